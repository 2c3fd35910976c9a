#!/usr/bin/env python
"""Summarise ncu outputs into the text files kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof_conv_fwd_r01.ncu-rep > profiles/r01_ncu_conv_fwd.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                name = re.sub(r"\(.*", "", d["Kernel Name"].replace("void ", "")).replace("dcll::", "")
                agg[name][0] += 1
                agg[name][1] += float(d["Metric Value"].replace(",", "")) * scale.get(d["Metric Unit"], 1.0)
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-46s %6s %12s %10s %7s" % ("kernel", "n", "total_ms", "avg_ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-46s %6d %12.3f %10.4f %6.1f%%" % (k[:46], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none; one block per captured launch (%s)" % path)
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        for k in KEYS:
            if k in d and d[k] != "":
                print("%-72s %s %s" % (k, d[k][:90], u[k]))
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
