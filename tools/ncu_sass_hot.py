#!/usr/bin/env python
"""Where a kernel's warps wait, from the SASS page of an ncu report (no GPU needed).

    python tools/ncu_sass_hot.py gpurun_out/r12/prof.ncu-rep conv_mma2_kernel [top_n]

Prints the stall-reason totals of the warp-state samples, the executed-instruction mix by opcode and the top_n
instructions by samples (address order kept, so neighbouring lines show the loop they belong to).
"""
import collections
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + pat, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print("#", rows[0][1] if rows and len(rows[0]) > 1 else pat)
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    mix = collections.Counter()
    insts = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":      # next launch of the same kernel: the first one is enough
            break
        if len(r) != len(hdr) or r == hdr:
            continue
        smp = int(r[col["# Samples"]] or 0)
        ex = int(r[col["Instructions Executed"]] or 0)
        src = r[col["Source"]].strip()
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        mix[op.split(".")[0]] += ex
        st = {h: int(r[col[h]] or 0) for h in stall_cols}
        for h, v in st.items():
            tot[h] += v
        insts.append((smp, ex, src, st, len(insts)))
    n_s = sum(i[0] for i in insts)
    n_e = sum(i[1] for i in insts)
    print("samples %d, warp-instructions executed %d, SASS lines %d" % (n_s, n_e, len(insts)))
    print("stall reasons:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(1, n_s)) for h, v in tot.most_common(8)))
    print("instruction mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, n_e)) for k, v in mix.most_common(14)))
    print("top %d instructions by samples:" % top_n)
    for smp, ex, src, st, idx in sorted(sorted(insts, key=lambda i: -i[0])[:top_n], key=lambda i: i[4]):
        why = max(st.items(), key=lambda kv: kv[1])
        print("  #%-5d %6d smp (%4.1f%%) %9d exec  %-14s %s" % (idx, smp, 100.0 * smp / max(1, n_s), ex, why[0][6:], src[:90]))


if __name__ == "__main__":
    main()
