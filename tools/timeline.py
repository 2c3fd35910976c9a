#!/usr/bin/env python
"""In-kernel stopwatch of the persistent tensor-core kernels (csrc/common.cuh TL_*), on the bench workload.

    DCLL_TIMELINE=1 python tools/timeline.py [--workload W] [--timesteps 24] [--burnin 4]

Runs one short window, then prints, per kernel kind, the mean / max over the CTAs of each slot in microseconds at the nominal
1.965 GHz (the slots are SM cycles) and as a share of the CTA's total.  The last launch of a kind wins, so with the default
kernel selection conv_mma<32> is layer 1, conv_mma2 the last layer and wgrad_tc2 the last layer.
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DCLL_TIMELINE", "1")

KINDS = ["conv_mma<32>", "conv_mma<1>", "conv_mma2", "wgrad_tc2"]
SLOTS = ["total", "prologue(entry->pdl)", "iss wait a_full / wg2 iss0 wait full", "iss wait acc_empty", "iss wait w_full / wg2 iss1 wait full",
         "iss loop / wg2 iss0 loop", "epi wait acc_full", "epi loop / wg2 iss1 loop", "wprod wait w_empty", "aprod wait (a_)empty",
         "entry->first mma", "drain (wg2)", "epi post-TMEM (math+memory)", "-", "-", "-"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="radio_ml_conv_train_128x128_B64")
    ap.add_argument("--timesteps", type=int, default=24)
    ap.add_argument("--burnin", type=int, default=4)
    ap.add_argument("--ghz", type=float, default=1.965)
    a = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from snn_modulation_classification_b200 import _lib
    from snn_modulation_classification_b200.data.utils import iq2spiketrain

    w = list(bench.WORKLOADS[a.workload])
    w[5] = a.burnin
    bench.WORKLOADS[a.workload] = tuple(w)
    spec, res, batch, train, arp, burnin, _ = bench.workload(a.workload, 1)
    net = bench.build_net(a.workload, 1)
    net.set_precision(os.environ.get("DCLL_PRECISION", "f16x2"))
    x, y = bench.synth(batch, 1)
    np.random.seed(1)
    cells, _ = iq2spiketrain(x.cuda(), y.cuda(), out_w=res, out_h=res, min_I=-1, max_I=1, min_Q=-1, max_Q=1, max_duration=a.timesteps,
                             as_cells=True)
    for _ in range(2):
        net.reset()
        net.learn_window(cells, y.cuda()) if train else net.test_window(cells)
    torch.cuda.synchronize()
    n = 4 * 148 * 16
    buf = (C.c_uint64 * n)()
    got = _lib.lib.dcll_debug_timeline(buf, n)
    assert got == n, _lib.lib.dcll_last_error()
    v = np.frombuffer(buf, dtype=np.uint64).astype(np.float64).reshape(4, 148, 16)
    for k, name in enumerate(KINDS):
        tot = v[k, :, 0]
        act = tot > 0
        if not act.any():
            continue
        print("== %s: %d CTAs, total mean %.1f us  max %.1f us (at %.3f GHz)" % (name, act.sum(), tot[act].mean() / a.ghz / 1e3,
                                                                               tot[act].max() / a.ghz / 1e3, a.ghz))
        t0, t1 = v[k, act, 13], v[k, act, 14]
        ghz = (tot[act] / np.maximum(t1 - t0, 1.0)).mean()
        print("   globaltimer: kernel span (first entry -> last exit) %.1f us; CTA entry skew %.1f us; exit skew %.1f us; SM clock %.3f GHz"
              % ((t1.max() - t0.min()) / 1e3, (t0.max() - t0.min()) / 1e3, (t1.max() - t1.min()) / 1e3, ghz))
        idx = np.nonzero(act)[0]
        dur = (t1 - t0) / 1e3
        print("   CTA lifetime (us): min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f" % tuple(np.percentile(dur, [0, 10, 50, 90, 100])))
        if name == "wgrad_tc2":
            ra = dur[idx < 78]
            rb = dur[idx >= 78]
            if len(ra) and len(rb):
                print("   role A (CTAs 0..77, 39 : 35 split) mean %.1f max %.1f | role B mean %.1f max %.1f" % (ra.mean(), ra.max(), rb.mean(), rb.max()))
        order = np.argsort(dur)
        print("   slowest CTAs:", ", ".join("%d:%.0f" % (idx[i], dur[i]) for i in order[-6:]), "| fastest:", ", ".join("%d:%.0f" % (idx[i], dur[i]) for i in order[:6]))
        for s in range(1, 13):
            col = v[k, act, s]
            if col.max() == 0:
                continue
            print("   %-40s mean %8.1f us  max %8.1f us   %5.1f %% of total" % (SLOTS[s], col.mean() / a.ghz / 1e3, col.max() / a.ghz / 1e3,
                                                                               100 * col.mean() / tot[act].mean()))


if __name__ == "__main__":
    main()
