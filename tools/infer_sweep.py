#!/usr/bin/env python
"""Inference sweep at 16x16: windows/s of ConvNetwork.test_window over B x T.

    python tools/infer_sweep.py > profiles/r01_infer_sweep_16x16.jsonl

One JSON line per (B, T): device-resident IQ records -> iq2spiketrain (cells) -> test_window (dcll_infer_stack16 +
dcll_conv_readout_rows in bf16x3 mode) -> device vote.  CUDA events, 1 warm-up + 3 timed windows, L2 flushed between.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from snn_modulation_classification_b200.data.utils import iq2spiketrain


def run(B, T, precision="bf16x3"):
    bench.WORKLOADS["sweep"] = ("radio_ml_conv", 16, B, False, 1.0, 20)
    net = bench.build_net("sweep")
    net.set_precision(precision)
    x, y = bench.synth(B, 1)
    x, y = x.cuda(), y.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    enc = dict(out_w=16, out_h=16, min_I=-1, max_I=1, min_Q=-1, max_Q=1, max_duration=T, as_cells=True)

    def step():
        cells, _ = iq2spiketrain(x, y, **enc)
        net.reset()
        net.test_window(cells)
        return net.dcll_slices[-1].clout.vote_device(bench.K_CLASSES)

    np.random.seed(1)
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    e0.record()
    for _ in range(steps):
        flush.zero_()
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"B": B, "T": T, "precision": precision, "ms_per_window_batch": ms, "windows_per_s": B / (ms / 1e3),
            "sample_timesteps_per_s": B * T / (ms / 1e3), "multi_timestep_kernel": bool(net._stack16_ok(iq2spiketrain(x, y, **enc)[0]))}


if __name__ == "__main__":
    for B in (256, 1024, 4096, 16384, 65536):
        for T in (64, 256, 512):
            if B * T > 65536 * 256:      # keep the sweep within a few GPU-minutes
                continue
            print(json.dumps(run(B, T)), flush=True)
    print(json.dumps(run(4096, 256, "fp32")), flush=True)   # per-timestep FP32 path for comparison
