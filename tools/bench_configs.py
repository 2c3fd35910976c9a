#!/usr/bin/env python
"""Throughput of the other SURVEY 8(d) training configurations through the public API (one JSON line each):

  config 2  mnist_conv.yaml, B = 128, K = 10, dense Bernoulli frames [T,B,1,28,28] (rate = gain*pixel/1000, gain 100), T = 500
  config 3  radio_ml_conv_ref.yaml as a 7-layer DCLL spec at 128x128 with int8-quantised weights, B = 32, T = 64

    python tools/bench_configs.py > profiles/r01_bench_other_configs.jsonl

Both run on the FP32 FMA convolution kernels (channel counts / kernel shapes without a tensor-core instantiation) with the
tcgen05 read-out; CUDA events, 1 warm-up + 2 timed windows.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from snn_modulation_classification_b200 import networks as N
from snn_modulation_classification_b200.data.utils import iq2spiketrain
from snn_modulation_classification_b200.quant import enable_quantized_weights


def build(spec, im, batch, K, burnin):
    torch.manual_seed(1)
    np.random.seed(1)
    net = N.ConvNetwork(bench.make_args(0.0), im, batch, N.load_network_spec(spec), K, act=torch.nn.Sigmoid(),
                        loss=torch.nn.SmoothL1Loss, opt=torch.optim.Adam,
                        opt_param={"betas": [0.0, 0.95], "weight_decay": 10.0}, learning_rates=[1e-6], burnin=burnin)
    net.reset(True)
    net.set_precision("bf16x3")
    return net


def timed(fn, steps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def mnist(B=128, T=500):
    net = build("mnist_conv", (1, 28, 28), B, 10, 50)
    g = torch.Generator().manual_seed(1)
    img = torch.rand(B, 1, 28, 28, generator=g)
    frames = (torch.rand(T, B, 1, 28, 28, generator=g) < img * (100.0 / 1000.0)).float().cuda()
    lab = torch.randint(0, 10, (B,), generator=g)
    y = torch.zeros(B, 10).scatter_(1, lab.unsqueeze(-1), 1).cuda()

    def step():
        net.reset()
        net.learn_window(frames, y)

    ms = timed(step)
    return {"config": "mnist_conv.yaml train, B=%d, T=%d, 28x28 dense Bernoulli frames, K=10" % (B, T), "ms_per_window_batch": ms,
            "windows_per_s": B / (ms / 1e3), "sample_timesteps_per_s": B * T / (ms / 1e3),
            "tensor_core_layers": [bool(s.dclllayer.i2h.tensor_core_ok()) for s in net.dcll_slices]}


def radio_ref_quant(B=32, T=64):
    net = build("radio_ml_conv_ref", (1, 128, 128), B, bench.K_CLASSES, 16)
    enable_quantized_weights(net)
    x, y = bench.synth(B, 1)
    x, y = x.cuda(), y.cuda()
    enc = dict(out_w=128, out_h=128, min_I=-1, max_I=1, min_Q=-1, max_Q=1, max_duration=T, as_cells=True)

    def step():
        cells, _ = iq2spiketrain(x, y, **enc)
        net.reset()
        net.learn_window(cells, y)

    ms = timed(step)
    return {"config": "radio_ml_conv_ref.yaml as a DCLL spec (7 layers, 1x3 kernels, 64 ch), int8-quantised weights, train, "
                      "B=%d, T=%d, 128x128" % (B, T), "ms_per_window_batch": ms, "windows_per_s": B / (ms / 1e3),
            "sample_timesteps_per_s": B * T / (ms / 1e3), "layers": len(net.dcll_slices)}


if __name__ == "__main__":
    print(json.dumps(mnist()), flush=True)
    print(json.dumps(radio_ref_quant()), flush=True)
