#!/usr/bin/env python
"""Benchmark of the DCLL hot path: RadioML IQ windows/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one batch of synthetic IQ windows pushed through the whole hot path: IQ->spike encoding
followed by T = 1024 timesteps of radio_ml_conv DCLL training (forward, local loss gradient, weight gradient
and Adam step per layer per timestep), or of inference for the infer workloads.  Prints ONE JSON line.

  value        windows/s with the IQ records already resident in HBM (encode + T timesteps timed)
  e2e          the same through the public API with HOST buffers: pinned-host IQ -> H2D -> iq2spiketrain ->
               ConvNetwork.learn_window -> device vote -> D2H of the per-sample predictions, all timed
  roofline     dominant kernel (convolution of a 32->32 layer) from CUDA events sampled inside the timed region
  cpu_baseline the oracle port (same operator sequence as the reference: F.conv2d / autograd / torch.optim.Adam)
               on the box's host cores, on a bounded sample of the same workload

--impl reference times that CPU port alone (the reference is pure Python/PyTorch and is not shipped to the GPU
box; oracle/ restates it operator for operator and is pinned bit-exactly against it, tests/test_oracle_vs_reference.py).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (spec, resolution, per-GPU batch, train, arp, burnin)
    "radio_ml_conv_train_128x128_B64": ("radio_ml_conv", 128, 64, True, 0.0, 50),   # train.py argparse defaults
    "radio_ml_conv_train_16x16_B64": ("radio_ml_conv", 16, 64, True, 0.0, 50),
    "radio_ml_conv_train_16x16_B512_arp": ("radio_ml_conv", 16, 512, True, 1.0, 20),  # scripts/train_radio_ml.sh
    "radio_ml_conv_infer_16x16_B4096": ("radio_ml_conv", 16, 4096, False, 1.0, 20),
    "radio_ml_conv_infer_128x128_B64": ("radio_ml_conv", 128, 64, False, 0.0, 50),
    "radio_ml_conv_train_16x16_B1024": ("radio_ml_conv", 16, 1024, True, 0.0, 50),   # config 5 shard (8192 / 8)
}
DEFAULT = "radio_ml_conv_train_128x128_B64"
K_CLASSES, N_IQ = 24, 1024


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops_sustained"], tensor_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, src="fallback")


def flops_per_sample_timestep(res, train):
    # conv fwd 2*Cin*kh*kw*Cout*H'*W' per layer, read-outs 2*F*K (x2 on the last layer)
    hw = res * res
    conv = 2 * 49 * hw * (1 * 32 + 32 * 32 + 32 * 32)
    ro = 2 * (32 * hw) * K_CLASSES * 4
    return (conv + ro) * (2 if train else 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 5]
        os.unlink(self.f.name)
        sm = sorted(int(r[0]) for r in rows if r[0].strip().isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any("Active" == r[2 + i].strip() for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(rows[0][1]) if rows and rows[0][1].strip().isdigit() else None,
                "reasons": reasons, "samples": len(rows)}


def synth(batch, seed):
    """SURVEY section 8d synthetic input: x ~ N(0, 0.4^2) f32 (B,2,1,1024) in the loader layout, uniform labels."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(batch, 2, 1, N_IQ, generator=g) * 0.4).float()
    lab = torch.randint(0, K_CLASSES, (batch,), generator=g)
    y = torch.zeros(batch, K_CLASSES).scatter_(1, lab.unsqueeze(-1), 1)
    return x, y


# --------------------------------------------------------------------------------------------------------
# CPU port (reference arm / cpu_baseline)
# --------------------------------------------------------------------------------------------------------
def cpu_port_sample(wl, T, n_fwd=1, n_train=2, reps=1):
    """Times the oracle port on a bounded sample and extrapolates linearly in T.
    Returns (windows_per_s, description, seconds_spent)."""
    import numpy as np
    import torch
    from oracle import dcll_oracle as O
    spec, res, batch, train, arp, burnin = WORKLOADS[wl]
    torch.set_num_threads(os.cpu_count() or 1)
    specs = O.make_specs(O.BUILTIN_SPECS[spec], (1, res, res), K_CLASSES, wrp=arp)
    params = O.random_params(specs, seed=1)
    net = O.OracleNet(specs, params, batch, burnin=0 if train else burnin, backend="autograd")
    x, y = synth(batch, 1)
    t_enc_n = max(8, 4 + n_fwd + n_train)
    t0 = time.perf_counter()
    cells = O.encode_cells(x.numpy(), res, res, t_start=0, max_duration=t_enc_n)
    frames = torch.from_numpy(O.cells_to_frames(cells, res, res))
    enc_per_t = (time.perf_counter() - t0) / t_enc_n
    spent0 = time.perf_counter()
    net.reset()
    with torch.no_grad():
        net.test(frames[0])                                          # untimed warm-up (allocations, oneDNN primitives)
    t_fw = t_tr = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        with torch.no_grad():
            for i in range(n_fwd):
                net.test(frames[1 + i])
        t_fw += (time.perf_counter() - t0) / n_fwd
        if train:
            net.learn(frames[2], y)                                  # untimed: first backward builds Adam state
            t0 = time.perf_counter()
            for i in range(n_train):
                net.learn(frames[3 + i], y)
            t_tr += (time.perf_counter() - t0) / n_train
    t_fw, t_tr = t_fw / reps, t_tr / reps
    if train:
        window_s = enc_per_t * T + (burnin - 1) * t_fw + (T - burnin + 1) * t_tr
        desc = ("oracle port (F.conv2d + autograd + torch.optim.Adam, %d threads): %d fwd-only + %d training timesteps "
                "of %s at B=%d timed, extrapolated linearly to T=%d (%d burn-in + %d training timesteps) + encode"
                % (torch.get_num_threads(), n_fwd, n_train, wl, batch, T, burnin - 1, T - burnin + 1))
    else:
        window_s = enc_per_t * T + T * t_fw
        desc = ("oracle port (%d threads): %d inference timesteps of %s at B=%d timed, extrapolated linearly to T=%d + encode"
                % (torch.get_num_threads(), n_fwd, wl, batch, T))
    return batch / window_s, desc, time.perf_counter() - spent0, dict(ms_fwd_timestep=1e3 * t_fw, ms_train_timestep=1e3 * t_tr)


def run_reference(a):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec, res, batch, train, arp, burnin = WORKLOADS[a.workload]
    vals = []
    big = res >= 64
    for i in range(a.warmup + a.steps):
        v, desc, _, extra = cpu_port_sample(a.workload, a.timesteps, n_fwd=2 if big else 8, n_train=6 if big else 64)
        if i >= a.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    unit = "windows/s"
    out = {"impl": "reference", "metric": "RadioML IQ windows/sec (DCLL %s)" % ("train" if train else "infer"),
           "value": v, "unit": unit, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": 1e3 * batch / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": a.workload, "timesteps": a.timesteps, "batch_per_gpu": batch, "resolution": res},
           "cpu_baseline": {"value": v, "unit": unit, "cores": torch.get_num_threads(), "kind": "port", "sample": desc},
           "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_json(out)


# --------------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------------
def make_args(arp=0.0):
    """train.py argparse defaults (ref train.py:75-90) as the namespace ConvNetwork reads."""
    import types
    return types.SimpleNamespace(netscale=1.0, alpha=0.92, alphas=0.85, alpharp=0.65, arp=arp, lc_ampl=0.5, random_tau=True)


NCU_SUMMARY = "r01_ncu_final_top_kernels.txt"


NCU_SUMMARY_MMA2 = "r01_ncu_conv_mma2.txt"


def ncu_traffic(kernel_substr, summary=None):
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel, from the committed `ncu --set full`
    summary of this round (profiles/); None when the capture is absent."""
    path = os.path.join(ROOT, "profiles", summary or NCU_SUMMARY)
    if not os.path.exists(path):
        return None
    cur, vals = None, {}
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    for line in open(path):
        parts = line.split()
        if line.startswith("Kernel Name"):
            cur = line
            if kernel_substr in cur and vals.get("done"):
                break
        elif cur and kernel_substr in cur and parts and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            vals[parts[0]] = float(parts[1]) * scale.get(parts[2], 1.0)
            if len(vals) == 2:
                vals["done"] = True
    if "dram__bytes_read.sum" in vals and "dram__bytes_write.sum" in vals:
        return vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
    return None


def build_net(wl):
    import numpy as np
    import torch
    from snn_modulation_classification_b200 import networks as N
    spec, res, batch, train, arp, burnin = WORKLOADS[wl]
    torch.manual_seed(1)
    np.random.seed(1)
    kw = dict(loss=torch.nn.SmoothL1Loss, opt=torch.optim.Adam, opt_param={"betas": [0.0, 0.95], "weight_decay": 10.0},
              learning_rates=[1e-6]) if train else dict(loss=None, opt=None, opt_param={}, learning_rates=None)
    net = N.ConvNetwork(make_args(arp), (1, res, res), batch, N.load_network_spec(spec), K_CLASSES,
                        act=torch.nn.Sigmoid(), burnin=burnin, **kw)
    net.reset(True)
    return net


def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from snn_modulation_classification_b200 import _lib
    from snn_modulation_classification_b200.data.utils import iq2spiketrain

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec, res, batch, train, arp, burnin = WORKLOADS[a.workload]
    T = a.timesteps
    net = build_net(a.workload)
    net.set_precision(a.precision)
    tc = any(sl.dclllayer.i2h.tensor_core_ok() for sl in net.dcll_slices)
    if world > 1:                                   # identical replicas: broadcast rank 0's parameters
        for p in net.state_dict().values():
            dist.broadcast(p, 0)
    x, y = synth(batch, 1 + rank)
    x_pin, y_pin = x.pin_memory(), y.pin_memory()
    x_dev, y_dev = x.cuda(), y.cuda()
    pred_pin = torch.empty(batch, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    enc = dict(out_w=res, out_h=res, min_I=-1, max_I=1, min_Q=-1, max_Q=1, max_duration=T, as_cells=True)

    def run_window(cells, labels):
        net.reset()
        if not train:
            net.test_window(cells)
        elif world > 1:
            net.learn_window_dp(cells, labels)
        else:
            net.learn_window(cells, labels)

    def step_device():
        cells, _ = iq2spiketrain(x_dev, y_dev, **enc)
        run_window(cells, y_dev)

    def step_e2e():
        xd = x_pin.to("cuda", non_blocking=True)
        yd = y_pin.to("cuda", non_blocking=True)
        cells, _ = iq2spiketrain(xd, yd, **enc)
        run_window(cells, yd)
        pred = net.dcll_slices[-1].clout.vote_device(K_CLASSES)
        pred_pin.copy_(pred, non_blocking=True)

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(steps):
            flush.zero_()                          # L2 flush between timed iterations (256 MiB > 126 MB L2)
            fn()
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    np.random.seed(1)
    for _ in range(a.warmup):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.lib.dcll_launch_count(1)
    _lib.check(_lib.lib.dcll_profile_enable(a.profile_every))
    ms = timed(step_device, a.steps)
    launches = int(_lib.lib.dcll_launch_count(0))
    prof = _lib.profile_read()
    _lib.check(_lib.lib.dcll_profile_enable(0))
    ms_e2e = timed(step_e2e, a.steps)
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * batch * a.steps / (ms / 1e3)
    e2e = world * batch * a.steps / (ms_e2e / 1e3)
    pk = peaks()
    # ---- roofline of the dominant kernel: the convolution of a 32->32 layer (layers 1 and 2 are identical).
    # In bf16x3 mode the conv_fwd bracket holds trace_image_kernel + conv_mma_kernel and the trace bracket the former alone.
    hw = res * res
    conv_flops = 2.0 * 32 * 49 * 32 * hw * batch                     # algorithmic FLOPs per launch
    # Quoted on the LAST layer: in the window driver layer 1's launch also carries layer 2's trace update (fused epilogue), and
    # layer 2's own trace pass is gone, so its conv_fwd bracket is the pure MMA kernel.
    lr = len(net.dcll_slices) - 1
    c_ms, c_n = prof.get(("conv_fwd", lr), (0.0, 0))
    t_ms, t_n = prof.get(("trace", lr), (0.0, 0))
    avg_ms = (c_ms / c_n - (t_ms / t_n if t_n else 0.0)) if c_n else None
    achieved = conv_flops / (avg_ms * 1e-3) / 1e12 if avg_ms else None
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    # issue floor of this decomposition: 2 M-tiles x 49 taps x 2 k-chunks x 2 MMAs per 16x16-position tile, ~44 cycles per
    # SS-mode M=128 MMA with N <= 64 (A-operand read; tools/mma_bench.cu), tiles spread over 148 SMs
    n_tiles = batch * ((res + 15) // 16) ** 2
    floor_ms = (n_tiles + 147) // 148 * 392 * 44 / (sm_mhz * 1e3)
    # The last layer takes conv_mma2_kernel (row-interleaved N-concatenation, conv_fwd_tc.cu) when its 32 x 8 tiles waste
    # <= 15 % of the plane: per tile 14 ring stages of 6 x (N=128: 64 + N=64: 49) + 2 x (N=64: 49 + N=32: 45) cycles.
    # Its conv_fwd bracket also holds the ~3 us weight re-layout launch that precedes it.
    t32, t8 = (res + 31) // 32, (res + 7) // 8
    mma2 = tc and os.environ.get("DCLL_CONV_MMA2", "1") != "0" and (t32 * 32) * (t8 * 8) <= 1.15 * res * res
    if mma2:
        floor_ms = (batch * t32 * t8 + 147) // 148 * 14 * (6 * (64 + 49) + 2 * (49 + 45)) / (sm_mhz * 1e3)
    summary = NCU_SUMMARY_MMA2 if mma2 else NCU_SUMMARY
    kname = ("conv_mma2_kernel (layer 2: 32->32 ch, tcgen05 split-bf16 x3, N = 128/64 row-interleaved MMAs, TMEM accumulators)"
             if mma2 else "conv_mma_kernel<7,7,32,32> (layer 2: 32->32 ch, tcgen05 split-bf16 x3, TMEM accumulators)")
    roofline = {"kernel": (kname if tc else "conv_fwd_kernel<7,7,...> (layer 2: 32->32 ch, FP32 FMA path)"), "bound": "tensor",
                "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": achieved / pk["tensor"] if achieved else None,
                "traffic": (ncu_traffic("conv_mma2_kernel" if mma2 else "conv_mma_kernel", summary)
                            if (tc and res == 128 and batch == 64) else None),
                "traffic_unit": "bytes per launch (dram read + write, ncu --set full, profiles/%s); algorithmic: 0.40e9 "
                                "(operand image in, spikes + pv out)" % summary,
                "peak_source": "%s bf16 dense sustained (MEASURED_PEAKS.json)" % pk["src"],
                "avg_launch_ms": avg_ms, "launches_sampled": c_n, "algorithmic_flops_per_launch": conv_flops,
                "fp32_fma_peak_tflops_at_clock": fp32_peak, "frac_of_fp32_fma_peak": achieved / fp32_peak if achieved else None,
                "executed_tensor_flops_per_launch": conv_flops * (3 if tc else 0),
                "frac_executed": (3 * achieved / pk["tensor"]) if (achieved and tc) else None,
                "mma_issue_floor_ms": floor_ms if tc else None,
                "frac_of_issue_floor": (floor_ms / avg_ms) if (tc and avg_ms) else None,
                "note": ("achieved counts ALGORITHMIC conv FLOPs; the split-bf16 mode executes 3 bf16 products per FLOP "
                         "(frac_executed = tensor-pipe share actually used). With Cout = 32 the MMAs are short (N = 128 / 64 "
                         "in conv_mma2_kernel, N = 64 / 32 in conv_mma_kernel) and bound by the shared-memory operand reads "
                         "(4 KB of A per MMA; 64 / 49 / 45 cycles for N = 128 / 64 / 32, measured), not by math: "
                         "mma_issue_floor_ms is the floor of the decomposition in use." if tc else
                         "FP32-exact parity mode runs on the CUDA-core FMA pipe")}
    per_class = {}
    for (name, layer), (tms, n) in sorted(prof.items()):
        per_class["%s[l%d]" % (name, layer)] = {"avg_ms": tms / n, "samples": n}

    out = {"metric": "RadioML IQ windows/sec (DCLL %s)" % ("train" if train else "infer"), "value": value,
           "unit": "windows/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16x3 (split-bf16 operands on tcgen05, fp32 accumulate; traces/neuron/update fp32)" if tc else "f32",
           "data": "synthetic",
           "config": {"workload": a.workload, "precision": a.precision, "network": spec + ".yaml", "timesteps": T, "batch_per_gpu": batch,
                      "global_batch": batch * world, "resolution": "%dx%d" % (res, res), "arp": arp, "burnin": burnin,
                      "parallelism": "dp%d (batch-sharded, NCCL allreduce of local-layer grads per timestep)" % world
                      if world > 1 else "single GPU",
                      "l2": "256 MiB flush buffer written between timed steps; per-step working set (state ping-pong "
                            "%.0f MB/layer) %s L2" % (2 * 2 * 4 * 32 * hw * batch / 1e6, ">" if res >= 64 else "<")},
           "sample_timesteps_per_s": value * T,
           "model_tflops": value * T * flops_per_sample_timestep(res, train) / 1e12,
           "e2e": {"value": e2e, "unit": "windows/s", "ms_per_step": ms_e2e / a.steps,
                   "h2d_bytes_per_step": int(x_pin.numel() * 4 + y_pin.numel() * 4),
                   "d2h_bytes_per_step": int(pred_pin.numel() * 4)},
           "gpu_launches": launches, "roofline": roofline, "kernel_ms": per_class, "clocks": clocks}
    if world == 1 and not a.no_cpu:
        big = res >= 64
        # ~10-20 s of CPU work on the bounded sample (16 host cores: ~0.45 s / 0.67 s per 128x128 timestep)
        v, desc, spent, extra = cpu_port_sample(a.workload, T, n_fwd=4 if big else 16, n_train=12 if big else 256,
                                                reps=1 if big else 3)
        out["cpu_baseline"] = {"value": v, "unit": "windows/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": desc, "seconds": spent, **extra}
    emit_json(out)
    if world > 1:
        dist.destroy_process_group()


class _StdoutGuard:
    """Everything libraries print to fd 1 while the benchmark runs (e.g. NCCL's version banner) goes to stderr;
    only the final JSON line reaches the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


_GUARD = None


def emit_json(obj):
    line = json.dumps(obj)
    if _GUARD is not None:
        _GUARD.emit(line)
    else:
        print(line, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument("--timesteps", type=int, default=1024)
    ap.add_argument("--profile-every", type=int, default=31, dest="profile_every")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3"],
                    help="fp32: FP32-exact parity mode on the FMA pipe; bf16x3: tcgen05 split-bf16 (headline mode)")
    ap.add_argument("--burnin", type=int, default=None, help="override the workload's burn-in (profiling runs)")
    a = ap.parse_args()
    if a.burnin is not None:
        w = list(WORKLOADS[a.workload])
        w[5] = a.burnin
        WORKLOADS[a.workload] = tuple(w)
    global _GUARD
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    with _StdoutGuard() as g:
        _GUARD = g
        if a.impl == "reference":
            run_reference(a)
        else:
            run_b200(a)
    _GUARD = None


if __name__ == "__main__":
    main()
