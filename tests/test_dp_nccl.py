"""Data-parallel training over NCCL on 2 GPUs (skipped with fewer): DP over two half-batches must equal the
single-process run on the full batch up to summation order."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu]
two_gpus = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, precision, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import numpy as np
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import dcll_oracle as O
    from util_build import build_pair
    from snn_modulation_classification_b200.data.utils import iq2spiketrain
    B, K, T, W, burnin = 16, 24, 7, 16, 3
    g = torch.Generator().manual_seed(4)
    xs = (torch.randn(B, 2, 1, 1024, generator=g) * 0.4).float()
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K)
    lo, hi = rank * B // world, (rank + 1) * B // world
    net, _ = build_pair("radio_ml_conv", (1, W, W), B // world, K, arp=1.0, burnin=burnin)
    net.set_precision(precision)
    np.random.seed(1)
    cells, _ = iq2spiketrain(xs[lo:hi], y[lo:hi], out_w=W, out_h=W, max_duration=T, as_cells=True)
    net.reset()
    net.learn_window_dp(cells, y[lo:hi].cuda())
    torch.cuda.synchronize()
    ok, worst = True, 0.0
    # identical replicas after the window
    for s in net.dcll_slices:
        w = s.dclllayer.i2h.weight.detach().clone()
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w)
        ok = ok and all(torch.equal(ws[0], t) for t in ws)
    if rank == 0:
        full, _ = build_pair("radio_ml_conv", (1, W, W), B, K, arp=1.0, burnin=burnin)
        full.set_precision(precision)
        np.random.seed(1)
        cells_f, _ = iq2spiketrain(xs, y, out_w=W, out_h=W, max_duration=T, as_cells=True)
        full.reset()
        full.learn_window(cells_f, y.cuda())
        for a, b in zip(net.dcll_slices, full.dcll_slices):
            d = float((a.dclllayer.i2h.weight - b.dclllayer.i2h.weight).abs().max()) / 1e-6
            worst = max(worst, d)
            ok = ok and len(a.clout) == len(b.clout) == T - burnin + 1
    out_q.put((rank, ok, worst))
    dist.destroy_process_group()


@two_gpus
@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16x2"])
def test_dp_two_gpus_matches_single_process(precision):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, precision, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    assert all(r[1] for r in res), res
    # 5 training steps, error x2 per step from summation-order differences
    assert res[0][2] <= (0.3 if precision == "fp32" else 2.0), res


# ---------------------------------------------------------------------------------------------------------------------
# The C data-parallel driver on ONE GPU: a single-rank NCCL communicator (the all-reduce is the identity), so this runs
# wherever the other GPU tests run.  dcll_net_window_dp -- side stream, events, bucket layout, Adam from the bucket -- must
# then reproduce learn_window bit for bit: the gradients are the same numbers, only the route to the Adam step differs.
# ---------------------------------------------------------------------------------------------------------------------
def _worker_single(port, precision, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import numpy as np
    import torch.distributed as dist
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    from oracle import dcll_oracle as O
    from util_build import build_pair
    from snn_modulation_classification_b200.data.utils import iq2spiketrain
    B, K, T, W, burnin = 12, 24, 9, 16, 3
    g = torch.Generator().manual_seed(4)
    xs = (torch.randn(B, 2, 1, 1024, generator=g) * 0.4).float()
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K).cuda()
    nets = []
    for _ in range(2):
        n, _ = build_pair("radio_ml_conv", (1, W, W), B, K, arp=1.0, burnin=burnin)
        nets.append(n.set_precision(precision))
    np.random.seed(1)
    cells, _ = iq2spiketrain(xs, y, out_w=W, out_h=W, max_duration=T, as_cells=True)
    for n in nets:
        n.reset()
    nets[0].learn_window(cells, y)
    nets[1].learn_window_dp(cells, y)
    torch.cuda.synchronize()
    ok = True
    for a, b in zip(nets[0].dcll_slices, nets[1].dcll_slices):
        ok = ok and torch.equal(a.dclllayer.i2h.weight, b.dclllayer.i2h.weight) and torch.equal(a.dclllayer.i2h.bias, b.dclllayer.i2h.bias)
        ok = ok and torch.equal(a.dclllayer.i2h.state.eps1, b.dclllayer.i2h.state.eps1)
        ok = ok and np.array_equal(np.array(a.clout), np.array(b.clout)) and a.iter == b.iter
        if a.dclllayer.output_layer:
            ok = ok and torch.equal(a.dclllayer.output_.weight, b.dclllayer.output_.weight)
            ok = ok and torch.equal(a.dclllayer.output_.bias, b.dclllayer.output_.bias)
        sa, sb = a.optimizer.state[a.dclllayer.i2h.weight], b.optimizer.state[b.dclllayer.i2h.weight]
        ok = ok and float(sa["step"]) == float(sb["step"]) == T - burnin + 1 and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"])
    out_q.put(ok)
    dist.destroy_process_group()


single_gpu = pytest.mark.skipif(torch.cuda.device_count() < 1, reason="needs a GPU")


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16x2"])
def test_dp_driver_single_rank_equals_learn_window(precision):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker_single, args=(port, precision, q))
    p.start()
    ok = q.get(timeout=600)
    p.join(timeout=120)
    assert ok
