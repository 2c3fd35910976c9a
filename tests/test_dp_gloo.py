"""Data-parallel path, host logic: world_size 2 over gloo on CPU (no GPU needed).

The kernels cannot run on CPU, so the per-rank compute is the oracle's closed-form local gradient; what is
tested is the product's bucket layout and averaging (`networks.grad_bucket`, `networks.allreduce_mean`) and the
claim they rest on: the mean over ranks of the per-shard local gradients equals the global-batch gradient, so
every rank applies the identical Adam step.
"""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import dcll_oracle as O
    from snn_modulation_classification_b200 import networks as N
    from snn_modulation_classification_b200.dcll import pytorch_libdcll as L
    L.device = N.device = "cpu"
    import numpy as np
    from util_build import make_args, state_dict_from_params
    B, K, im = 8, 24, (1, 8, 8)
    specs = O.make_specs(O.BUILTIN_SPECS["radio_ml_conv"], im, K)
    params = O.random_params(specs, seed=3)
    torch.manual_seed(0)
    np.random.seed(0)
    net = N.ConvNetwork(make_args(), im, B // world, N.load_network_spec("radio_ml_conv"), K, act=torch.nn.Sigmoid(),
                        loss=torch.nn.SmoothL1Loss, opt=torch.optim.Adam, opt_param={"betas": [0.0, 0.95], "weight_decay": 10.0},
                        learning_rates=[1e-6], burnin=0)
    net.load_state_dict(state_dict_from_params(params))
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(3, B, *im, generator=g) < 0.15).float()
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K)
    lo, hi = rank * B // world, (rank + 1) * B // world
    ok = True
    states = [O.zero_state(s, B // world) for s in specs]
    full_states = [O.zero_state(s, B) for s in specs]
    for t in range(3):
        inp, inp_full = x[t, lo:hi], x[t]
        for i, (sp, p) in enumerate(zip(specs, params)):
            fo = O.conv_step_fwd(sp, p, states[i], inp)
            states[i] = fo.state
            gr = O.conv_local_grads(sp, p, fo, y[lo:hi])
            flat, views = N.grad_bucket(net.dcll_slices[i].dclllayer, "cpu")
            parts = [gr.gW, gr.gb] + ([gr.gWout, gr.gbout] if sp.output_layer else [])
            assert [v.numel() for v in views] == [q.numel() for q in parts]
            for v, q in zip(views, parts):
                v.copy_(q.reshape(-1))
            N.allreduce_mean(flat)
            # single-process, global batch
            fo_f = O.conv_step_fwd(sp, p, full_states[i], inp_full)
            full_states[i] = fo_f.state
            gf = O.conv_local_grads(sp, p, fo_f, y)
            want = torch.cat([q.reshape(-1) for q in ([gf.gW, gf.gb] + ([gf.gWout, gf.gbout] if sp.output_layer else []))])
            err = float((flat - want).abs().max()) / float(want.abs().max())
            ok = ok and err < 1e-5
            # identical buckets on all ranks -> identical Adam steps
            other = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(other, flat)
            ok = ok and all(torch.equal(o, flat) for o in other)
            inp, inp_full = fo.spikes, fo_f.spikes
    out_q.put((rank, ok))
    dist.destroy_process_group()


def test_dp_bucket_mean_equals_global_batch_gradient():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
