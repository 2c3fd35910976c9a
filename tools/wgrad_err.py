#!/usr/bin/env python
"""Raw weight-gradient error of the tensor-core kernels against the oracle's closed form (what tests/test_gpu_tensorcore.py
test_tc_weight_gradient asserts), printed per layer: max |err| / max |g| and the worst element.  A/B aid:

    DCLL_WG2_PAIR=0 python tools/wgrad_err.py ; DCLL_WG2_PAIR=1 python tools/wgrad_err.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from oracle import dcll_oracle as O
from util_build import force_state, make_args, state_dict_from_params


def main():
    from snn_modulation_classification_b200 import networks as N
    im, B, arp, K = (128, 128), int(os.environ.get("B", "8")), 0.0, 24
    specs = O.make_specs(O.BUILTIN_SPECS["radio_ml_conv"], (1,) + im, K, wrp=arp)
    params = O.random_params(specs, seed=2)
    sd = state_dict_from_params(params)
    net = N.ConvNetwork(make_args(arp), (1,) + im, B, N.load_network_spec("radio_ml_conv"), K, act=torch.nn.Sigmoid(),
                        loss=torch.nn.SmoothL1Loss, opt=torch.optim.SGD, opt_param={}, learning_rates=[0.0], burnin=0)
    net.load_state_dict(sd)
    net = net.to("cuda")
    net.reset(True)
    net.load_state_dict(sd)
    net.set_precision(os.environ.get("DCLL_PRECISION", "bf16x3"))
    onet = O.OracleNet(specs, params, B, burnin=0)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(3, B, 1, *im, generator=g) < 0.1).float()
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K)
    for t in range(3):
        force_state(net, onet)
        fos, grads, inp = [], [], x[t]
        for i, sp in enumerate(specs):
            fo = O.conv_step_fwd(sp, params[i], onet.states[i], inp)
            grads.append(O.conv_local_grads(sp, params[i], fo, y))
            fos.append(fo)
            onet.states[i], inp = fo.state, fo.spikes
        for i, s in enumerate(net.dcll_slices):
            s.train_dcll(x[t].cuda() if i == 0 else fos[i - 1].spikes.cuda(), y.cuda(), regularize=False)
            gw, ref = s.dclllayer.i2h.weight.grad.cpu(), grads[i].gW
            err = (gw - ref).abs()
            idx = torch.nonzero(err == err.max())[0].tolist()
            gb, refb = s.dclllayer.i2h.bias.grad.cpu(), grads[i].gb
            print("t %d layer %d: gW max err %.3e of scale %.3e (rel %.2e), mean rel %.2e, worst at %s; gb rel %.2e" % (
                t, i, float(err.max()), float(ref.abs().max()), float(err.max() / ref.abs().max()), float(err.mean() / ref.abs().max()),
                idx, float((gb - refb).abs().max() / refb.abs().max())))
            # per (kh, kw) maximum error: a structural mistake shows as one tap standing out
            if i > 0:
                per_tap = err.amax(dim=(0, 1)) / ref.abs().max()
                print("   per-tap max rel err (rows kh):", "; ".join(" ".join("%.0e" % float(v) for v in row) for row in per_tap))


if __name__ == "__main__":
    main()
