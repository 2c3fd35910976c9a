"""Test helpers: build the CUDA-backed network next to an oracle network with identical parameters."""
import types

import numpy as np
import torch

from oracle import dcll_oracle as O


def state_dict_from_params(params):
    """Oracle parameters -> state_dict with the reference's key names."""
    sd = {}
    for i, p in enumerate(params):
        pre = "dcll_slices.%d.dclllayer." % i
        sd[pre + "i2h.weight"], sd[pre + "i2h.bias"] = p.weight.detach().clone(), p.bias.detach().clone()
        sd[pre + "i2h.alpha"], sd[pre + "i2h.tau_m__dt"] = p.alpha.clone(), p.tau_m.clone()
        sd[pre + "i2h.alphas"], sd[pre + "i2h.tau_s__dt"] = p.alphas.clone(), p.tau_s.clone()
        sd[pre + "i2o.weight"], sd[pre + "i2o.bias"] = p.wo.clone(), p.bo.clone()
        if p.wout is not None:
            sd[pre + "output_.weight"], sd[pre + "output_.bias"] = p.wout.detach().clone(), p.bout.detach().clone()
    return sd


def make_args(arp=0.0, random_tau=True, netscale=1.0):
    return types.SimpleNamespace(netscale=netscale, alpha=0.92, alphas=0.85, alpharp=0.65, arp=arp, lc_ampl=0.5,
                                 random_tau=random_tau)


def build_pair(spec_name, im_dims, B, K, arp=0.0, burnin=3, seed=0, train=True, state_dict=None, lr=1e-6,
               backend="closed"):
    """(cuda ConvNetwork, OracleNet) with identical parameters."""
    from snn_modulation_classification_b200 import networks as N

    specs = O.make_specs(O.BUILTIN_SPECS[spec_name], im_dims, K, wrp=arp)
    if state_dict is None:
        params = O.random_params(specs, seed=seed)
        state_dict = state_dict_from_params(params)
    else:
        params = O.params_from_state_dict(state_dict, len(specs))
    onet = O.OracleNet(specs, params, B, burnin=burnin, lrs=(lr,), backend=backend)
    torch.manual_seed(seed)
    np.random.seed(seed)
    convs = N.load_network_spec(spec_name)
    if train:
        net = N.ConvNetwork(make_args(arp), im_dims, B, convs, K, act=torch.nn.Sigmoid(), loss=torch.nn.SmoothL1Loss,
                            opt=torch.optim.Adam, opt_param={"betas": [0.0, 0.95], "weight_decay": 10.0},
                            learning_rates=[lr], burnin=burnin)
    else:
        net = N.ConvNetwork(make_args(arp), im_dims, B, convs, K, act=torch.nn.Sigmoid(), loss=None, opt=None,
                            opt_param={}, learning_rates=None, burnin=burnin)
    missing = net.load_state_dict({k: v.clone() for k, v in state_dict.items()}, strict=True)
    net = net.to("cuda")
    net.reset(True)
    # the refractory core re-randomises tau on every init_state (reference quirk): restore the loaded values
    net.load_state_dict({k: v.clone() for k, v in state_dict.items()}, strict=True)
    return net, onet


def force_state(net, onet):
    """Teacher forcing: copy the oracle's neuron state and weights into the CUDA network."""
    for s, st, p, sl in zip(net.dcll_slices, onet.states, onet.params, onet.slots):
        i2h = s.dclllayer.i2h
        vals = [st.eps0.cuda(), st.eps1.cuda()] + ([st.arp.cuda()] if st.arp is not None else [])
        i2h.state = i2h.NeuronState(*[v.contiguous().clone() for v in vals])
        with torch.no_grad():
            i2h.weight.copy_(p.weight.detach())
            i2h.bias.copy_(p.bias.detach())
            if p.wout is not None:
                s.dclllayer.output_.weight.copy_(p.wout.detach())
                s.dclllayer.output_.bias.copy_(p.bout.detach())
        if hasattr(s, "optimizer"):
            pairs = [(s.optimizer, i2h.weight, sl["w"]), (s.optimizer, i2h.bias, sl["b"])]
            if p.wout is not None:
                pairs += [(s.optimizer2, s.dclllayer.output_.weight, sl["wout"]),
                          (s.optimizer2, s.dclllayer.output_.bias, sl["bout"])]
            for opt, prm, slot in pairs:
                if slot.exp_avg is None:
                    opt.state.pop(prm, None)
                    continue
                opt.state[prm] = {"step": torch.tensor(float(slot.step)), "exp_avg": slot.exp_avg.cuda().clone(),
                                  "exp_avg_sq": slot.exp_avg_sq.cuda().clone()}


def rel_err(a, b):
    """max |a-b| relative to the tensor's own scale (element-wise relative error is meaningless at the
    zero crossings of a membrane potential)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    scale = float(b.abs().max())
    return float((a - b).abs().max()) / (scale if scale > 0 else 1.0)


def assert_adam_step_close(w, w_ref, lr, who):
    """Weights after ONE Adam step (betas (0, .95): the step is lr * g / (|g| + eps), i.e. +-lr whatever |g| is) against the
    oracle's.  The split-bf16 gradient is exact to ~1e-5 of the gradient's scale (tools/wgrad_err.py), so an element moves the
    other way only when gW cancels against weight_decay * w to that level -- a handful of the 50 K weights of a layer per
    step, whichever kernel computes the sum, and then by at most 2 lr.  Bounds: mean <= 2e-3 lr, at most 1e-4 of the
    elements beyond 0.5 lr, none beyond 2.1 lr."""
    dw = (w - w_ref).abs()
    frac = float((dw > 0.5 * lr).float().mean())
    assert float(dw.mean()) <= 2e-3 * lr and frac <= 1e-4 and float(dw.max()) <= 2.1 * lr, \
        (who, float(dw.max()) / lr, float(dw.mean()) / lr, frac)
