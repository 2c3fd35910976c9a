"""Debug/parity probe of the bf16x3 tensor-core conv against the oracle (teacher-forced, per layer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import dcll_oracle as O
from util_build import build_pair, force_state, rel_err

def run(im, B, arp, steps=4, mode='bf16x3'):
    K = 24
    net, onet = build_pair("radio_ml_conv", (1,)+im, B, K, arp=arp, train=False)
    net.set_precision(mode)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(steps, B, 1, *im, generator=g) < 0.1).float()
    net.reset(); onet.reset()
    for t in range(steps):
        force_state(net, onet)
        onet.test(x[t])
        for i, s in enumerate(net.dcll_slices):
            inp = x[t].cuda() if i == 0 else onet.last[i-1].output.cuda()
            out, pvo, pv, pvmem = s.forward(inp, ignore_burnin=True)
            fo = onet.last[i]
            st = s.dclllayer.i2h.state
            eq = torch.equal(st.eps1.cpu(), fo.state.eps1) and torch.equal(st.eps0.cpu(), fo.state.eps0)
            flips = float((s.dclllayer._ctx[1]['spikes'].cpu() != fo.spikes).float().mean())
            print("im=%s B=%d arp=%g t=%d L%d tc=%s traces_exact=%s pvmem_rel=%.2e flips=%.2e pvo_rel=%.2e" % (
                im, B, arp, t, i, s.dclllayer.i2h.tensor_core_ok(), eq, rel_err(pvmem, fo.pvmem), flips, rel_err(pvo, fo.pvoutput)))
    torch.cuda.synchronize()

MODE = os.environ.get("DCLL_PRECISION", "bf16x3")
if __name__ == "__main__":
    run((16, 16), 4, 0.0, mode=MODE)
    run((40, 24), 3, 1.0, mode=MODE)
    run((128, 128), 2, 0.0, steps=2, mode=MODE)


def check_wgrad(im, B, arp=0.0, mode='bf16x3'):
    from snn_modulation_classification_b200 import networks as N
    from util_build import make_args, state_dict_from_params
    K = 24
    specs = O.make_specs(O.BUILTIN_SPECS["radio_ml_conv"], (1,) + im, K, wrp=arp)
    params = O.random_params(specs, seed=2)
    sd = state_dict_from_params(params)
    net = N.ConvNetwork(make_args(arp), (1,) + im, B, N.load_network_spec("radio_ml_conv"), K, act=torch.nn.Sigmoid(),
                        loss=torch.nn.SmoothL1Loss, opt=torch.optim.SGD, opt_param={}, learning_rates=[0.0], burnin=0)
    net.load_state_dict(sd); net = net.to("cuda"); net.reset(True); net.load_state_dict(sd)
    net.set_precision(mode)
    onet = O.OracleNet(specs, params, B, burnin=0)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(3, B, 1, *im, generator=g) < 0.1).float()
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K)
    for t in range(3):
        force_state(net, onet)
        fos, grads = [], []
        inp = x[t]
        for i, sp in enumerate(specs):
            fo = O.conv_step_fwd(sp, params[i], onet.states[i], inp)
            grads.append(O.conv_local_grads(sp, params[i], fo, y)); fos.append(fo)
            onet.states[i] = fo.state; inp = fo.spikes
        for i, s in enumerate(net.dcll_slices):
            lay = s.dclllayer
            s.train_dcll(x[t].cuda() if i == 0 else fos[i - 1].spikes.cuda(), y.cuda(), regularize=False)
            gr = grads[i]
            print("wgrad im=%s B=%d t=%d L%d tc=%s gW_rel=%.2e gb_rel=%.2e" % (im, B, t, i, lay.i2h.tensor_core_ok(),
                  rel_err(lay.i2h.weight.grad, gr.gW), rel_err(lay.i2h.bias.grad, gr.gb)))


if __name__ == "__main__" and "--wgrad" in sys.argv:
    check_wgrad((16, 16), 8, mode=MODE)
    check_wgrad((40, 24), 3, 1.0, mode=MODE)
    check_wgrad((128, 128), 2, mode=MODE)
