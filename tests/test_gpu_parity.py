"""CUDA path vs oracle and vs the golden fixtures (needs a B200: pytest -m gpu).

Every call goes through the host mirror (snn_modulation_classification_b200) and therefore through the C ABI
of libdcll_b200.so.  Tolerances (BASELINE.json north_star, FP32-exact mode):
  * encoder cells, quantiser codes, vote                       bit-exact
  * synaptic traces eps0/eps1                                  bit-exact (one rounding per reference operation)
  * membrane / pv / read-outs                                  <= 1e-5 of the tensor's scale
  * spikes                                                     flip rate <= 1e-4 (teacher-forced, per layer)
  * weights after the per-timestep Adam step                   <= 5e-2 lr max, <= 2e-4 lr mean (see DESIGN.md:
                                                                Adam normalises by sqrt(v), so gW rounding shows
                                                                up in units of lr, not of |w|)
"""
import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR, NetFixture, sub
from oracle import dcll_oracle as O
from util_build import build_pair, force_state, rel_err

pytestmark = pytest.mark.gpu

MEM_TOL = 1e-5
FLIP_TOL = 1e-4


# ------------------------------------------------------------------------------------------------
# encoder
# ------------------------------------------------------------------------------------------------
def _encode_gpu(x, W, H, bounds, t_start, T, gamma, seed_for_start=None):
    from snn_modulation_classification_b200 import _lib
    xc = torch.as_tensor(x).reshape(x.shape[0], 2, x.shape[-1]).float().cuda().contiguous()
    cells = torch.empty((T, xc.shape[0], 2), dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib.dcll_iq_encode(_lib.ptr(xc), xc.shape[0], xc.shape[-1], bounds[0], bounds[1], bounds[2], bounds[3],
                                       W, H, t_start, T, int(gamma), _lib.ptr(cells), _lib.current_stream()))
    return cells


def test_encoder_golden_bit_exact():
    z = np.load(GOLDEN_DIR + "/encoder.npz")
    x = z["x"]
    n = 0
    for k in z.files:
        if not k.startswith("cells__"):
            continue
        key = k[len("cells__"):]
        parts = key.split("_")
        W, H, T, gamma = int(parts[0][1:]), int(parts[1][1:]), int(parts[2][1:]), bool(int(parts[3][1:]))
        bounds = [float(v) for v in key.split("_b")[1].split("_")]
        cells = _encode_gpu(x, W, H, bounds, int(z["tstart__" + key]), T, gamma)
        assert np.array_equal(cells.cpu().numpy(), z[k]), key
        n += 1
    assert n == 4


@pytest.mark.parametrize("W,B,T", [(16, 64, 1024), (128, 64, 1024), (28, 37, 300), (16, 512, 1024)])
def test_encoder_vs_oracle_bit_exact(W, B, T):
    g = torch.Generator().manual_seed(W + B)
    x = (torch.randn(B, 2, 1, 1024, generator=g) * 0.4).float()
    x[0, :, 0, :6] = torch.tensor([[-1.0, 1.0, 0.0, 3.0, -3.0, 1e-8], [1.0, -1.0, -0.0, 0.5, -0.5, -1e-8]])
    t_start = (1024 - T) // 2
    want = O.encode_cells(x.numpy(), W, W, t_start=t_start, max_duration=T)
    got = _encode_gpu(x.numpy(), W, W, (-1, 1, -1, 1), t_start, T, True).cpu().numpy()
    assert np.array_equal(got, want)
    assert got.min() >= 0 and got.max() <= W - 1


def test_iq2spiketrain_api_one_spike_per_frame():
    from snn_modulation_classification_b200.data.utils import iq2spiketrain, to_one_hot
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(64, 2, 1, 1024, generator=g) * 0.4).float()
    y = to_one_hot(torch.arange(64) % 24, 24)
    np.random.seed(5)
    frames, tgt = iq2spiketrain(x, y, out_w=16, out_h=16, max_duration=200)
    np.random.seed(5)
    t_start = np.random.randint(0, 1024 - 200 + 1)
    assert frames.shape == (200, 64, 1, 16, 16) and frames.dtype == torch.float32 and frames.is_cuda
    assert float(frames.sum()) == 200 * 64                                    # exactly one spike per (t, b)
    assert torch.equal(frames.flatten(2).sum(-1), torch.ones(200, 64, device="cuda"))
    want = O.cells_to_frames(O.encode_cells(x.numpy(), 16, 16, t_start=t_start, max_duration=200), 16, 16)
    assert np.array_equal(frames.cpu().numpy(), want)
    assert tgt.shape == (200, 64, 24) and torch.equal(tgt[7].cpu(), y.cpu())
    with pytest.raises(AssertionError):
        iq2spiketrain(x, y, max_duration=2000)                                # data/utils.py:56
    np.random.seed(5)
    cells, _ = iq2spiketrain(x, y, out_w=16, out_h=16, max_duration=200, as_cells=True)
    assert torch.equal(cells.dense(), frames)


# ------------------------------------------------------------------------------------------------
# per-layer, teacher-forced: forward, local gradient, Adam step
# ------------------------------------------------------------------------------------------------
CONFIGS = [
    # name, spec, im_dims, B, K, arp
    ("radio16", "radio_ml_conv", (1, 16, 16), 8, 24, 0.0),
    ("radio16_arp", "radio_ml_conv", (1, 16, 16), 8, 24, 1.0),
    ("radio128", "radio_ml_conv", (1, 128, 128), 2, 24, 0.0),
    ("radio40x24_arp", "radio_ml_conv", (1, 40, 24), 3, 24, 1.0),       # ragged tiles
    ("mnist", "mnist_conv", (1, 28, 28), 5, 10, 0.0),                   # pooling 2, odd sizes
    # F = 32*56*56 = 100 352: the large-F output layer (packed g_u sweep + wout_grad_adam2_kernel); B = 21 = two full
    # groups of 8 prefetched rows + a remainder of 5
    ("radio56_b21", "radio_ml_conv", (1, 56, 56), 21, 24, 0.0),
    ("radioref", "radio_ml_conv_ref", (1, 4, 128), 2, 24, 0.0),         # (1,3) kernels, (1,2) pooling, 64 ch
]


def _inputs(im_dims, B, K, steps, seed=5, rate=0.1):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(steps, B, *im_dims, generator=g) < rate).float()
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K)
    return x, y


@pytest.mark.parametrize("name,spec,im_dims,B,K,arp", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_layer_steps_teacher_forced(name, spec, im_dims, B, K, arp):
    burnin, steps, lr = 3, 7, 1e-6
    net, onet = build_pair(spec, im_dims, B, K, arp=arp, burnin=burnin, lr=lr)
    x, y = _inputs(im_dims, B, K, steps)
    yc = y.cuda()
    net.reset()
    onet.reset()
    flips = np.zeros(len(net.dcll_slices))
    total = np.zeros(len(net.dcll_slices))
    for t in range(steps):
        force_state(net, onet)                           # identical state and weights going into the step
        onet.learn(x[t], y)
        spikes_in = x[t].cuda()
        for i, s in enumerate(net.dcll_slices):
            fo = onet.last[i]
            inp = spikes_in if i == 0 else onet.last[i - 1].output.cuda()   # teacher-forced layer input
            out, pvo, pv, pvmem, _loss = s.train_dcll(inp, yc, regularize=False)
            st = s.dclllayer.i2h.state
            assert torch.equal(st.eps0.cpu(), fo.state.eps0), (name, t, i, "eps0")
            assert torch.equal(st.eps1.cpu(), fo.state.eps1), (name, t, i, "eps1")
            assert rel_err(pvmem, fo.pvmem) <= MEM_TOL, (name, t, i, "pvmem", rel_err(pvmem, fo.pvmem))
            spk = out if not s.dclllayer.output_layer else None
            if spk is not None:
                flips[i] += float((spk.cpu() != fo.spikes).sum())
                total[i] += spk.numel()
            # pv differs from the oracle only where the membrane does
            assert float((pv.cpu() - fo.pv).abs().max()) <= 2e-6 + 0.25 * MEM_TOL * float(fo.pvmem.abs().max())
            assert rel_err(pvo, fo.pvoutput) <= MEM_TOL, (name, t, i, "pvoutput")
            if s.dclllayer.output_layer:
                assert rel_err(out, fo.output) <= MEM_TOL
            if arp > 0:
                # refractory state differs only where a spike flipped
                bad = float((st.arp.cpu() - fo.state.arp).abs().gt(1e-5).float().mean())
                assert bad <= FLIP_TOL
            if onet.iters[i] >= burnin:
                p = onet.params[i]
                dw = (s.dclllayer.i2h.weight.detach().cpu() - p.weight).abs()
                assert float(dw.max()) <= 5e-2 * lr and float(dw.mean()) <= 2e-4 * lr, (name, t, i, float(dw.max()) / lr)
                db = (s.dclllayer.i2h.bias.detach().cpu() - p.bias).abs()
                assert float(db.max()) <= 5e-2 * lr
                if s.dclllayer.output_layer:
                    dwo = (s.dclllayer.output_.weight.detach().cpu() - p.wout).abs()
                    assert float(dwo.max()) <= 5e-2 * 1e-4 and float(dwo.mean()) <= 2e-4 * 1e-4
                    dbo = (s.dclllayer.output_.bias.detach().cpu() - p.bout).abs()
                    assert float(dbo.max()) <= 5e-2 * 1e-4
    for i in range(len(flips)):
        if total[i]:
            assert flips[i] / total[i] <= FLIP_TOL, (name, i, flips[i] / total[i])
    # per-step API bookkeeping: clout rows counted from burn-in on (ref :724)
    for s in net.dcll_slices:
        assert len(s.clout) == steps - burnin + 1


@pytest.mark.parametrize("name,spec,im_dims,B,K,arp", CONFIGS[:6], ids=[c[0] for c in CONFIGS[:6]])
def test_raw_gradients_match_closed_form(name, spec, im_dims, B, K, arp):
    """g_u / gW / gb / gWout straight from the kernels (apply_update = 0 path, optimizer = SGD lr 0)."""
    from snn_modulation_classification_b200 import networks as N
    from util_build import make_args, state_dict_from_params
    specs = O.make_specs(O.BUILTIN_SPECS[spec], im_dims, K, wrp=arp)
    params = O.random_params(specs, seed=2)
    sd = state_dict_from_params(params)
    net = N.ConvNetwork(make_args(arp), im_dims, B, N.load_network_spec(spec), K, act=torch.nn.Sigmoid(),
                        loss=torch.nn.SmoothL1Loss, opt=torch.optim.SGD, opt_param={}, learning_rates=[0.0], burnin=0)
    net.load_state_dict(sd)
    net = net.to("cuda")
    net.reset(True)
    net.load_state_dict(sd)
    onet = O.OracleNet(specs, params, B, burnin=0)
    x, y = _inputs(im_dims, B, K, 3)
    yc = y.cuda()
    for t in range(3):
        force_state(net, onet)
        fos, grads = [], []
        inp = x[t]
        for i, sp in enumerate(specs):
            fo = O.conv_step_fwd(sp, params[i], onet.states[i], inp)
            grads.append(O.conv_local_grads(sp, params[i], fo, y))
            fos.append(fo)
            onet.states[i] = fo.state
            inp = fo.spikes
        for i, s in enumerate(net.dcll_slices):
            lay = s.dclllayer
            s.train_dcll(x[t].cuda() if i == 0 else fos[i - 1].spikes.cuda(), yc, regularize=False)
            g = grads[i]
            scale = float(g.gW.abs().max())
            assert float((lay.i2h.weight.grad.cpu() - g.gW).abs().max()) <= 2e-5 * scale, (name, t, i)
            assert float((lay.i2h.bias.grad.cpu() - g.gb).abs().max()) <= 2e-5 * float(g.gb.abs().max())
            # g_u on the pooled grid equals the oracle's scattered gradient gathered at the argmax
            gu = lay._g_u.cpu()
            want = g.g_u.flatten(2).gather(2, fos[i].pool_idx.flatten(2)).reshape(gu.shape)
            assert float((gu - want).abs().max()) <= 2e-5 * float(want.abs().max())
            if lay.output_layer:
                assert rel_err(lay.output_.weight.grad, g.gWout) <= 1e-5
                assert rel_err(lay.output_.bias.grad, g.gbout) <= 1e-5


# ------------------------------------------------------------------------------------------------
# free-running network: inference (flip rate, clout, vote) and window API == per-step API
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("arp", [0.0, 1.0])
def test_inference_free_running_and_window(arp):
    from snn_modulation_classification_b200.data.utils import iq2spiketrain
    B, K, T, W = 16, 24, 60, 16
    net, onet = build_pair("radio_ml_conv", (1, W, W), B, K, arp=arp, train=False)
    net2, _ = build_pair("radio_ml_conv", (1, W, W), B, K, arp=arp, train=False)
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(B, 2, 1, 1024, generator=g) * 0.4).float()
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K)
    np.random.seed(1)
    cells, tgt = iq2spiketrain(x, y, out_w=W, out_h=W, max_duration=T, as_cells=True)
    frames = cells.dense()
    net.reset()
    net.eval()
    net2.reset()
    onet.reset()
    fl = np.zeros(3)
    for t in range(T):
        net.test(frames[t])
        onet.test(frames[t].cpu())
        for i, s in enumerate(net.dcll_slices):
            if i < 2:
                fl[i] += float((s.dclllayer._ctx[1]["spikes"].cpu() != onet.last[i].spikes).float().mean())
    assert (fl / T).max() <= FLIP_TOL, fl / T
    # window API on the compact cell input == per-step API on dense frames, bit for bit
    clout = net2.test_window(cells)
    for i, (a, b) in enumerate(zip(net.dcll_slices, net2.dcll_slices)):
        assert np.array_equal(np.array(a.clout), np.array(b.clout))
        assert torch.equal(a.dclllayer.i2h.state.eps1, b.dclllayer.i2h.state.eps1)
        assert np.array_equal(clout[:, i, :].cpu().numpy(), np.array(a.clout))
        agree = (np.array(a.clout) == np.array(onet.clout[i])).mean()
        assert agree >= 1 - 5e-3, (i, agree)
    assert net.accuracy(tgt) == net2.accuracy(tgt)
    labels = tgt.cpu().numpy()
    cm = net.confusion_matrix(tgt)
    assert cm.sum() == B and cm.shape == (K, K)
    # device vote == Counter vote on the same predictions
    pred_host, _ = O.predictions_by_vote(list(np.array(net.dcll_slices[-1].clout)), labels)
    assert np.array_equal(net.dcll_slices[-1].clout.vote(K), pred_host)


def test_training_window_equals_per_step():
    """dcll_net_window (C loop) and the per-timestep Python API run the same kernels in the same order."""
    from snn_modulation_classification_b200.data.utils import iq2spiketrain
    B, K, T, W, burnin = 8, 24, 12, 16, 4
    a, _ = build_pair("radio_ml_conv", (1, W, W), B, K, arp=1.0, burnin=burnin)
    b, _ = build_pair("radio_ml_conv", (1, W, W), B, K, arp=1.0, burnin=burnin)
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(B, 2, 1, 1024, generator=g) * 0.4).float()
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K).cuda()
    np.random.seed(1)
    cells, tgt = iq2spiketrain(x, y, out_w=W, out_h=W, max_duration=T, as_cells=True)
    a.reset()
    b.reset()
    for t in range(T):
        a.learn(cells[t], tgt[t])
    b.learn_window(cells, y)
    for sa, sb in zip(a.dcll_slices, b.dcll_slices):
        assert torch.equal(sa.dclllayer.i2h.weight, sb.dclllayer.i2h.weight)
        assert torch.equal(sa.dclllayer.i2h.bias, sb.dclllayer.i2h.bias)
        assert torch.equal(sa.dclllayer.i2h.state.eps1, sb.dclllayer.i2h.state.eps1)
        assert torch.equal(sa.dclllayer.i2h.state.arp, sb.dclllayer.i2h.state.arp)
        assert np.array_equal(np.array(sa.clout), np.array(sb.clout)) and len(sa.clout) == T - burnin + 1
        assert sa.iter == sb.iter == T
        wa, wb = sa.optimizer.state[sa.dclllayer.i2h.weight], sb.optimizer.state[sb.dclllayer.i2h.weight]
        assert float(wa["step"]) == float(wb["step"]) == T - burnin + 1
        assert torch.equal(wa["exp_avg_sq"], wb["exp_avg_sq"])
    assert torch.equal(a.dcll_slices[-1].dclllayer.output_.weight, b.dcll_slices[-1].dclllayer.output_.weight)
    # a second window continues from the carried state (no state reset between batches)
    a.reset()
    b.reset()
    for t in range(T):
        a.learn(cells[t], tgt[t])
    b.learn_window(cells, y)
    assert torch.equal(a.dcll_slices[1].dclllayer.i2h.weight, b.dcll_slices[1].dclllayer.i2h.weight)


# ------------------------------------------------------------------------------------------------
# golden fixtures (frozen reference outputs)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["radio8_train", "radio8_arp_train", "radio8_arp_infer", "mnist_train", "radioref_train"])
def test_network_golden(name):
    fx = NetFixture(name)
    net, _ = build_pair(fx.spec_name, fx.im_dims, fx.B, fx.K, arp=fx.arp, burnin=fx.burnin, train=fx.train,
                        state_dict=fx.state_dict)
    net.reset()
    lr = 1e-6
    first_train_t = fx.burnin - 1
    yc = fx.y.cuda()
    for t in range(fx.steps):
        if fx.train:
            net.learn(fx.x[t].cuda(), yc)
        else:
            net.test(fx.x[t].cuda())
        horizon = t - first_train_t if fx.train else 0
        for i, s in enumerate(net.dcll_slices):
            lay = s.dclllayer
            if horizon <= 0:
                # until the first weight update the run is free of training chaos: traces are exact up to
                # spike flips upstream (none expected at these sizes)
                np.testing.assert_allclose(sub(lay.i2h.state.eps1), fx.get(t, i, "eps1"), rtol=1e-6, atol=0)
            if fx.train and 0 <= horizon <= 1:
                d = np.abs(sub(lay.i2h.weight) - fx.get(t, i, "w"))
                assert d.max() <= 5e-2 * (2 ** horizon) * lr, (name, t, i, d.max() / lr)
                db = np.abs(lay.i2h.bias.detach().cpu().numpy() - fx.get(t, i, "b"))
                assert db.max() <= 5e-2 * (2 ** horizon) * lr
                if lay.output_layer:
                    assert np.abs(sub(lay.output_.weight) - fx.get(t, i, "wout")).max() <= 5e-2 * 1e-4
    if not fx.train:
        for i, s in enumerate(net.dcll_slices):
            assert np.array_equal(np.array(s.clout), fx.clout(i))
        labels = torch.stack([fx.y] * fx.steps)
        assert net.accuracy(labels) == list(fx.z["acc"])
        assert np.array_equal(net.confusion_matrix(labels), fx.z["confusion"])


# ------------------------------------------------------------------------------------------------
# API behaviour the reference defines
# ------------------------------------------------------------------------------------------------
def test_batch_size_change_reallocates_state(caplog):
    net, _ = build_pair("radio_ml_conv", (1, 16, 16), 4, 24, train=False)
    x = (torch.rand(6, 1, 16, 16) < 0.1).float().cuda()
    net.test(x)                                             # ref :410-413: warning, not an error
    assert net.dcll_slices[0].dclllayer.i2h.state.eps0.shape[0] == 6
    assert net.dcll_slices[2].dclllayer.i2h.state.eps1.shape[0] == 6


def test_state_not_reset_between_batches():
    net, _ = build_pair("radio_ml_conv", (1, 16, 16), 4, 24, train=False)
    x = (torch.rand(4, 1, 16, 16) < 0.1).float().cuda()
    net.test(x)
    e = net.dcll_slices[0].dclllayer.i2h.state.eps1.clone()
    net.reset()                                             # clears clout/iter only
    assert torch.equal(net.dcll_slices[0].dclllayer.i2h.state.eps1, e) and float(e.abs().sum()) > 0
    assert len(net.dcll_slices[0].clout) == 0 and net.dcll_slices[0].iter == 0
    net.reset(True)
    assert float(net.dcll_slices[0].dclllayer.i2h.state.eps1.abs().sum()) == 0


def test_unsupported_shapes_fail_loudly():
    from snn_modulation_classification_b200.dcll.pytorch_libdcll import Conv2dDCLLlayer
    with pytest.raises(NotImplementedError):
        Conv2dDCLLlayer(1, 8, kernel_size=7, im_dims=(16, 16), pooling=3)
    lay = Conv2dDCLLlayer(1, 8, kernel_size=9, padding=4, im_dims=(16, 16), pooling=1).to("cuda").init_hiddens(2)
    with pytest.raises(NotImplementedError):
        lay.forward(torch.zeros(2, 1, 16, 16, device="cuda"))
    with pytest.raises(ValueError):
        Conv2dDCLLlayer(3, 8, kernel_size=7, im_dims=(16, 16)).i2h.__class__(3, 8, 7, groups=2)


# ------------------------------------------------------------------------------------------------
# dense layers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wrp,random_tau", [(0.0, False), (0.0, True), (2.0, True)])
def test_dense_layer_vs_oracle(wrp, random_tau):
    from snn_modulation_classification_b200.dcll import pytorch_libdcll as L
    torch.manual_seed(0)
    np.random.seed(0)
    B, In, Out, K, lr = 9, 70, 45, 10, 1e-3
    lay = L.DenseDCLLlayer(In, Out, target_size=K, wrp=wrp, random_tau=random_tau).to("cuda").init_hiddens(B)
    sl = L.DCLLClassification(dclllayer=lay, batch_size=B, loss=torch.nn.SmoothL1Loss, optimizer=torch.optim.Adam,
                              kwargs_optimizer={"lr": lr, "betas": [0.0, 0.95], "weight_decay": 10.0}, burnin=2)
    m = lay.i2h
    c = lambda t: t.detach().cpu().clone()
    p = O.ConvParams(c(m.weight), c(m.bias), c(m.alpha), c(m.alphas), c(m.tau_m__dt), c(m.tau_s__dt), c(lay.i2o.weight),
                     c(lay.i2o.bias))
    st = O.ConvState(torch.zeros(B, In), torch.zeros(B, In), torch.zeros(B, Out) if wrp > 0 else None)
    slots = dict(w=O.AdamSlot(), b=O.AdamSlot())
    g = torch.Generator().manual_seed(1)
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K)
    for t in range(6):
        x = (torch.rand(B, In, generator=g) < 0.3).float()
        # teacher forcing of the weights (training is chaotic)
        with torch.no_grad():
            m.weight.copy_(p.weight)
            m.bias.copy_(p.bias)
        out, pvo, pv, vmem, _ = sl.train_dcll(x.cuda(), y.cuda(), regularize=False)
        fo = O.dense_step_fwd(p, st, x, wrp=wrp)
        st = fo.state
        assert torch.equal(m.state.eps0.cpu(), st.eps0) and torch.equal(m.state.eps1.cpu(), st.eps1)
        assert rel_err(vmem, fo.vmem) <= MEM_TOL and rel_err(pvo, fo.pvoutput) <= MEM_TOL
        assert float((out.cpu() != fo.spikes).float().mean()) <= 1e-3
        if wrp > 0:
            assert float((m.state.arp.cpu() - st.arp).abs().gt(1e-5).float().mean()) <= 1e-3
        if t + 1 >= 2:
            g_o, g_u, gW, gb = O.dense_local_grads(p, fo, y)
            assert rel_err(m.weight.grad, gW) <= 2e-5 and rel_err(m.bias.grad, gb) <= 2e-5
            kw = dict(lr=lr, beta1=0.0, beta2=0.95, weight_decay=10.0)
            O.adam_update(p.weight, gW, slots["w"], **kw)
            O.adam_update(p.bias, gb, slots["b"], **kw)
            # moments follow the oracle's (same gradients up to rounding)
            stt = sl.optimizer.state[m.weight]
            assert float(stt["step"]) == slots["w"].step
            assert float((m.weight.detach().cpu() - p.weight).abs().max()) <= 5e-2 * lr
            stt["exp_avg_sq"].copy_(slots["w"].exp_avg_sq)
            sl.optimizer.state[m.bias]["exp_avg_sq"].copy_(slots["b"].exp_avg_sq)
    assert len(sl.clout) == 5
    # stand-alone i2h call (ref :131-148)
    o2, pv2, vm2 = m.forward(torch.zeros(B, In, device="cuda"))
    assert o2.shape == (B, Out) and vm2.shape == (B, Out)


def test_window_activity_histogram_matches_numpy():
    """collect_stats in the C window driver (ref :658-661): 19-bin pv histogram every 20 iterations."""
    net, onet = build_pair("radio_ml_conv", (1, 16, 16), 4, 24, arp=0.0, train=False)
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(45, 4, 1, 16, 16, generator=g) < 0.1).float()
    net.reset()
    onet.reset()
    net.test_window(x.cuda())
    bins = np.linspace(0, 1, 20)
    want = [[] for _ in range(3)]
    for t in range(45):
        onet.test(x[t])
        if (t + 1) % 20 == 0:
            for i in range(3):
                want[i].append(np.histogram(onet.last[i].pv.numpy(), bins=bins)[0])
    for i, s in enumerate(net.dcll_slices):
        assert len(s.activity_hist) == 2
        got = torch.stack(s.activity_hist).cpu().numpy()
        # pv differs from the oracle by ~1e-7 at bin edges: compare with a tolerance of a few elements per bin
        assert np.abs(got - np.array(want[i])).max() <= 3, (i, got, want[i])
        assert got.sum(axis=1).tolist() == [4 * 32 * 256] * 2

    class W:
        def __init__(self):
            self.scalars = {}

        def add_histogram(self, *a, **k):
            pass

        def add_scalar(self, name, v, epoch):
            self.scalars[name] = v

    w = W()
    net.accuracy(torch.zeros(45, 4, 24))
    net.write_stats(w, epoch=0)
    assert "conv0/low_pv/test" in w.scalars and "conv2/acc/test" in w.scalars


# ------------------------------------------------------------------------------------------------
# image2spiketrain on the device (data/utils.py:15-40)
# ------------------------------------------------------------------------------------------------
def test_image2spiketrain_device_matches_numpy_stream():
    from snn_modulation_classification_b200.data.utils import image2spiketrain
    g = torch.Generator().manual_seed(3)
    x = torch.rand(9, 1, 28, 28, generator=g)
    y = O.to_one_hot(torch.randint(0, 10, (9,), generator=g), 10)
    for kw in (dict(gain=100, max_duration=60), dict(gain=50, min_duration=7, max_duration=41)):
        np.random.seed(5)
        want, want_t = O.image2spiketrain(x, y.numpy(), (1, 28, 28), **kw)
        np.random.seed(5)
        got, got_t = image2spiketrain(x, y, (1, 28, 28), **kw)
        assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == want.shape
        assert np.array_equal(got.cpu().numpy(), want.astype(np.float32))              # bit-exact for the same draws
        assert np.array_equal(got_t.cpu().numpy(), want_t)
    # device generator: same distribution (rate = gain * pixel / 1000 per timestep), silent after T_i, reproducible per seed
    np.random.seed(5)
    a, _ = image2spiketrain(x, y, (1, 28, 28), gain=100, min_duration=30, max_duration=400, device_rng=7)
    np.random.seed(5)
    b, _ = image2spiketrain(x, y, (1, 28, 28), gain=100, min_duration=30, max_duration=400, device_rng=7)
    np.random.seed(5)
    T = np.random.randint(30, 400, 9)
    assert torch.equal(a, b)
    for i in range(9):
        assert float(a[T[i]:, i].sum()) == 0
        rate = a[:T[i], i].mean(0).cpu().reshape(-1)
        want_rate = (100 * x[i].reshape(-1) / 1000).clamp(0, 1)
        # per-pixel sampling error of a Bernoulli rate over T_i timesteps ~ sqrt(p(1-p)/T_i); the image mean is 784x tighter
        assert float((rate - want_rate).abs().mean()) < 1.5 * float(np.sqrt(0.05 / T[i]))
        assert abs(float(rate.mean() - want_rate.mean())) < 5 * float(np.sqrt(0.05 / (T[i] * 784)))
