"""Quantised-weight path: CUDA codes vs the defining restatement in oracle/quant.py (parity unpinned:
the reference has no quantised code, see oracle/quant.py)."""
import numpy as np
import pytest
import torch

from oracle import quant as Q


def test_oracle_quantizer_properties():
    rs = np.random.RandomState(0)
    w = (rs.uniform(-1, 1, size=(8, 3, 7, 7)) * 1e-6).astype(np.float32)
    w[3] = 0
    w[4, 0, 0, 0] = 0.5e-6
    q, s = Q.quantize(w)
    assert q.dtype == np.int8 and np.abs(q).max() == 127 and s[3] == 1.0 and not q[3].any()
    assert np.abs(Q.dequantize(q, s) - w).max() <= 0.5 * s.max() * (1 + 1e-6)
    assert np.array_equal(Q.quantize(Q.dequantize(q, s))[0], q)            # idempotent
    assert np.array_equal(Q.quantize(-w)[0], -q)                           # symmetric


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(32, 32, 7, 7), (32, 1, 7, 7), (64, 64, 1, 3), (5, 3)])
def test_cuda_codes_bit_exact(shape):
    from snn_modulation_classification_b200 import quant
    g = torch.Generator().manual_seed(3)
    w = (torch.rand(shape, generator=g) * 2 - 1) * 2e-6
    w[1] = 0
    if w[0].numel() >= 8:
        w.view(shape[0], -1)[2, :4] = torch.tensor([0.5, 1.5, 2.5, -0.5]) * (w[2].abs().max() / 127)   # ties -> even
    codes, scales = quant.quantize(w.cuda())
    q, s = Q.quantize(w.numpy())
    assert np.array_equal(codes.cpu().numpy(), q)
    assert np.array_equal(scales.cpu().numpy(), s)
    assert np.array_equal(quant.dequantize(codes, scales).cpu().numpy(), Q.dequantize(q, s))


@pytest.mark.gpu
def test_quantized_forward_matches_oracle_with_dequantized_weights():
    from snn_modulation_classification_b200 import quant
    from oracle import dcll_oracle as O
    from util_build import build_pair, rel_err
    net, onet = build_pair("radio_ml_conv", (1, 16, 16), 4, 24, train=False)
    quant.enable_quantized_weights(net)
    for p in onet.params:
        p.weight = torch.from_numpy(Q.fake_quantize(p.weight.numpy()))
    x = (torch.rand(5, 4, 1, 16, 16) < 0.1).float()
    net.reset()
    onet.reset()
    for t in range(5):
        net.test(x[t].cuda())
        onet.test(x[t])
    for i, s in enumerate(net.dcll_slices):
        assert torch.equal(s.dclllayer.i2h.state.eps1.cpu(), onet.states[i].eps1)
        assert rel_err(s.dclllayer._ctx[1]["pvmem"], onet.last[i].pvmem) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_quantized_training_window_equals_per_step_and_uses_requantized_weights(precision):
    """Quantised mode inside the C window driver: after every Adam step the library re-derives the int8 image of
    the updated float32 master weights (straight-through training, oracle/quant.py)."""
    from snn_modulation_classification_b200 import quant
    from snn_modulation_classification_b200.data.utils import iq2spiketrain
    from oracle import dcll_oracle as O
    from util_build import build_pair
    B, K, T, W, burnin = 8, 24, 8, 16, 3
    nets = []
    for _ in range(2):
        n, _ = build_pair("radio_ml_conv", (1, W, W), B, K, arp=0.0, burnin=burnin)
        n.set_precision(precision)
        nets.append(quant.enable_quantized_weights(n))
    g = torch.Generator().manual_seed(4)
    xs = (torch.randn(B, 2, 1, 1024, generator=g) * 0.4).float()
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K).cuda()
    np.random.seed(1)
    cells, tgt = iq2spiketrain(xs, y, out_w=W, out_h=W, max_duration=T, as_cells=True)
    for n in nets:
        n.reset()
    for t in range(T):
        nets[0].learn(cells[t], tgt[t])
    nets[1].learn_window(cells, y)
    for a, b in zip(nets[0].dcll_slices, nets[1].dcll_slices):
        ia, ib = a.dclllayer.i2h, b.dclllayer.i2h
        assert torch.equal(ia.weight, ib.weight) and torch.equal(ia.state.eps1, ib.state.eps1)
        # kernel-side weights = transpose of the fake-quantised master weights, bit for bit
        cinkk = ib.weight[0].numel()
        cout_pad = (ib.out_channels + 31) // 32 * 32
        wt = ib._wt[:cinkk * cout_pad].view(cinkk, cout_pad)[:, :ib.out_channels].t().cpu().numpy()
        want = Q.fake_quantize(ib.weight.detach().cpu().numpy()).reshape(ib.out_channels, cinkk)
        assert np.array_equal(wt, want)
        assert not np.array_equal(want, ib.weight.detach().cpu().numpy().reshape(ib.out_channels, cinkk))
