"""Single Conv2dDCLLlayer on odd shapes vs the oracle (needs a B200): every kernel-size instantiation, channel counts
that are not multiples of 8/32, images smaller than a tile, ragged tiles, pooling with odd sizes, batch 1."""
import math

import numpy as np
import pytest
import torch

from oracle import dcll_oracle as O
from util_build import rel_err

pytestmark = pytest.mark.gpu

SHAPES = [
    # cin, cout, kernel, pad, pool, (H, W), B, K, wrp, output_layer
    (1, 8, 5, 2, None, (28, 28), 3, 10, 0.0, False),          # Conv2dDCLLlayer defaults (kernel 5, pad 2, no pooling)
    (3, 5, 5, 2, 2, (17, 23), 2, 7, 0.0, True),               # Cout not a multiple of 8, odd sizes, pool 2
    (6, 40, 3, 1, 2, (9, 33), 1, 10, 1.5, False),             # Cout > 32 (two channel chunks), batch 1, refractory
    (2, 16, 3, 0, 1, (5, 6), 4, 3, 0.0, True),                # image smaller than a tile, no padding
    (10, 24, 7, 3, 2, (20, 12), 2, 10, 0.0, False),           # 7x7 with pooling
    (33, 9, (1, 3), (0, 1), (1, 2), (3, 50), 2, 5, 0.0, True),  # 1x3 kernels, (1,2) pooling, Cin > chunk
    (32, 32, 7, 3, 1, (7, 9), 1, 24, 1.0, True),              # tensor-core shape on an image smaller than a tile
]


def _layer_pair(cin, cout, kernel, pad, pool, hw, B, K, wrp, output_layer, seed=0):
    from snn_modulation_classification_b200.dcll import pytorch_libdcll as L
    torch.manual_seed(seed)
    np.random.seed(seed)
    lay = L.Conv2dDCLLlayer(cin, cout, kernel_size=kernel, im_dims=hw, target_size=K, pooling=pool, padding=pad,
                            alpha=.92, alphas=.85, alpharp=.65, wrp=wrp, random_tau=True, output_layer=output_layer)
    lay = lay.to("cuda").init_hiddens(B)
    return lay


def _oracle_of(lay, cin, cout, kernel, pad, pool, hw, B, K, wrp, output_layer):
    """Oracle spec/params captured from the live module (after every init_hiddens: the refractory core re-draws its
    time constants on each init_state, reference quirk ref:479-481)."""
    spec = O.ConvSpec(cin, cout, O._pair(kernel), O._pair(pad), O._pair(pool if pool is not None else 1), tuple(hw), K,
                      0.65, wrp, output_layer)
    m = lay.i2h
    c = lambda t: t.detach().cpu().clone()
    p = O.ConvParams(c(m.weight), c(m.bias), c(m.alpha), c(m.alphas), c(m.tau_m__dt), c(m.tau_s__dt), c(lay.i2o.weight),
                     c(lay.i2o.bias), c(lay.output_.weight) if output_layer else None,
                     c(lay.output_.bias) if output_layer else None)
    return spec, p


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("shape", SHAPES, ids=[str(i) for i in range(len(SHAPES))])
def test_single_layer_forward_and_gradients(shape, precision):
    from snn_modulation_classification_b200.dcll import pytorch_libdcll as L
    cin, cout, kernel, pad, pool, hw, B, K, wrp, output_layer = shape
    lay = _layer_pair(*shape)
    lay.i2h.precision = precision
    tc = lay.i2h.tensor_core_ok()
    if precision == "bf16x3" and not tc and not (cin == 1 and cout == 32):
        pytest.skip("no tensor-core instantiation for this shape: identical to the fp32 run")
    sl = L.DCLLClassification(dclllayer=lay, batch_size=B, loss=torch.nn.SmoothL1Loss, optimizer=torch.optim.SGD,
                              kwargs_optimizer={"lr": 0.0}, burnin=0)
    spec, p = _oracle_of(lay, *shape)
    st = O.zero_state(spec, B)
    g = torch.Generator().manual_seed(3)
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K)
    tol_m, tol_g = (1e-5, 2e-5) if not tc else (5e-5, 5e-5)
    for t in range(4):
        x = (torch.rand(B, cin, *hw, generator=g) < 0.2).float()
        if output_layer:      # optimizer2 is hard-wired to lr = 1e-4 (ref :636-638): keep output_ teacher-forced
            with torch.no_grad():
                lay.output_.weight.copy_(p.wout)
                lay.output_.bias.copy_(p.bout)
        out, pvo, pv, pvmem, _ = sl.train_dcll(x.cuda(), y.cuda(), regularize=False)
        fo = O.conv_step_fwd(spec, p, st, x)
        st = fo.state
        assert torch.equal(lay.i2h.state.eps0.cpu(), st.eps0) and torch.equal(lay.i2h.state.eps1.cpu(), st.eps1)
        assert pvmem.shape == fo.pvmem.shape and pv.shape == fo.pv.shape
        assert rel_err(pvmem, fo.pvmem) <= tol_m, (t, rel_err(pvmem, fo.pvmem))
        assert rel_err(pvo, fo.pvoutput) <= tol_m
        if output_layer:
            assert rel_err(out, fo.output) <= tol_m
        else:
            assert float((out.cpu() != fo.spikes).float().mean()) <= 2e-3      # tiny tensors: a single flip is ~1e-3
        if wrp > 0:
            assert float((lay.i2h.state.arp.cpu() - st.arp).abs().gt(1e-5).float().mean()) <= 2e-3
        gr = O.conv_local_grads(spec, p, fo, y)
        assert rel_err(lay.i2h.weight.grad, gr.gW) <= tol_g, (t, rel_err(lay.i2h.weight.grad, gr.gW))
        assert rel_err(lay.i2h.bias.grad, gr.gb) <= tol_g
        if output_layer:
            assert rel_err(lay.output_.weight.grad, gr.gWout) <= 1e-5
            assert rel_err(lay.output_.bias.grad, gr.gbout) <= 1e-5
    assert len(sl.clout) == 4


def test_i2h_core_alone_matches_layer():
    """ContinuousConv2D.forward / RRP.forward (ref :407-426 / :485-509) called directly, un-pooled outputs."""
    shape = (4, 12, 5, 2, 2, (14, 18), 3, 6, 2.0, False)
    lay = _layer_pair(*shape)
    spec, p = _oracle_of(lay, *shape)
    st = O.zero_state(spec, 3)
    g = torch.Generator().manual_seed(1)
    for t in range(3):
        x = (torch.rand(3, 4, 14, 18, generator=g) < 0.3).float()
        out, pv, pvmem = lay.i2h(x.cuda())
        fo = O.conv_step_fwd(spec, p, st, x)
        st = fo.state
        assert out.shape == fo.pvmem.shape                          # un-pooled
        assert rel_err(pvmem, fo.pvmem) <= 1e-5
        assert torch.equal(lay.i2h.state.eps1.cpu(), st.eps1)
        assert float((out.cpu() != (fo.pvmem > 0).float()).float().mean()) <= 1e-3
