"""Local read-outs through the C ABI (dcll_conv_readout_rows) vs float64 torch, ragged shapes (needs a B200: pytest -m gpu).

Reference: dcll/pytorch_libdcll.py:602-606 (pvoutput = i2o(flat(pv)), output = output_(flat(pv))) and :725-728 (argmax
into clout).  FP32 mode runs readout_fwd_kernel (FMA pipe), bf16x3 mode readout_tc_kernel (tcgen05, split-bf16 x3);
both reduce their partial blocks in a fixed order.
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

# rows: below / at / above one 128-row MMA tile, many tiles;  (Cout, H, W): F with and without a 32-feature tail;
# K: 10 (MNIST), 24 (RadioML), with and without the trainable output_ (Ktot = 2K)
CASES = [(1, (4, 3, 3), 10, 0), (5, (32, 16, 16), 24, 1), (127, (8, 5, 5), 24, 0), (128, (32, 16, 16), 24, 0),
         (300, (32, 7, 9), 10, 1), (1000, (4, 16, 16), 24, 1), (64, (32, 40, 24), 24, 1), (130, (1, 2, 2), 24, 1)]


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("rows,chw,K,outl", CASES)
def test_readout_rows_matches_float64(rows, chw, K, outl, precision):
    from snn_modulation_classification_b200 import _lib
    C, H, W = chw
    F = C * H * W
    g = torch.Generator().manual_seed(rows * 7 + F)
    pv = torch.rand(rows, F, generator=g).cuda()
    wo = ((torch.rand(K, F, generator=g) - 0.5) * 0.1).cuda()
    bo = (torch.rand(K, generator=g) - 0.5).cuda()
    wout = ((torch.rand(K, F, generator=g) - 0.5) * 0.1).cuda() if outl else None
    bout = (torch.rand(K, generator=g) - 0.5).cuda() if outl else None
    pvo = torch.full((rows, K), float("nan"), device="cuda")
    out = torch.full((rows, K), float("nan"), device="cuda") if outl else None
    clout = torch.full((rows,), -1, dtype=torch.int32, device="cuda")
    d = _lib.ConvLayer()
    # geometry such that the conv grid is (C, H, W): 1x1 kernel, no padding, no pooling
    d.B, d.Cin, d.H, d.W, d.Cout, d.KH, d.KW, d.padH, d.padW, d.poolH, d.poolW = rows, 1, H, W, C, 1, 1, 0, 0, 1, 1
    d.K, d.output_layer = K, outl
    d.precision = _lib.PREC_BF16X3 if precision == "bf16x3" else _lib.PREC_FP32
    d.pv, d.wo, d.bo, d.wout, d.bout = _lib.ptr(pv), _lib.ptr(wo), _lib.ptr(bo), _lib.ptr(wout), _lib.ptr(bout)
    d.pvoutput, d.output = _lib.ptr(pvo), _lib.ptr(out)
    ws = torch.empty(_lib.lib.dcll_conv_workspace_bytes(ctypes.byref(d)), dtype=torch.uint8, device="cuda")
    d.workspace, d.workspace_bytes = _lib.ptr(ws), ws.numel()
    _lib.check(_lib.lib.dcll_conv_readout_rows(ctypes.byref(d), _lib.ptr(clout), _lib.current_stream()))
    torch.cuda.synchronize()
    ref = pv.double() @ wo.double().t() + bo.double()
    scale = float(ref.abs().max())
    tol = 2e-6 if precision == "fp32" else 2e-5          # relative to the tensor scale; bf16x3 drops the lo*lo product
    assert float((pvo.double() - ref).abs().max()) <= tol * scale
    last = ref
    if outl:
        ref2 = pv.double() @ wout.double().t() + bout.double()
        assert float((out.double() - ref2).abs().max()) <= tol * float(ref2.abs().max())
        last = ref2
    # argmax of output_ on the output layer, else of pvoutput; ties / near-ties are excluded from the comparison
    top2 = last.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 10 * tol * scale
    assert torch.equal(clout.long()[clear], last.argmax(1)[clear])
    assert int(clout.min()) >= 0 and int(clout.max()) < K
