"""Quantised-weight mode (BASELINE.json config 3).  The reference has no quantised path;
the scheme is defined in oracle/quant.py: per-output-channel symmetric int8, s = max|w|/127,
q = clamp(rint(w/s), -127, 127), straight-through training on float32 master weights.
Both directions run in libdcll_b200 (dcll_quantize / dcll_dequantize)."""
import torch

from . import _lib


def quantize(w):
    """float32 CUDA tensor [Cout, ...] -> (int8 codes of the same shape, float32 scales [Cout])."""
    if w.device.type != 'cuda':
        raise RuntimeError("libdcll_b200 has no CPU path: quantize() needs a CUDA tensor")
    w = w.detach().float().contiguous()
    rows, cols = int(w.shape[0]), int(w.numel() // w.shape[0])
    codes = torch.empty(w.shape, dtype=torch.int8, device=w.device)
    scales = torch.empty(rows, dtype=torch.float32, device=w.device)
    _lib.check(_lib.lib.dcll_quantize(_lib.ptr(w), rows, cols, _lib.ptr(codes), _lib.ptr(scales), _lib.current_stream()))
    return codes, scales


def dequantize(codes, scales):
    rows, cols = int(codes.shape[0]), int(codes.numel() // codes.shape[0])
    w = torch.empty(codes.shape, dtype=torch.float32, device=codes.device)
    _lib.check(_lib.lib.dcll_dequantize(_lib.ptr(codes.contiguous()), _lib.ptr(scales.contiguous()), rows, cols,
                                        _lib.ptr(w), _lib.current_stream()))
    return w


def fake_quantize(w):
    return dequantize(*quantize(w))


def enable_quantized_weights(net, enabled=True):
    """Switch every conv core of a ConvNetwork to the quantised-weight forward (master weights stay fp32)."""
    for s in net.dcll_slices:
        s.dclllayer.i2h.quantized = bool(enabled)
        s.dclllayer.i2h._wt_key = None
    return net
