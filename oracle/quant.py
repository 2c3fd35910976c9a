"""Defining restatement of the quantised-weight path -- TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED: the reference contains no quantised-weight code at all ("(quantized)" appears only in
README.md:3; apex.amp is imported but every use is commented out, dcll/pytorch_libdcll.py:705-710).
BASELINE.json config 3 asks for bit-exact weight codes "vs reference", so this file *is* the definition
the CUDA kernels (dcll_quantize / dcll_dequantize) are held to:

  per output channel r (row of the [Cout, Cin*KH*KW] matrix):
      s_r   = max_j |w[r,j]| / 127        (float32 division; s_r = 1 when the row is all zero)
      q[r,j] = clamp(round_half_even(w[r,j] / s_r), -127, 127)   as int8
      w~[r,j] = float32(q[r,j]) * s_r

Training keeps float32 master weights: the forward convolution uses w~, the local gradient is taken as if
w.r.t. w~ (straight-through) and Adam updates the master weights.
"""
import numpy as np


def quantize(w):
    w = np.asarray(w, dtype=np.float32)
    rows = w.reshape(w.shape[0], -1)
    amax = np.abs(rows).max(axis=1)
    scale = np.where(amax > 0, amax / np.float32(127.0), np.float32(1.0)).astype(np.float32)
    q = np.rint(rows / scale[:, None])          # numpy rint = round half to even, like rintf
    q = np.clip(q, -127, 127).astype(np.int8)
    return q.reshape(w.shape), scale


def dequantize(q, scale):
    q = np.asarray(q)
    rows = q.reshape(q.shape[0], -1).astype(np.float32)
    return (rows * np.asarray(scale, dtype=np.float32)[:, None]).reshape(q.shape)


def fake_quantize(w):
    return dequantize(*quantize(w))
