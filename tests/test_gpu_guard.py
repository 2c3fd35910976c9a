"""Guard-band checks for the tensor-core operand buffers (needs a B200: pytest -m gpu).

compute-sanitizer is not available on the GPU pool, so the kernels that write the bf16 operand images
(trace_image_kernel / trace_image1_kernel -> eps1_mma, weight_mma*_kernel and reduce_adam_kernel -> weight_mma) are
run with their buffers embedded in sentinel-filled allocations on ragged shapes: no write may land outside, and the
image must be written completely.
"""
import pytest
import torch

from oracle import dcll_oracle as O
from util_build import build_pair

pytestmark = pytest.mark.gpu

GUARD, SENT = 4096, 12345.0


def _guarded(n, dev):
    big = torch.full((n + 2 * GUARD,), SENT, dtype=torch.bfloat16, device=dev)
    return big, big[GUARD:GUARD + n]


@pytest.mark.parametrize("im,B", [((21, 45), 3), ((16, 16), 5), ((40, 24), 2)])
def test_operand_images_stay_inside_their_buffers(im, B):
    K, lr = 24, 1e-6
    net, onet = build_pair("radio_ml_conv", (1,) + im, B, K, arp=1.0, burnin=0, lr=lr)
    net.set_precision("bf16x3")
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(4, B, 1, *im, generator=g) < 0.1).float().cuda()
    y = O.to_one_hot(torch.randint(0, K, (B,), generator=g), K).cuda()
    net.reset()
    net.learn(x[0], y)                                   # allocates the kernel-side buffers
    bigs = []
    for s in net.dcll_slices:
        i2h = s.dclllayer.i2h
        assert i2h.tensor_core_ok()
        be, ve = _guarded(i2h._e1mma.numel(), x.device)
        bw, vw = _guarded(i2h._wmma.numel(), x.device)
        i2h._e1mma, i2h._wmma, i2h._wt_key = ve, vw, None   # force a fresh weight_mma through dcll_conv_sync_weights
        bigs.append((be, ve, bw, vw, i2h))
    for t in range(1, 4):
        net.learn(x[t], y)
    torch.cuda.synchronize()
    for be, ve, bw, vw, i2h in bigs:
        assert i2h._e1mma.data_ptr() == ve.data_ptr() and i2h._wmma.data_ptr() == vw.data_ptr()   # buffers were kept
        for big in (be, bw):
            assert bool((big[:GUARD] == SENT).all()) and bool((big[-GUARD:] == SENT).all()), "write outside the buffer"
        assert not bool((ve == SENT).any()), "operand image not completely written"
        # weight_mma: every element the kernels own is written; for one input channel the allocation is rounded up
        n_w = 2 * i2h.weight.numel() if i2h.in_channels > 1 else 4 * 2 * 2 * i2h.out_channels * 8
        assert not bool((vw[:n_w] == SENT).any())
