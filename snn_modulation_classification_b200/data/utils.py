"""Spike encoders: mirror of the reference's ``data/utils.py`` (ref lines cited inline)."""
import numpy as np
import torch

from .. import _lib
from ..dcll.pytorch_libdcll import SpikeCells, _as_cuda_f32


def add_gaussian(x, gs_stdev):
    """ref:5-7 -- isotropic additive Gaussian noise."""
    return x + torch.empty_like(x).normal_(0, gs_stdev)


def to_one_hot(t, width):
    """ref:10-12."""
    t_onehot = torch.zeros(*t.shape + (width,), device=t.device)
    return t_onehot.scatter_(1, t.unsqueeze(-1), 1)


class WindowCells(SpikeCells):
    """All timesteps of a window: int32 [T, B, 2]; ``wc[t]`` is the SpikeCells of timestep t."""

    def __len__(self):
        return self.cells.shape[0]

    def __getitem__(self, t):
        return SpikeCells(self.cells[t], self.height, self.width)

    @property
    def shape(self):
        return torch.Size([self.cells.shape[0], self.cells.shape[1], 1, self.height, self.width])

    def dense(self):
        T, B = self.cells.shape[0], self.cells.shape[1]
        out = torch.empty((T, B, 1, self.height, self.width), dtype=torch.float32, device=self.cells.device)
        _lib.check(_lib.lib.dcll_cells_to_frames(_lib.ptr(self.cells), T, B, self.height, self.width, _lib.ptr(out),
                                                 _lib.current_stream()))
        return out


def iq2spiketrain(x, y, out_w=28, out_h=28, min_I=-1, max_I=1, min_Q=-1, max_Q=1, max_duration=500, do_gamma=True,
                  gs_stdev=0, as_cells=False):
    """Convert each I/Q sample to a spike in the I/Q plane over time (ref:43-87), on the GPU.

    Same arguments and the same numpy-RNG consumption (one ``randint`` draw for the window start,
    ref:58) as the reference.  Returns ``(spike_trains, all_target)``:
      * spike_trains -- float32 CUDA tensor [T, B, 1, out_h, out_w] with exactly one 1 per (t, b)
        (the reference returns the same values as a float64 numpy array that train.py:243 then
        converts and uploads), or, with ``as_cells=True``, a ``WindowCells`` holding only the int32
        [T, B, 2] (row, col) cells, which layer 0 consumes directly;
      * all_target   -- labels repeated over time [T, B, K] (ref:85), on the device of ``y``.
    """
    x = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))
    x = _as_cuda_f32(x.squeeze())
    if gs_stdev > 0:
        x = add_gaussian(x, gs_stdev)
    if x.dim() != 3 or x.shape[1] != 2:
        raise ValueError('expected x of (squeezed) shape (batch, 2, num_timesteps), got %s' % (tuple(x.shape),))
    batch_size, num_timesteps = int(x.shape[0]), int(x.shape[-1])
    assert max_duration <= num_timesteps
    t_start = int(np.random.randint(0, num_timesteps - max_duration + 1))
    cells = torch.empty((max_duration, batch_size, 2), dtype=torch.int32, device=x.device)
    _lib.check(_lib.lib.dcll_iq_encode(_lib.ptr(x), batch_size, num_timesteps, float(min_I), float(max_I), float(min_Q),
                                       float(max_Q), int(out_w), int(out_h), t_start, int(max_duration),
                                       1 if do_gamma else 0, _lib.ptr(cells), _lib.current_stream()))
    wc = WindowCells(cells, out_h, out_w)
    y_t = y if torch.is_tensor(y) else torch.as_tensor(np.asarray(y))
    all_target = y_t.unsqueeze(0).expand(max_duration, *y_t.shape)
    return (wc if as_cells else wc.dense()), all_target


def image2spiketrain(x, y, input_shape, gain=50, min_duration=None, max_duration=500, device_rng=None):
    """Frozen Poisson spike train of an image (ref:15-40), on the GPU.

    Same arguments and the same numpy-RNG consumption as the reference: one ``randint`` draw of the per-sample durations
    (ref:26), then per sample one ``uniform(size=(T_i, Nin))`` draw (ref:32).  Those host-drawn uniforms are compared with
    ``p = (1000 - gain*x)/1000`` by ``image_encode_kernel`` (csrc/encode.cu), so for a given numpy seed the spikes are the
    reference's, element for element.  ``device_rng=<seed>`` skips the host draws after the durations and uses the kernel's
    counter-based generator instead (same distribution, not the numpy stream) -- no 8-byte-per-element upload.

    Returns ``(spike_trains, all_target)``: a float32 CUDA tensor [max_duration, B, *input_shape] (the reference returns the
    same values as a float64 numpy array that train.py then converts and uploads) and the labels repeated over time."""
    if min_duration is None:
        min_duration = max_duration - 1
    x_t = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))
    batch_size = int(x_t.shape[0])
    nin = int(np.prod(input_shape))
    x_d = _as_cuda_f32(x_t.reshape(batch_size, -1))
    if x_d.shape[1] != nin:
        raise ValueError('input_shape %r does not match x of shape %s' % (tuple(input_shape), tuple(x_t.shape)))
    T = np.random.randint(min_duration, max_duration, batch_size)          # ref:26
    u_d = None
    if device_rng is None:
        u = torch.zeros((batch_size, max_duration, nin), dtype=torch.float64,
                        pin_memory=torch.cuda.is_available())
        un = u.numpy()
        for i in range(batch_size):
            un[i, :T[i]] = np.random.uniform(size=(T[i], nin))             # ref:32, same draw order
        u_d = u.to(x_d.device, non_blocking=True)
    t_len = torch.as_tensor(T.astype(np.int32)).to(x_d.device)
    out = torch.empty((max_duration, batch_size, nin), dtype=torch.float32, device=x_d.device)
    _lib.check(_lib.lib.dcll_image_encode(_lib.ptr(x_d), _lib.ptr(u_d), _lib.ptr(t_len), batch_size, nin, int(max_duration),
                                          float(gain), int(device_rng or 0), _lib.ptr(out), _lib.current_stream()))
    y_t = y if torch.is_tensor(y) else torch.as_tensor(np.asarray(y))
    all_target = y_t.unsqueeze(0).expand(max_duration, *y_t.shape)
    return out.reshape(max_duration, batch_size, *input_shape), all_target
