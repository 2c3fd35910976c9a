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


def image2spiketrain(x, y, input_shape, gain=50, min_duration=None, max_duration=500):
    """Frozen Poisson spike train of an image (ref:15-40).  Host-side data generation (numpy RNG, as the
    reference); used only to shape the MNIST-style input of mnist_conv.yaml, not a hot-path kernel."""
    if min_duration is None:
        min_duration = max_duration - 1
    batch_size = x.shape[0]
    nin = int(np.prod(input_shape))
    rates = gain * np.asarray(x).reshape(batch_size, -1)
    p = (1000.0 - rates) / 1000
    T = np.random.randint(min_duration, max_duration, batch_size)
    all_inputs = np.zeros((max_duration, batch_size, nin))
    for i in range(batch_size):
        spikes = np.ones((T[i], nin))
        spikes[(np.random.uniform(size=(T[i], nin)) < p[i]).astype('bool')] = 0
        all_inputs[:T[i], i, :] = spikes
    all_inputs = all_inputs.reshape(max_duration, batch_size, *input_shape)
    all_target = np.repeat(np.asarray(y)[np.newaxis, :, :], max_duration, axis=0)
    return all_inputs, all_target
