"""RadioML 2018.01A reader: host-side mirror of the reference's ``data/load_radio_ml.py``.

Same public surface -- ``RadioMLDataset(data_dir, train, normalize=False, min_snr=6, max_snr=30, per_h5_frac=0.5,
train_frac=0.9)`` and ``get_radio_ml_loader(batch_size, train, **kwargs)`` -- and the same contents, element for element
(checked against the reference classes on an in-memory HDF5 stand-in, tests/test_host_cpu_loader.py):

* ref :20-56   the 2.5 M-record ``GOLD_XYZ_OSC.0001_1024.hdf5`` is split once into ``class<c>_snr<z>.hdf5`` files;
* ref :69-95   per (class, SNR) file the first ``int(per_h5_frac * 4096)`` records are used, the leading ``train_frac`` of
               them for training and the rest for testing, and the records are INTERLEAVED: global index
               ``n * 24 * n_snr + (class * n_snr + snr_index)`` holds record n of that pair;
* ref :98-101  batches are ``X: (B, 2, 1, 1024) float32`` (I/Q planes first, a dummy height axis) and ``Y: (B,) int64``;
               ``normalize`` maps X to [0, 1] with the min/max over every file that was opened.

What is new: the loader hands out page-locked batches when a GPU is present (``pin_memory``), so that the encoder's
``x.cuda(non_blocking=True)`` overlaps the previous window's kernels, and the HDF5 module is injectable (``h5=``) so the
host logic can be tested without h5py or the 20 GB data set.  There is no synthetic fallback in here: a missing h5py or
data directory fails loudly (``snn_modulation_classification_b200.data.synthetic`` is the explicit stand-in).
"""
import os

import numpy as np
import torch
from torch.utils import data

NUM_CLASSES = 24
RECORDS_PER_PAIR = 4096                       # per (class, SNR) pair in the 2018.01A release
GOLD_FILE = 'GOLD_XYZ_OSC.0001_1024.hdf5'
ALL_SNRS = range(-26, 32, 2)                  # the split writes one file per value in this range (ref :46)


def _h5_module(h5):
    if h5 is not None:
        return h5
    try:
        import h5py
        return h5py
    except ImportError:                        # install-free reader for the earliest-format files the split writes
        from . import minih5                   # (contiguous datasets; anything else raises NotImplementedError, never synthetic data)
        return minih5


def pair_file(data_dir, class_idx, snr):
    return os.path.join(data_dir, 'class%d_snr%d.hdf5' % (class_idx, snr))


def split_gold_file(data_dir, h5=None, verbose=True):
    """One-off split of the monolithic file into per-(class, SNR) files holding only 'X' (ref :20-56)."""
    h5 = _h5_module(h5)
    src = h5.File(os.path.join(data_dir, GOLD_FILE), 'r')
    try:
        labels = np.argmax(src['Y'], axis=1)                      # one-hot -> class index
        for c in range(NUM_CLASSES):
            of_class = labels == c
            sig = src['X'][of_class, :, :]
            snr_of = src['Z'][of_class, 0]
            for z in ALL_SNRS:
                path = pair_file(data_dir, c, z)
                out = h5.File(path, 'w')
                out.create_dataset('X', data=sig[snr_of == z, :, :])
                out.close()
                if verbose:
                    print('split: class %d, SNR %d dB -> %s' % (c, z, path))
    finally:
        src.close()


class RadioMLDataset(data.Dataset):
    """RadioML 2018.01A records of the SNR range [min_snr, max_snr] (step 2 dB), interleaved over (class, SNR)."""

    def __init__(self, data_dir, train, normalize=False, min_snr=6, max_snr=30, per_h5_frac=0.5, train_frac=0.9, h5=None):
        h5 = _h5_module(h5)
        self.train = train
        if not os.path.exists(pair_file(data_dir, NUM_CLASSES - 1, 30)):
            split_gold_file(data_dir, h5)

        snrs = list(range(min_snr, max_snr + 2, 2))
        n_snr = (max_snr - min_snr) // 2 + 1                      # ref :68 (the stride of the interleave)
        used = int(per_h5_frac * RECORDS_PER_PAIR)                 # records taken from each file
        n_train = int(train_frac * used)
        per_pair = n_train if train else used - n_train
        stride = NUM_CLASSES * n_snr

        self.X = np.zeros((stride * per_pair, 1024, 2), dtype=np.float32)
        self.Y = np.zeros(stride * per_pair, dtype=np.int64)
        lo, hi = float('inf'), float('-inf')
        for c in range(NUM_CLASSES):
            for zi, z in enumerate(snrs):
                f = h5.File(pair_file(data_dir, c, z), 'r')
                rec = f['X'][:]
                f.close()
                lo, hi = min(lo, rec.min()), max(hi, rec.max())    # over the WHOLE file, train and test alike (ref :82-83)
                rows = slice(0, n_train) if train else slice(n_train, used)
                slot = c * n_snr + zi
                self.X[slot::stride] = rec[rows]
                self.Y[slot::stride] = c
        # (N, 1024, 2) -> (N, 2, 1, 1024): I/Q first, then a dummy height axis (ref :98)
        self.X = self.X.transpose(0, 2, 1)[:, :, np.newaxis, :]
        if normalize:
            self.X = (self.X - lo) / (hi - lo)

    def __len__(self):
        return len(self.X)

    def __getitem__(self, index):
        return self.X[index], self.Y[index]


def get_radio_ml_loader(batch_size, train, **kwargs):
    """ref :111-132: shuffled for training, in file order for testing; ``loader.name`` as in the reference."""
    dataset = RadioMLDataset(kwargs['data_dir'], train, normalize=False, min_snr=kwargs.get('min_snr', 6),
                             max_snr=kwargs.get('max_snr', 30), per_h5_frac=kwargs.get('per_h5_frac', 0.5),
                             train_frac=kwargs.get('train_frac', 0.9), h5=kwargs.get('h5'))
    which = 'train' if train else 'test'
    print('RadioML %s set: %d records' % (which, len(dataset)))
    loader = data.DataLoader(dataset=dataset, batch_size=batch_size, shuffle=train,
                             pin_memory=torch.cuda.is_available())
    loader.name = 'RadioML_{}'.format(which)
    return loader
