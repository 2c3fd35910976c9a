"""Synthetic stand-in for data/load_radio_ml.py (the RadioML 2018.01A HDF5 file and h5py are not available).

Only the loader's OUTPUT LAYOUT matters to the hot path (ref data/load_radio_ml.py:99): batches of
``X: (B, 2, 1, 1024) float32`` IQ records and ``Y: (B,) int64`` class indices.  Records are class-dependent
PSK/QAM-like constellations with additive noise set by the SNR, so that accuracy above chance is possible.
"""
import numpy as np
import torch


class SyntheticRadioML(torch.utils.data.Dataset):
    def __init__(self, n, num_classes=24, n_iq=1024, snr_db=18.0, seed=0):
        rs = np.random.RandomState(seed)
        self.y = rs.randint(0, num_classes, size=n).astype(np.int64)
        t = np.arange(n_iq, dtype=np.float32)
        x = np.empty((n, 2, 1, n_iq), dtype=np.float32)
        noise = 10.0 ** (-snr_db / 20.0)
        for i in range(n):
            c = int(self.y[i])
            order = 2 + c % 6                                   # constellation order
            sps = 4 + 2 * (c // 6)                              # samples per symbol
            sym = rs.randint(0, order, size=n_iq // sps + 1)
            phase = 2 * np.pi * sym[(t // sps).astype(int)] / order + 0.05 * c
            amp = 0.35 + 0.1 * ((c * 7) % 5) / 4.0
            x[i, 0, 0] = amp * np.cos(phase) + noise * 0.3 * rs.randn(n_iq)
            x[i, 1, 0] = amp * np.sin(phase) + noise * 0.3 * rs.randn(n_iq)
        self.x = torch.from_numpy(x)

    def __len__(self):
        return len(self.y)

    def __getitem__(self, i):
        return self.x[i], int(self.y[i])


def get_radio_ml_loader(batch_size, train, n=None, min_snr=6, max_snr=30, seed=None, **kwargs):
    """Same call shape as ref data/load_radio_ml.py:111-132 (unknown kwargs such as data_dir are ignored)."""
    n = n or batch_size * 8
    snr = 0.5 * (min_snr + max_snr)
    ds = SyntheticRadioML(n, snr_db=snr, seed=(0 if train else 1) if seed is None else seed)
    return torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=train, drop_last=True)
