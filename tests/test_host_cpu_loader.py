"""RadioML reader: host logic of data/load_radio_ml.py on an in-memory HDF5 stand-in.

h5py and the 20 GB data set are not available, so the HDF5 module is replaced by a dictionary-backed fake with the four
calls the reader makes (File, create_dataset, item access, close).  When /root/reference is present (build container) the
reference's own RadioMLDataset runs on the SAME fake files and the two are compared element for element; everywhere the
interleave / split / layout properties are checked directly.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = os.path.join(os.environ.get("DCLL_REFERENCE_ROOT", "/root/reference"), "data", "load_radio_ml.py")


class FakeH5:
    """Dictionary-backed stand-in for the h5py module; 'w' also touches the path so os.path.exists() sees the file."""

    def __init__(self):
        self.store = {}
        outer = self

        class File:
            def __init__(self, path, mode='r'):
                self.path, self.mode = path, mode
                if mode == 'w':
                    outer.store[path] = {}
                    open(path, 'wb').close()
                elif path not in outer.store:
                    raise OSError('no such file: %s' % path)

            def create_dataset(self, name, data):
                outer.store[self.path][name] = np.array(data)

            def __getitem__(self, name):
                return outer.store[self.path][name]

            def close(self):
                pass

        self.File = File


def _write_pairs(h5, data_dir, snrs, n_rec, seed=0):
    rs = np.random.RandomState(seed)
    for c in range(24):
        for z in snrs:
            f = h5.File(os.path.join(data_dir, 'class%d_snr%d.hdf5' % (c, z)), 'w')
            # value encodes (class, snr, record) so that misplaced records are visible, plus noise for min/max
            x = rs.randn(n_rec, 1024, 2).astype(np.float32) * 0.1
            x[:, 0, 0] = c
            x[:, 0, 1] = z
            x[:, 1, 0] = np.arange(n_rec)
            f.create_dataset('X', data=x)
            f.close()


def _reference_dataset_class(h5):
    if not os.path.isfile(REF):
        return None
    mod = types.ModuleType('h5py')
    mod.File = h5.File
    saved = sys.modules.get('h5py')
    sys.modules['h5py'] = mod
    try:
        spec = importlib.util.spec_from_file_location('_ref_load_radio_ml', REF)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        if saved is None:
            sys.modules.pop('h5py', None)
        else:
            sys.modules['h5py'] = saved
    return m.RadioMLDataset


@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("normalize", [False, True])
def test_dataset_layout_split_and_interleave(tmp_path, train, normalize):
    from snn_modulation_classification_b200.data.load_radio_ml import RadioMLDataset
    h5 = FakeH5()
    d = str(tmp_path)
    snrs = [26, 28, 30]
    frac = 20 / 4096.0                                            # 20 records per pair, 18 train + 2 test
    _write_pairs(h5, d, snrs, n_rec=24)
    ds = RadioMLDataset(d, train, normalize=normalize, min_snr=26, max_snr=30, per_h5_frac=frac, train_frac=0.9, h5=h5)
    per_pair = 18 if train else 2
    assert len(ds) == 24 * 3 * per_pair
    assert ds.X.shape == (len(ds), 2, 1, 1024) and ds.X.dtype == np.float32 and ds.Y.dtype == np.int64
    if not normalize:
        # global index n * 72 + (class * 3 + snr_index) holds record n (train) / 18 + n (test) of that pair
        for idx in (0, 1, 2, 3, 71, 72, 100, len(ds) - 1):
            slot, n = idx % 72, idx // 72
            c, zi = slot // 3, slot % 3
            x, y = ds[idx]
            assert y == c and x[0, 0, 0] == c and x[1, 0, 0] == snrs[zi] and x[0, 0, 1] == (n if train else 18 + n)
    else:
        assert float(ds.X.min()) >= 0.0 and float(ds.X.max()) <= 1.0
    ref_cls = _reference_dataset_class(h5)
    if ref_cls is not None:
        ref = ref_cls(d, train, normalize=normalize, min_snr=26, max_snr=30, per_h5_frac=frac, train_frac=0.9)
        assert np.array_equal(ref.X, ds.X) and np.array_equal(ref.Y, ds.Y) and ref.X.dtype == ds.X.dtype


def test_gold_file_split_and_loader(tmp_path, capsys):
    from snn_modulation_classification_b200.data import load_radio_ml as L
    h5 = FakeH5()
    d = str(tmp_path)
    rs = np.random.RandomState(1)
    snr_all = list(range(-26, 32, 2))
    n = 24 * len(snr_all) * 2
    cls = np.repeat(np.arange(24), len(snr_all) * 2)
    z = np.tile(np.repeat(snr_all, 2), 24)
    f = h5.File(os.path.join(d, L.GOLD_FILE), 'w')
    x = rs.randn(n, 1024, 2).astype(np.float32)
    f.create_dataset('X', data=x)
    f.create_dataset('Y', data=np.eye(24, dtype=np.int64)[cls])
    f.create_dataset('Z', data=z[:, None].astype(np.int64))
    f.close()
    loader = L.get_radio_ml_loader(5, train=False, data_dir=d, min_snr=28, max_snr=30, per_h5_frac=2 / 4096.0, train_frac=0.5,
                                   h5=h5)
    assert os.path.exists(L.pair_file(d, 23, 30)) and loader.name == 'RadioML_test'
    assert len(loader.dataset) == 24 * 2 * 1
    xb, yb = next(iter(loader))
    assert xb.shape == (5, 2, 1, 1024) and xb.dtype == torch.float32 and yb.dtype == torch.int64      # ref :98-101
    # test split = second record of each pair, in (class, snr) order
    first = x[(cls == 0) & (z == 28)][1]
    assert np.array_equal(xb[0, :, 0, :].numpy(), first.T) and yb.tolist() == [0, 0, 1, 1, 2]
    ref_cls = _reference_dataset_class(h5)
    if ref_cls is not None:
        ref = ref_cls(d, True, min_snr=28, max_snr=30, per_h5_frac=2 / 4096.0, train_frac=0.5)
        mine = L.RadioMLDataset(d, True, min_snr=28, max_snr=30, per_h5_frac=2 / 4096.0, train_frac=0.5, h5=h5)
        assert np.array_equal(ref.X, mine.X) and np.array_equal(ref.Y, mine.Y)


def test_missing_h5py_fails_loudly(tmp_path):
    from snn_modulation_classification_b200.data.load_radio_ml import RadioMLDataset
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py is installed here")
    except ImportError:
        pass
    with pytest.raises(ImportError, match="h5py"):
        RadioMLDataset(str(tmp_path), True)
