"""Entry points train.py / test_radio_ml.py on synthetic data."""
import os

import numpy as np
import pytest
import torch


def test_cli_flags_match_reference_defaults():
    from snn_modulation_classification_b200 import train
    a = train.parse_args([])
    assert (a.I_resolution, a.Q_resolution, a.burnin, a.batch_size, a.n_iters) == (128, 128, 50, 64, 1024)
    assert a.learning_rates == [1e-6] and a.beta == .95 and a.arp == 0 and a.random_tau is True
    assert a.network_spec == 'networks/radio_ml_conv.yaml' and a.loss_type == 'SmoothL1Loss'


def test_synthetic_loader_layout():
    from snn_modulation_classification_b200.data.synthetic import get_radio_ml_loader
    x, y = next(iter(get_radio_ml_loader(8, train=True, data_dir='nowhere')))
    assert x.shape == (8, 2, 1, 1024) and x.dtype == torch.float32 and y.dtype == torch.int64   # ref load_radio_ml.py:99


@pytest.mark.gpu
def test_train_save_restore_evaluate(tmp_path):
    from snn_modulation_classification_b200 import test_radio_ml, train
    common = ['--I_resolution', '16', '--Q_resolution', '16', '--burnin', '4', '--batch_size', '8', '--batch_size_test', '8',
              '--n_test_samples', '8', '--n_iters', '12', '--n_iters_test', '12', '--arp', '1.0', '--radio_ml_data_dir',
              str(tmp_path / 'no_such_dir'), '--synthetic', '--output', str(tmp_path / 'results')]
    os.chdir(tmp_path)
    r = train.main(common + ['--n_steps', '3', '--n_test_interval', '2'])
    saved = sorted(f for f in os.listdir(r['out_dir']) if f.endswith('.pth'))
    assert saved == ['parameters_0.pth', 'parameters_2.pth']
    assert len(r['acc_train']) == 3 and np.isfinite(r['acc_test'][0]).all()
    sd = torch.load(os.path.join(r['out_dir'], saved[-1]))
    assert 'dcll_slices.2.dclllayer.output_.weight' in sd and sd['dcll_slices.1.dclllayer.i2h.alpha'].shape == (32, 16, 16)
    # weights moved (Adam steps were applied) and the .pth holds what the live network holds
    live = r['net'].state_dict()
    assert torch.equal(sd['dcll_slices.1.dclllayer.i2h.weight'], live['dcll_slices.1.dclllayer.i2h.weight'].cpu())
    ev = test_radio_ml.main(common + ['--restore_path', os.path.join(r['out_dir'], saved[-1])], snrs=[6, 18])
    assert ev['accs'].shape == (2, 3) and ev['confusion'].sum() == 16
    assert os.path.isfile(os.path.join(r['out_dir'], 'snr_evaluation.txt'))
    # restored weights are the saved ones (time constants of RRP cores are re-drawn by reset(True): reference quirk)
    assert torch.equal(ev['net'].state_dict()['dcll_slices.0.dclllayer.i2h.weight'].cpu(), sd['dcll_slices.0.dclllayer.i2h.weight'])
    # lr halving (ref train.py:221-227) touches the live optimizer objects
    s = r['net'].dcll_slices[0]
    s.optimizer.param_groups[-1]['lr'] /= 2
    assert s.optimizer.param_groups[-1]['lr'] == 0.5e-6


def test_missing_data_dir_fails_loudly(tmp_path):
    """A wrong --radio_ml_data_dir must not silently fall back to synthetic data (the reference fails loudly)."""
    from snn_modulation_classification_b200.train import get_loader, parse_args
    with pytest.raises(FileNotFoundError):
        get_loader(8, train=True, data_dir=str(tmp_path / 'no_such_dir'))
    assert parse_args([]).synthetic is False and parse_args(['--synthetic']).synthetic is True
    x, y = next(iter(get_loader(8, train=True, synthetic=True, data_dir=str(tmp_path / 'no_such_dir'))))
    assert x.shape == (8, 2, 1, 1024)
