"""Oracle vs the frozen reference outputs in tests/golden/ (CPU; runs everywhere)."""
import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR, NetFixture, sub
from oracle import dcll_oracle as O


def test_encoder_golden():
    z = np.load(GOLDEN_DIR + "/encoder.npz")
    x = z["x"]
    n = 0
    for k in z.files:
        if not k.startswith("cells__"):
            continue
        key = k[len("cells__"):]
        parts = key.split("_")
        W, H, T, gamma = int(parts[0][1:]), int(parts[1][1:]), int(parts[2][1:]), bool(int(parts[3][1:]))
        bounds = [float(v) for v in key.split("_b")[1].split("_")]
        cells = O.encode_cells(x, W, H, bounds[0], bounds[1], bounds[2], bounds[3],
                               t_start=int(z["tstart__" + key]), max_duration=T, do_gamma=gamma)
        assert cells.dtype == np.int32 and np.array_equal(cells, z[k]), key
        n += 1
    assert n == 4


@pytest.mark.parametrize("name", ["radio8_train", "radio8_arp_train", "radio8_arp_infer", "mnist_train",
                                  "radioref_train"])
@pytest.mark.parametrize("backend", ["autograd", "closed"])
def test_network_golden(name, backend):
    fx = NetFixture(name)
    if not fx.train and backend == "closed":
        pytest.skip("inference has a single backend")
    specs = O.make_specs(O.BUILTIN_SPECS[fx.spec_name], fx.im_dims, fx.K, wrp=fx.arp)
    params = O.params_from_state_dict(fx.state_dict, fx.n_layers)
    net = O.OracleNet(specs, params, fx.B, burnin=fx.burnin, backend=backend)
    net.reset()
    lr = 1e-6
    first_train_t = fx.burnin - 1
    for t in range(fx.steps):
        if fx.train:
            net.learn(fx.x[t], fx.y)
        else:
            net.test(fx.x[t])
        # Free-running comparison: exact-ish until the first weight update, then only over a short
        # horizon (training amplifies rounding differences x2 per step).
        horizon = t - first_train_t if fx.train else 0
        for i in range(fx.n_layers):
            if horizon <= 1:
                np.testing.assert_allclose(sub(net.states[i].eps1), fx.get(t, i, "eps1"), rtol=1e-6, atol=0)
                if fx.arp > 0:
                    np.testing.assert_allclose(sub(net.states[i].arp), fx.get(t, i, "arp"), rtol=1e-6, atol=1e-7)
            if fx.train and 0 <= horizon <= 1:
                d = np.abs(sub(net.params[i].weight) - fx.get(t, i, "w"))
                assert d.max() <= (5e-2 * (2 ** horizon)) * lr, (name, t, i, d.max() / lr)
                db = np.abs(net.params[i].bias.detach().numpy() - fx.get(t, i, "b"))
                assert db.max() <= 5e-2 * (2 ** horizon) * lr
                if specs[i].output_layer:
                    dw = np.abs(sub(net.params[i].wout) - fx.get(t, i, "wout"))
                    assert dw.max() <= 5e-2 * 1e-4
    if not fx.train:
        for i in range(fx.n_layers):
            assert np.array_equal(np.array(net.clout[i]), fx.clout(i))
        labels = np.stack([fx.y.numpy()] * fx.steps)
        assert net.accuracy(labels) == list(fx.z["acc"])
        assert np.array_equal(O.confusion_matrix(net.clout[-1], labels, fx.K), fx.z["confusion"])
    else:
        for i in range(fx.n_layers):
            assert np.array(net.clout[i]).shape == fx.clout(i).shape


def test_vote_tie_break_first_seen():
    # Counter.most_common keeps insertion order among equal counts (dcll/pytorch_libdcll.py:51-52)
    clout = [np.array([2, 1]), np.array([3, 1]), np.array([3, 0]), np.array([2, 0])]
    lab = np.zeros((4, 2, 4))
    lab[:, 0, 2] = 1
    lab[:, 1, 0] = 1
    pred, true = O.predictions_by_vote(clout, lab)
    assert list(pred) == [2, 1] and list(true) == [2, 0]
    assert O.accuracy_by_vote(clout, lab) == 0.5
