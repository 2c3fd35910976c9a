"""FP32 parity mode vs bf16x3 tensor-core mode: train the same network on the same synthetic RadioML-shaped windows
and compare held-out vote accuracy (BASELINE.json north_star: 'top-1 accuracy on held-out synthetic data within
0.5 pt after a fixed number of steps').  Writes a small report to stdout."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from util_build import make_args
from snn_modulation_classification_b200 import networks as N
from snn_modulation_classification_b200.data.synthetic import SyntheticRadioML
from snn_modulation_classification_b200.data.utils import iq2spiketrain, to_one_hot

def run(mode, res=16, B=256, T=300, burnin=50, n_train=24, n_test=4, K=24, lr=1e-6, seed=1, arp=1.0, snr=18.0):
    torch.manual_seed(seed); np.random.seed(seed)
    net = N.ConvNetwork(make_args(arp), (1, res, res), B, N.load_network_spec("radio_ml_conv"), K, act=torch.nn.Sigmoid(),
                        loss=torch.nn.SmoothL1Loss, opt=torch.optim.Adam, opt_param={"betas": [0.0, 0.95], "weight_decay": 10.0},
                        learning_rates=[lr], burnin=burnin)
    net.reset(True); net.set_precision(mode)
    train = SyntheticRadioML(B * n_train, snr_db=snr, seed=10)
    test = SyntheticRadioML(B * n_test, snr_db=snr, seed=11)
    kw = dict(out_w=res, out_h=res, max_duration=T, as_cells=True)
    accs_train = []
    for i in range(n_train):
        x = train.x[i*B:(i+1)*B]; y = to_one_hot(torch.from_numpy(train.y[i*B:(i+1)*B]), K).cuda()
        np.random.seed(100 + i)
        cells, tgt = iq2spiketrain(x, y, **kw)
        net.reset(); net.learn_window(cells, y)
        accs_train.append(net.accuracy(tgt))
    accs = []
    for i in range(n_test):
        x = test.x[i*B:(i+1)*B]; y = to_one_hot(torch.from_numpy(test.y[i*B:(i+1)*B]), K).cuda()
        np.random.seed(500 + i)
        cells, tgt = iq2spiketrain(x, y, **kw)
        net.reset(); net.test_window(cells)
        accs.append(net.accuracy(tgt))
    return np.mean(accs, axis=0), np.mean(accs_train[-4:], axis=0)

def seed_sweep(snr=6.0, seeds=(1, 2, 3, 4, 5)):
    """Mid-SNR accuracy over several network-initialisation seeds per mode: training is chaotic, so single runs of the
    two modes (or of one mode with another summation order) differ by points; the distributions are what compare."""
    out = {}
    for mode in ("fp32", "bf16x3"):
        accs = np.array([run(mode, snr=snr, seed=sd)[0] for sd in seeds])
        out[mode] = dict(test_acc_per_layer_by_seed=[[round(float(a), 4) for a in r] for r in accs],
                         mean=[round(float(a), 4) for a in accs.mean(0)], std=[round(float(a), 4) for a in accs.std(0)])
    print(json.dumps(dict(config="seed sweep: radio_ml_conv 16x16, B=256, T=300, 24 training windows, 1024 held-out samples, "
                                 "%g dB, seeds %s" % (snr, list(seeds)), **out)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "seeds":
        seed_sweep()
        sys.exit(0)
    for snr, n_train in ((18.0, 24), (6.0, 24), (0.0, 24)):
        out = {}
        for mode in ("fp32", "bf16x3"):
            t0 = time.time()
            test_acc, train_acc = run(mode, snr=snr, n_train=n_train)
            out[mode] = dict(test_acc_per_layer=[round(float(a), 4) for a in test_acc],
                             train_acc_last4_per_layer=[round(float(a), 4) for a in train_acc], seconds=round(time.time() - t0, 1))
        d = [round(abs(a - b), 4) for a, b in zip(out["fp32"]["test_acc_per_layer"], out["bf16x3"]["test_acc_per_layer"])]
        print(json.dumps(dict(config="radio_ml_conv 16x16, B=256, T=300, burnin=50, arp=1, %d training windows, 1024 held-out samples, "
                                     "synthetic constellations at %g dB, chance=0.0417" % (n_train, snr),
                              **out, abs_diff_test_acc=d)), flush=True)
