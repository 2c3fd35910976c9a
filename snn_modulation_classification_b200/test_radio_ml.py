"""Inference entry point: the reference's ``test_radio_ml.py`` (ref :65-188) on the B200 hot path.

Restores a ``.pth`` written by either implementation, evaluates per SNR (6..30 dB in steps of 2) with the hot
loop ``net.test(x[t])`` (ref :144-145) as one ``test_window`` call, writes ``snr_evaluation.txt``, the
confusion matrices and ``snr_evaluation_accs.npy``.  Plotting (matplotlib) is out of scope.
"""
import os
import time

import numpy as np
import torch

from .data.utils import iq2spiketrain, to_one_hot
from .dcll.pytorch_libdcll import device
from .networks import ConvNetwork, load_network_spec
from .train import get_loader, parse_args as _train_args


def parse_args(argv=None):
    return _train_args(argv)


def main(argv=None, snrs=None):
    args = parse_args(argv)
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    out_dir = os.path.dirname(args.restore_path) if args.restore_path else '.'
    log = open(os.path.join(out_dir, 'snr_evaluation.txt'), 'a+')

    def print_and_log(text):
        print(text)
        log.write(text + '\n')

    im_dims = (1, args.Q_resolution, args.I_resolution)
    target_size = 24
    st_kw = dict(out_w=args.I_resolution, out_h=args.Q_resolution, min_I=args.I_bounds[0], max_I=args.I_bounds[1],
                 min_Q=args.Q_bounds[0], max_Q=args.Q_bounds[1], max_duration=args.n_iters_test, as_cells=True)
    n_test = int(np.ceil(float(args.n_test_samples) / args.batch_size_test))
    convs = load_network_spec(args.network_spec)
    net = ConvNetwork(args, im_dims, args.batch_size_test, convs, target_size, act=torch.nn.Sigmoid(), loss=None,
                      opt=None, opt_param={}, learning_rates=None, burnin=args.burnin)          # ref :92-95
    if args.restore_path:
        if not os.path.isfile(args.restore_path):
            print_and_log('ERROR: Cannot load `%s`. File does not exist! Aborting...' % args.restore_path)
            return None
        net.load_state_dict(torch.load(args.restore_path))
        print_and_log('Loaded the SNN model from `%s`.' % args.restore_path)
    net = net.to(device)
    net.reset(True)                                                     # ref :110 (RRP cores re-draw tau here)

    accs = []
    snrs = np.array(range(6, 32, 2)) if snrs is None else np.array(snrs)
    total_cm = np.zeros((target_size, target_size), dtype=int)
    for snr in snrs:
        start = time.time()
        gen_test = iter(get_loader(args.batch_size_test, train=False, data_dir=args.radio_ml_data_dir, min_snr=int(snr),
                                   max_snr=int(snr), per_h5_frac=args.per_h5_frac, train_frac=args.train_frac,
                                   synthetic=args.synthetic,
                                   n=n_test * args.batch_size_test))
        data = [next(gen_test) for _ in range(n_test)]
        data = [(x, to_one_hot(y, target_size)) for x, y in data]
        acc_test = np.zeros([n_test, len(net.dcll_slices)])
        cm = np.zeros((target_size, target_size), dtype=int)
        for i, (x, y) in enumerate(data):
            test_input, test_labels = iq2spiketrain(x, y.to(device), **st_kw)
            net.reset()                                                 # ref :142-147
            net.eval()
            net.test_window(test_input)
            acc_test[i, :] = net.accuracy(test_labels)
            cm += net.confusion_matrix(test_labels)
        acc = np.mean(acc_test, axis=0)
        print_and_log('SNR {} \t Accuracy {} \t Time Elapsed {}'.format(str(snr).zfill(2), acc, '%.2f s' % (time.time() - start)))
        np.save(os.path.join(out_dir, 'confusion_matrix_snr_%d.npy' % snr), cm)
        accs.append(acc)
        total_cm += cm
    print_and_log('---\nTotal confusion matrix:')
    print_and_log(np.array2string(total_cm, max_line_width=300))
    np.save(os.path.join(out_dir, 'snr_evaluation_accs.npy'), accs)
    log.close()
    return dict(accs=np.array(accs), confusion=total_cm, net=net)


if __name__ == '__main__':
    main()
