"""Generates tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

The reference ships no golden vectors; these fixtures freeze
what its own classes compute on seeded synthetic inputs so that the oracle and the
CUDA path can be checked where /root/reference is absent (the GPU box).
Fixtures are kept small: binary inputs are bit-packed, large tensors are stored
as strided samples (every ``SUB``-th element of the flattened tensor).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402

SUB = 7


def sub(t):
    return t.detach().reshape(-1)[::SUB].numpy().copy()


def golden_encoder():
    _, _, U = refshim.load_reference()
    g = torch.Generator().manual_seed(11)
    B, N = 32, 256                      # B % 32 == 0: torch's vectorised pow body only
    x = (torch.randn(B, 2, 1, N, generator=g) * 0.45).float()
    x[0, 0, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, -0.0, 1.5, -2.0, 0.999999, -0.999999])
    y = torch.zeros(B, 24)
    out = {"x": x.numpy()}
    for W, H, T, gamma, bounds in [(16, 16, 256, True, (-1, 1, -1, 1)), (128, 128, 200, True, (-1, 1, -1, 1)),
                                   (28, 20, 256, False, (-1, 1, -1, 1)), (16, 16, 256, True, (-0.8, 1.3, -2.0, 0.5))]:
        np.random.seed(4)
        t_start = np.random.randint(0, N - T + 1)
        np.random.seed(4)
        frames, _ = U.iq2spiketrain(x, y, out_w=W, out_h=H, min_I=bounds[0], max_I=bounds[1], min_Q=bounds[2],
                                    max_Q=bounds[3], max_duration=T, do_gamma=gamma)
        assert frames.sum() == T * B
        flat = frames.reshape(T, B, H * W).argmax(-1)
        cells = np.stack([flat // W, flat % W], -1).astype(np.int32)
        key = "W%d_H%d_T%d_g%d_b%s" % (W, H, T, gamma, "_".join(str(b) for b in bounds))
        out["cells__" + key] = cells
        out["tstart__" + key] = np.int64(t_start)
    np.savez_compressed(os.path.join(HERE, "encoder.npz"), **out)


def golden_network(name, spec, im_dims, B, K, arp, burnin, steps, train):
    net = refshim.build_reference_net(spec, im_dims, B, K, arp=arp, burnin=burnin, train=train)
    out = {"meta": np.array([B, K, burnin, steps, int(train)], dtype=np.int64), "arp": np.float64(arp),
           "im_dims": np.array(im_dims, dtype=np.int64)}
    for k, v in net.state_dict().items():
        out["sd__" + k] = v.detach().numpy().copy()
    g = torch.Generator().manual_seed(21)
    x = (torch.rand(steps, B, *im_dims, generator=g) < 0.08).float()
    lab = torch.randint(0, K, (B,), generator=g)
    y = torch.zeros(B, K).scatter_(1, lab.unsqueeze(-1), 1)
    out["x_packed"] = np.packbits(x.numpy().astype(np.uint8).reshape(-1))
    out["x_shape"] = np.array(x.shape, dtype=np.int64)
    out["labels"] = lab.numpy()
    net.reset()
    net.train() if train else net.eval()
    for t in range(steps):
        if train:
            net.learn(x[t], y)
        else:
            net.test(x[t])
        for i, s in enumerate(net.dcll_slices):
            lay = s.dclllayer
            pre = "t%d_l%d_" % (t, i)
            # re-run the read-out on the stored state is not possible (weights moved); record from the module
            out[pre + "eps1"] = sub(lay.i2h.state.eps1)
            out[pre + "w"] = sub(lay.i2h.weight)
            out[pre + "b"] = lay.i2h.bias.detach().numpy().copy()
            if arp > 0:
                out[pre + "arp"] = sub(lay.i2h.state.arp)
            if lay.output_layer:
                out[pre + "wout"] = sub(lay.output_.weight)
    for i, s in enumerate(net.dcll_slices):
        out["clout_l%d" % i] = np.array(s.clout, dtype=np.int64)
    labels = torch.stack([y] * steps)
    out["acc"] = np.array(net.accuracy(labels))
    out["confusion"] = net.confusion_matrix(labels)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(1)            # fixed summation order for the fixtures
    golden_encoder()
    # small I/Q planes keep the frozen read-out matrices (24 x F) small
    golden_network("radio8_train", "radio_ml_conv", (1, 8, 8), 4, 24, 0.0, 3, 6, True)
    golden_network("radio8_arp_train", "radio_ml_conv", (1, 8, 8), 4, 24, 1.0, 3, 6, True)
    golden_network("radio8_arp_infer", "radio_ml_conv", (1, 8, 8), 4, 24, 1.0, 3, 10, False)
    golden_network("mnist_train", "mnist_conv", (1, 28, 28), 4, 10, 0.0, 2, 5, True)
    golden_network("radioref_train", "radio_ml_conv_ref", (1, 1, 128), 2, 24, 0.0, 2, 4, True)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
