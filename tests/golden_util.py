"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py from the live reference)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUB = 7
SPEC_OF = {"radio8_train": "radio_ml_conv", "radio8_arp_train": "radio_ml_conv", "radio8_arp_infer": "radio_ml_conv",
           "mnist_train": "mnist_conv", "radioref_train": "radio_ml_conv_ref"}


def sub(t):
    return t.detach().cpu().reshape(-1)[::SUB].numpy()


class NetFixture:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.name = name
        self.spec_name = SPEC_OF[name]
        self.B, self.K, self.burnin, self.steps, train = [int(v) for v in z["meta"]]
        self.train = bool(train)
        self.arp = float(z["arp"])
        self.im_dims = tuple(int(v) for v in z["im_dims"])
        shape = tuple(int(v) for v in z["x_shape"])
        n = int(np.prod(shape))
        self.x = torch.from_numpy(np.unpackbits(z["x_packed"])[:n].reshape(shape).astype(np.float32))
        self.labels = torch.from_numpy(z["labels"])
        self.y = torch.zeros(self.B, self.K).scatter_(1, self.labels.unsqueeze(-1), 1)
        self.state_dict = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}
        self.n_layers = 1 + max(int(k.split(".")[1]) for k in self.state_dict)

    def get(self, t, layer, what):
        return self.z["t%d_l%d_%s" % (t, layer, what)]

    def clout(self, layer):
        return self.z["clout_l%d" % layer]
