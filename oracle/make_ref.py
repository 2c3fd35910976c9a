#!/usr/bin/env python
"""Vendors the UNMODIFIED reference sources of the hot path into oracle/_ref/ (TEST / BASELINE INFRASTRUCTURE).

    python oracle/make_ref.py [--src /root/reference]

The reference is pure Python over torch, so "building" it is a byte-for-byte copy of the few files the path lives in:

    dcll/__init__.py, dcll/pytorch_libdcll.py      the layers and DCLLClassification      (a2-a9)
    networks/__init__.py, networks/*.yaml           ConvNetwork, load_network_spec, specs  (a10)
    data/__init__.py, data/utils.py                 iq2spiketrain, to_one_hot              (a1)

oracle/_ref/ is git-ignored (reference sources never enter the history) but not gpurun-ignored, so it travels to the GPU
box like the built .so: `bench.py --impl reference` and the cpu_baseline leg import the reference classes from there
(oracle/refshim.py: apex stub, yaml Loader default, device = 'cpu' -- none touches arithmetic) and time THEM on the box's
host cores.  A MANIFEST with sha256 of every file is written next to the copies; refshim checks it on import, so a
modified copy is refused.  Called by __graft_entry__.build() when /root/reference is present.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["dcll/__init__.py", "dcll/pytorch_libdcll.py", "networks/__init__.py", "networks/radio_ml_conv.yaml",
         "networks/mnist_conv.yaml", "networks/radio_ml_conv_ref.yaml", "data/__init__.py", "data/utils.py"]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def make_ref(src="/root/reference", verbose=True):
    if not os.path.isfile(os.path.join(src, "dcll", "pytorch_libdcll.py")):
        if verbose:
            print("make_ref: no reference at %s, keeping %s as it is" % (src, DST))
        return False
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = sha256(d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("make_ref: %d reference files -> %s" % (len(FILES), DST))
    return True


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else "/root/reference"
    sys.exit(0 if make_ref(src) else 1)
