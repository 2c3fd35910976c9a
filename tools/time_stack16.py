import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from snn_modulation_classification_b200 import _lib
from snn_modulation_classification_b200.data.utils import iq2spiketrain

def main(B=4096, T=64, tc=16):
    bench.WORKLOADS["x"] = ("radio_ml_conv", 16, B, False, 1.0, 20)
    net = bench.build_net("x"); net.set_precision("bf16x3")
    x, y = bench.synth(B, 1)
    np.random.seed(1)
    cells, _ = iq2spiketrain(x.cuda(), y.cuda(), out_w=16, out_h=16, max_duration=T, as_cells=True)
    net.reset(); net._run_stack16(cells, chunk=tc); torch.cuda.synchronize()
    # time phases by monkeypatching the two entry points
    lib = _lib.lib
    ev = lambda: torch.cuda.Event(enable_timing=True)
    times = {"stack": 0.0, "readout": 0.0}
    orig_s, orig_r = lib.dcll_infer_stack16, lib.dcll_conv_readout_rows
    recs = []
    def wrap(name, fn):
        def f(*a):
            e0, e1 = ev(), ev(); e0.record(); rc = fn(*a); e1.record(); recs.append((name, e0, e1)); return rc
        return f
    lib.dcll_infer_stack16 = wrap("stack", orig_s); lib.dcll_conv_readout_rows = wrap("readout", orig_r)
    net.reset(); net._run_stack16(cells, chunk=tc); torch.cuda.synchronize()
    for name, e0, e1 in recs: times[name] += e0.elapsed_time(e1)
    lib.dcll_infer_stack16, lib.dcll_conv_readout_rows = orig_s, orig_r
    print("B=%d T=%d tc=%d: stack %.2f ms/timestep, readouts %.2f ms/timestep" % (B, T, tc, times["stack"] / T, times["readout"] / T))

if __name__ == "__main__":
    main(4096, 64, 16)
    main(4096, 64, 4)
    main(148, 64, 16)
