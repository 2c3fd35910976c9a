#!/usr/bin/env python
"""Static SASS instruction counts per kernel of libdcll_b200.so for the mnemonics that identify the hardware path.

    python tools/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt

No GPU needed (cuobjdump + c++filt).  Kernels without any tcgen05 / TMA / FFMA2 instruction are omitted.
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "snn_modulation_classification_b200", "libdcll_b200.so")
# column -> SASS opcode prefix
COLS = [("UTCHMMA.2CTA", "UTCHMMA.2CTA"), ("UTCHMMA", "UTCHMMA"), ("UTCBAR", "UTCBAR"), ("LDTM", "LDTM"), ("UTCATOMSWS", "UTCATOMSWS"),
        ("UTMALDG", "UTMALDG"), ("UBLKCP", "UBLKCP"), ("LDGSTS", "LDGSTS"), ("SYNCS", "SYNCS"), ("USETMAXREG", "USETMAXREG"),
        ("UCGABAR", "UCGABAR_"), ("FFMA2", "FFMA2"), ("F2FP", "F2FP"), ("F2F", "F2F."), ("ACQBULK", "ACQBULK")]
KEY = ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "FFMA2", "LDTM")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            for col, pre in COLS:
                if col == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                    continue
                if op.startswith(pre):
                    counts[cur][col] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass libdcll_b200.so (sm_100a), final round-2 code: static instruction counts per kernel for the mnemonics that identify the hardware path.")
    print("# UTCHMMA = tcgen05.mma (.2CTA: cta_group::2, CTA pairs), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTCATOMSWS = tcgen05.alloc/dealloc, UTMALDG = cp.async.bulk.tensor (TMA tensor-map load),")
    print("# UBLKCP = cp.async.bulk (1-D bulk copy), LDGSTS = cp.async, SYNCS = mbarrier operations, USETMAXREG = setmaxnreg, UCGABAR = barrier.cluster, FFMA2 = fma.rn.f32x2,")
    print("# F2FP = packed fp32 -> bf16x2 / fp16x2 conversion, F2F = scalar conversion, ACQBULK = griddepcontrol.wait (PDL; every kernel has one).")
    print("# Kernels without any tcgen05 / TMA / FFMA2 instruction are omitted (FP32 parity kernels, dense / encode / quantise kernels).  tools/sass_mnemonics.py")
    print("%-78s" % "kernel" + "".join("%13s" % c for c, _ in COLS))
    tot = collections.Counter()
    for (mangled, c), name in sorted(zip(counts.items(), names), key=lambda t: t[1]):
        tot.update(c)
        if not any(c[k] for k in KEY):
            continue
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*$", "", name).replace("dcll::", "").replace("(int)", "").replace("(bool)", "")
        print("%-78s" % name[:78] + "".join("%13d" % c[col] for col, _ in COLS))
    print("%-78s" % "whole library (all kernels)" + "".join("%13d" % tot[col] for col, _ in COLS))


if __name__ == "__main__":
    main()
