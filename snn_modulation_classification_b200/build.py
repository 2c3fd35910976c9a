"""Builds libdcll_b200.so (sm_100a only) in-tree with nvcc.

    python -m snn_modulation_classification_b200.build [--force]

The shared library carries a plain C ABI (include/dcll_b200.h); it links against
cudart only.  Objects are cached under csrc/_build and rebuilt when a source or
header is newer.  The built .so stays in the package directory so that it travels
to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import concurrent.futures as cf
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_build")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")
LIB = os.path.join(PKG, "libdcll_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nccl_include():
    """nccl.h for dp.cu (types only: the library itself is dlopen'ed at run time): torch's bundled copy, else the system's."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations or []) if spec else []:
            inc = os.path.join(loc, "include")
            if os.path.isfile(os.path.join(inc, "nccl.h")):
                return ["-I", inc]
    except Exception:
        pass
    return ["-I", "/usr/include"] if os.path.isfile("/usr/include/nccl.h") else []


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(INCLUDE, "dcll_b200.h"))
    return hs


def _compile(src, force):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    deps = [os.path.join(CSRC, src)] + _headers()
    if not force and os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(d) for d in deps):
        return obj, False
    cmd = [NVCC] + FLAGS + (_nccl_include() if src == "dp.cu" else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, True


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile(s, force), _sources()))
    objs = [o for o, _ in results]
    if any(c for _, c in results) or not os.path.exists(LIB) or force:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("built", LIB)
    elif verbose:
        print("up to date:", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
