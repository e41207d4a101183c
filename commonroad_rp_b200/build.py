"""In-tree build of the CUDA library (sm_100a only).  The .so is git-ignored but travels to the GPU box."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "rp_capi.cu")
OUT = os.path.join(HERE, "librp_b200.so")
DEPS = [SRC] + [os.path.join(HERE, "csrc", f) for f in ("rp_kernels.cuh", "rp_fused.cuh", "rp_cand.cuh", "rp_device.cuh")] + \
       [os.path.join(ROOT, "include", "rp_b200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              # the reference evaluates every expression as separate IEEE multiplies/adds: no contraction
              "--fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def build_library(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc"), "-o", OUT, SRC]
    subprocess.run(cmd, check=True)
    return OUT


PACK_SRC = os.path.join(HERE, "csrc", "rp_pack.c")


def pack_module_path():
    import sysconfig
    return os.path.join(HERE, "_rp_pack" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_pack_module(force=False):
    """The CPython helper for plan()'s output packing (csrc/rp_pack.c); host glue, optional at run time."""
    import sysconfig
    out = pack_module_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(PACK_SRC):
        return out
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"]]
    try:            # with the numpy headers the helper also creates the states' position arrays
        import numpy as np
        if os.path.exists(os.path.join(np.get_include(), "numpy", "arrayobject.h")):
            cmd += ["-DRP_PACK_NUMPY", "-I", np.get_include()]
    except ImportError:
        pass
    subprocess.run(cmd + ["-o", out, PACK_SRC, "-lm"], check=True)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_pack_module(force=True))
