"""Builds csrc/nmpc_kernels.cu into libnmpc_b200.so (in-tree) with nvcc for sm_100a.

Replaces the `cargo build` of the OpEn-generated crate that MpcModule.build() triggers
(src/mpc/mpc_generator.py:188-193).  --fmad=false + explicit fma() is the arithmetic
contract."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "nmpc_kernels.cu")
DEV = os.path.join(_HERE, "csrc", "nmpc_device.cuh")
HDR = os.path.join(os.path.dirname(_HERE), "include", "nmpc_b200.h")
LIB = os.path.join(_HERE, "libnmpc_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "128"]


def find_nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def sources():
    """Every file the library is compiled from (csrc/*.cu, csrc/*.cuh and the public header)."""
    import glob
    return sorted(glob.glob(os.path.join(_HERE, "csrc", "*.cu*"))) + [HDR]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in sources())


def build_library(force=False, verbose=False):
    """Compile the CUDA library if missing or older than its sources; returns its path."""
    if not force and not is_stale():
        return LIB
    return build_variant(LIB, verbose=verbose)


def build_variant(out, defines=(), verbose=False):
    """Compile csrc/ into `out` with extra -D defines (tools/: profiling and tuning builds)."""
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libnmpc_b200.so")
    # compile next to the target and rename: a concurrent loader (other torchrun ranks) never sees a partial file
    tmp = f"{out}.tmp.{os.getpid()}"
    cmd = ([nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])
           + ["-o", tmp, SRC])
    try:
        subprocess.check_call(cmd)
        os.replace(tmp, out)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return out
