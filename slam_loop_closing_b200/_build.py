"""In-tree nvcc build of liblc_b200.so (sm_100a only; no torch JIT cache, the .so travels with the repo)."""
import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "liblc_b200.so")
SOURCES = ["lc_abi.cu"]
HEADERS = ["lc_kernels.cuh", "lc_tensor.cuh", "lc_mxf4.cuh", "lc_l2.cuh", os.path.join("..", "..", "include", "lc_abi.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: liblc_b200.so cannot be built (there is no CPU fallback)")


def have_nvcc():
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into csrc/liblc_b200.so; returns the path."""
    if force or stale():
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
              [os.path.join(CSRC, s) for s in SOURCES]
        subprocess.check_call(cmd)
    return LIB
