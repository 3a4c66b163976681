"""In-tree nvcc build of liblc_b200.so (sm_100a only; no torch JIT cache, the .so travels with the repo)."""
import fcntl
import hashlib
import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "liblc_b200.so")
STAMP = LIB + ".srchash"  # hash of the sources the library was built from (file times do not survive every copy of the tree)
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


def source_hash():
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def stale():
    """True when the library is missing or was built from other sources than the tree holds now."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into csrc/liblc_b200.so; returns the path.  Safe against several processes (the ranks
    of a torchrun launch) finding the library stale at once: one builds, the others wait and find it fresh."""
    if not (force or stale()):
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or stale():
                tmp = LIB + f".tmp{os.getpid()}"
                cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + \
                      [os.path.join(CSRC, s) for s in SOURCES]
                subprocess.check_call(cmd)
                os.replace(tmp, LIB)  # never a half-written library under the final name
                with open(STAMP, "w") as fh:
                    fh.write(source_hash() + "\n")
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB
