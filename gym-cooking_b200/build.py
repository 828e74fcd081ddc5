"""Builds libgymcook.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
# GC_LIBGYMCOOK=<path>: load / build another copy of the library (kernel experiments: scripts/build_variants.sh)
LIB_PATH = os.environ.get("GC_LIBGYMCOOK") or os.path.join(PKG_DIR, "libgymcook.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "--threads", "0",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    root = os.path.dirname(PKG_DIR)
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + [
        os.path.join(root, "include", "gymcook.h")]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    if os.environ.get("GC_LIBGYMCOOK"):
        return False  # an explicitly named copy (a kernel variant) is used as it is
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False, extra=()):
    """Compile every .cu under csrc/ into gym-cooking_b200/libgymcook.so."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = list(extra) + os.environ.get("GC_NVCC_EXTRA", "").split()  # e.g. -DGC_LUT_MIN_CTAS=6 for experiments
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB_PATH] + sources()
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libgymcook.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    build(force=True, verbose=True, extra=["-Xptxas", "-v"] if "-v" in sys.argv else [])
