"""Builds libk3d.so (the C-ABI library of include/k3d.h) in-tree with nvcc for sm_100a.

    python pygeostatistics_b200/csrc/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = [os.path.join(HERE, "k3d_kernels.cu")]
HDR = [os.path.join(ROOT, "include", "k3d.h")]
OUT = os.path.join(HERE, "libk3d.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
         "-I", os.path.join(ROOT, "include"), "--cudart", "shared"]


def build(force=False, verbose=False, out=None, defines=()):
    """`out` / `defines` build a tuning variant next to the product library (K3D_LIB selects
    which one _native.py loads)."""
    out = out or OUT
    newest = max(os.path.getmtime(f) for f in SRC + HDR + [os.path.abspath(__file__)])
    if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + \
        (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SRC
    subprocess.check_call(cmd)
    return out


DEBUG_OUT = os.path.join(HERE, "libk3d_debug.so")


def build_debug(force=False):
    """The bounds-checking variant (-DK3D_DEBUG) the debug-build test runs the parity cases on."""
    return build(force=force, out=DEBUG_OUT, defines=("K3D_DEBUG",))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                out=args[0] if args else None, defines=args[1:]))
