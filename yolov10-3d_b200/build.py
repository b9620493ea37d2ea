"""Build liby3d_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.  No torch involved: the library is a
plain C-ABI shared object (include/y3d.h); nvcc cross-compiles without a GPU."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liby3d_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-shared", "-Xcompiler",
         "-fPIC", "--expt-relaxed-constexpr", "-Xcudafe", "--diag_suppress=177"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def up_to_date():
    if not os.path.exists(OUT):
        return False
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "y3d.h")]
    return all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; liby3d_b200.so cannot be built")
    cmd = [NVCC, *FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", OUT, *sources()]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
