"""In-tree build of libfizz_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "fizz_engine.cu")]
OUT = os.path.join(HERE, "libfizz_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false",                      # never contract a*b+c: the FMAs numpy performs are written explicitly
         "--shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "-I", os.path.join(ROOT, "include")]


def needs_build():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    import glob
    deps = SRC + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + [os.path.join(ROOT, "include", "fizz_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_extension(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    extra = os.environ.get("FIZZ_NVCC_EXTRA", "").split()
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build_extension(force="--force" in sys.argv, verbose="-v" in sys.argv))
