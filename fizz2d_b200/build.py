"""In-tree build of libfizz_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "fizz_engine.cu")]
OUT = os.path.join(HERE, "libfizz_b200.so")


def out_path(dot_mode=1):
    """One library per numpy dot/norm contract (csrc/fizz_common.cuh FIZZ_DOT_MODE); 1 is the default build."""
    return OUT if dot_mode == 1 else os.path.join(HERE, "libfizz_b200_dot%d.so" % dot_mode)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false",                      # never contract a*b+c: the FMAs numpy performs are written explicitly
         "--shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "-I", os.path.join(ROOT, "include")]


def needs_build(dot_mode=1):
    out = out_path(dot_mode)
    if not os.path.isfile(out):
        return True
    t = os.path.getmtime(out)
    import glob
    deps = SRC + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "csrc", "*.inc")) + \
        [os.path.join(ROOT, "include", "fizz_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_extension(force=False, verbose=False, dot_mode=1):
    if dot_mode not in (0, 1, 2):
        raise ValueError("dot_mode must be 0, 1 or 2")
    out = out_path(dot_mode)
    if not force and not needs_build(dot_mode):
        return out
    extra = os.environ.get("FIZZ_NVCC_EXTRA", "").split()
    cmd = [NVCC] + FLAGS + ["-DFIZZ_DOT_MODE=%d" % dot_mode] + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SRC
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    mode = int(sys.argv[sys.argv.index("--dot-mode") + 1]) if "--dot-mode" in sys.argv else 1
    print(build_extension(force="--force" in sys.argv, verbose="-v" in sys.argv, dot_mode=mode))
