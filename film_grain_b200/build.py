"""Build libfg_b200.so (the C-ABI engine) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libfg_b200.so")
SOURCES = ["fg_api.cu", os.path.join("..", "host", "film_grain.cpp"), os.path.join("..", "host", "image_io.cpp")]


def deps() -> list[str]:
    """Every file the library is compiled from: globbed, so a new header cannot be forgotten."""
    import glob
    out = []
    for pat in ("csrc/*.cu", "csrc/*.cuh", "csrc/*.h", "host/*", "../include/*.h"):
        out += glob.glob(os.path.join(HERE, pat))
    return out


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           # bit-exactness: never contract a*b+c (the reference's CPU path never fuses), IEEE div/sqrt
           "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
           "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off", "-ccbin", "/usr/bin/g++",
           "-shared", "-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lz"]  # PNG (host/image_io.cpp)
    cmd += os.environ.get("FG_NVCC_EXTRA", "").split()  # experiments only (e.g. -DFG_TILE_WARPS=32)
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
