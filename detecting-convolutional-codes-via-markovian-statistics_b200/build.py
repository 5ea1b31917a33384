#!/usr/bin/env python3
"""Build libmvd.so in-tree for sm_100a (explicit nvcc; the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = [os.path.join(PKG, "csrc", "mvd.cu")]
DEPS = SRC + [os.path.join(PKG, "csrc", "mvd_kernels.cuh"), os.path.join(PKG, "csrc", "mvd_detect2.cuh"), os.path.join(ROOT, "include", "mvd.h")]
OUT = os.path.join(PKG, "libmvd.so")


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", OUT] + SRC
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
