#!/usr/bin/env python3
"""Build libmvd.so in-tree for sm_100a (explicit nvcc; the .so travels with the repo snapshot).
The kernels are split over a few translation units that compile in parallel."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
UNITS = ["mvd.cu", "mvd_tu_generic_acs.cu", "mvd_tu_generic_fsm.cu", "mvd_tu_det2_acs.cu", "mvd_tu_det2_fsm.cu",
         "mvd_tu_learn.cu", "mvd_tu_bfs.cu", "mvd_tu_chernoff.cu", "mvd_tu_parity.cu", "mvd_tu_det3_pair.cu", "mvd_tu_acsp.cu"]
HEADERS = ["mvd_types.h", "mvd_launch.h", "mvd_kernels.cuh", "mvd_detect2.cuh", "mvd_learn2.cuh", "mvd_bfs.cuh", "mvd_chernoff.cuh", "mvd_parity.cuh", "mvd_detect3p.cuh", "mvd_split.cuh", "mvd_acsp.cuh"]
DEPS = [os.path.join(CSRC, f) for f in UNITS + HEADERS] + [os.path.join(ROOT, "include", "mvd.h")]
OBJ = os.path.join(PKG, "build")
OUT = os.path.join(PKG, "libmvd.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    common = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        common.insert(0, "-Xptxas=-v")

    def compile_unit(unit: str) -> str:
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        subprocess.check_call([nvcc] + common + ["-c", "-o", obj, os.path.join(CSRC, unit)])
        return obj

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_unit, UNITS))
    subprocess.check_call([nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", OUT] + objs)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
