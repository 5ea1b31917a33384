#!/usr/bin/env python3
"""GPU state enumeration (mvd_enumerate_states_gpu): time and size per code, beside the
host C++ BFS of libmvd and the numpy BFS of mvd/codes.py.  One JSON line per case.
usage: scripts/gpu_bfs.py [small] [m4] [m5 LOG2_BUDGET] [m6 LOG2_BUDGET]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
import numpy as np
from mvd import codes
from mvd.engine import Detector, HashOnlyDetector


def octal_taps(o, m):
    """Octal generator -> D^0-first tap list of length m + 1 (binary digits, MSB first)."""
    b = bin(int(str(o), 8))[2:]
    assert len(b) == m + 1, (o, m)
    return [int(c) for c in b]


def gen_of(pair, m):
    return [[octal_taps(pair[0], m)], [octal_taps(pair[1], m)]]


def run(name, gen, m, budget, count_only=False, host=True, chunk=0):
    cls = HashOnlyDetector
    with cls(gen, 1, 2, m) as det:
        det._enumerate_gpu(min(budget, 1 << 12), install=False, count_only=True, allow_partial=True)   # warm-up (module load)
        t0 = time.perf_counter()
        st = det._enumerate_gpu(budget, install=False, count_only=count_only, allow_partial=True, chunk_parents=chunk)
        wall = time.perf_counter() - t0
        KW = max(1, (1 << m) // 8)
        out = dict(case=name, m=m, budget=budget, count_only=count_only, **st, wall_s=round(wall, 4),
                   candidates_per_s=st["candidates"] / (st["ms"] * 1e-3) if st["ms"] else None,
                   key_bytes=4 * KW,
                   # bytes a candidate must move at 32-byte sector granularity: parent key (shared by R
                   # candidates), candidate key write, one slot sector, one key sector to compare, cslot + NEXT writes
                   sector_gbs=st["candidates"] * (4 * KW / 4 + 4 * KW + 32 + 32 + 4 + (0 if count_only else 4)) / (st["ms"] * 1e-3) / 1e9 if st["ms"] else None)
        if host and st["closed"]:
            import ctypes as C
            S = C.c_uint32()
            t0 = time.perf_counter()
            det._ck(det.lib.mvd_enumerate_states(det.ctx, budget, C.byref(S)))
            out["host_cpp_bfs_s"] = round(time.perf_counter() - t0, 4)
            assert S.value == st["S"]
            if st["S"] <= 30000:
                t0 = time.perf_counter()
                tab = codes.enumerate_states(codes.freeze_generator(gen), m, 1, 2)
                out["host_numpy_bfs_s"] = round(time.perf_counter() - t0, 4)
                assert tab.S == st["S"]
    print(json.dumps(out), flush=True)


args = sys.argv[1:] or ["small", "m4"]
if "small" in args:
    run("(7,5) m2", [[[1, 1, 1]], [[1, 0, 1]]], 2, 1 << 12)
    run("demo m3 (17,13)", [[[1, 1, 1, 1]], [[1, 0, 1, 1]]], 3, 1 << 12)
    run("(15,13) m3", gen_of((15, 13), 3), 3, 1 << 12)
if "m4" in args:
    for pair in ((23, 33), (31, 33), (23, 35), (25, 37), (31, 27), (37, 21)):
        run(f"{pair} m4", gen_of(pair, 4), 4, 1 << 19)
for key, pair, m in (("m5", (53, 75), 5), ("m6", (133, 171), 6)):
    if key in args:
        lg = int(args[args.index(key) + 1])
        run(f"{pair} m{m}", gen_of(pair, m), m, 1 << lg, count_only=True, host=False)
