"""Rate-1/3 code (n = 3) through the fast NEXT-walk kernels (run_trial_n3) against the generic kernel: kernel time per launch.
usage: python scripts/micro/rate13.py"""
import json, sys
import numpy as np
sys.path.insert(0, "detecting-convolutional-codes-via-markovian-statistics_b200")
import Pd_plotter as pdp
from mvd import bitsource
from mvd.engine import Detector, Seg
g1, g2 = [[[1, 1, 1]], [[1, 0, 1]], [[1, 1, 0]]], [[[1, 0, 1]], [[1, 1, 1]], [[0, 1, 1]]]
det = Detector(g1, 1, 3, 2)
counts, tables = pdp._learn_edge_tables(det, [0.1], None, 200, 1.0, 12345)
det.set_models(tables)
T = bitsource.bsc_threshold(0.1)
N, trials = 500, 400000
segs = [Seg(N=N, threshold=T, stream=d, table=0, enc_taps=det.taps_of((g1, g2)[d]), decide=d, trial_begin=0, trial_end=trials) for d in (0, 1)]
out = dict(case="rate 1/3, m = 2", S=det.S, N=N, trials=trials, steps=2 * N * trials)
for label, gen in (("fast_walk", False), ("generic", True)):
    det.force_generic(gen)
    det.detect(segs, seed=1, engine="fsm")
    ms = []
    for _ in range(3):
        tal = det.detect(segs, seed=1, engine="fsm")
        ms.append(det.last_kernel_ms())
    out[label] = dict(kernel_ms=round(float(np.median(ms)), 3), steps_per_s=out["steps"] / (np.median(ms) * 1e-3), kind=det.last_kernel_kind(),
                      tallies=tal.tolist())
det.force_generic(False)
assert out["fast_walk"]["tallies"] == out["generic"]["tallies"]
print(json.dumps(out))
