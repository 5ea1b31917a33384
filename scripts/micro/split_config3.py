"""BASELINE config 3 (few long trials: Pd vs N, N = 1e4 .. 1e5 at 2 000 trials per hypothesis): the split path with the
float64 sums re-associated inside binades (mvd_split.cuh) against the same path adding every term in order and against
one thread per trial.
usage: python scripts/micro/split_config3.py"""
import json, sys, time
import numpy as np
sys.path.insert(0, "detecting-convolutional-codes-via-markovian-statistics_b200")
import Pd_plotter as pdp
from mvd import bitsource
from mvd.engine import Detector, Seg

g1, g2 = [[[1, 1, 1]], [[1, 0, 1]]], [[[1, 1, 0]], [[1, 0, 1]]]
det = Detector(g1, 1, 2, 2)
counts, tables = pdp._learn_edge_tables(det, [0.1], None, 200, 1.0, 12345)
det.set_models(tables)
t1, t2 = det.taps_of(g1), det.taps_of(g2)
T = bitsource.bsc_threshold(0.1)


def timed(segs, reps=5):
    det.detect(segs, seed=1, engine="fsm")
    ms = []
    for _ in range(reps):
        tal = det.detect(segs, seed=1, engine="fsm")
        ms.append(det.last_kernel_ms())
    return float(np.median(ms)), det.last_kernel_kind(), tal.tolist(), det.learn_dirty_chunks()


def point(N, trials):
    return [Seg(N=N, threshold=T, stream=0, table=0, enc_taps=t1, decide=0, trial_begin=0, trial_end=trials),
            Seg(N=N, threshold=T, stream=1, table=0, enc_taps=t2, decide=1, trial_begin=0, trial_end=trials)]


cases = [("N=1e4 x 2000 x 2", point(10000, 2000)), ("N=1e5 x 2000 x 2", point(100000, 2000)),
         ("N=1e5 x 250 x 2 (one of 8 GPUs)", point(100000, 250)),
         ("Pd vs N sweep: N = 1e4, 2e4, 5e4, 1e5 x 2000 x 2", sum((point(N, 2000) for N in (10000, 20000, 50000, 100000)), []))]
for name, segs in cases:
    steps = sum(s.N * s.ntrials for s in segs)
    out = dict(case=name, steps=steps)
    for label, setup in (("reassociated", lambda: (det.split_trials(1), det.split_sequential(False))),
                         ("term_by_term", lambda: (det.split_trials(1), det.split_sequential(True))),
                         ("one_thread_per_trial", lambda: (det.split_trials(2), det.split_sequential(False)))):
        setup()
        ms, kind, tal, dirty = timed(segs)
        out[label] = dict(kernel_ms=round(ms, 4), steps_per_s=steps / (ms * 1e-3), kind=kind, dirty_chunks=dirty)
        if kind & 16384:
            out[label]["subchunks"], out[label]["subchunks_term_by_term"] = det.split_stats()
        out.setdefault("tallies", tal)
        assert tal == out["tallies"], (label, tal, out["tallies"])
    det.split_trials(0)
    det.split_sequential(False)
    ms, kind, tal, _ = timed(segs)
    out["automatic"] = dict(kernel_ms=round(ms, 4), kind=kind)
    print(json.dumps(out), flush=True)
