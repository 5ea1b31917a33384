#!/usr/bin/env python3
"""Debug aid: m = 3 / m = 4 pair kernel (ACS, two trials per thread) against the NEXT-walk engine over a ladder of N."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
import numpy as np
import Pd_plotter as pdp
from mvd import bitsource
from mvd.engine import Detector, Seg

def run(gen1, gen2, m, learn_len=None, enum="python"):
    det = Detector(gen1, 1, 2, m, enumerate_with=enum)
    counts, tables = pdp._learn_edge_tables(det, [0.05], learn_len, 200, 1.0, 12345)
    det.set_models(tables)
    t1, t2 = det.taps_of(gen1), det.taps_of(gen2)
    T = bitsource.bsc_threshold(0.05)
    for N in (1, 2, 3, 8, 9, 31, 32, 33, 64, 100, 128, 129, 333):
        segs = [Seg(N=N, threshold=T, stream=3, table=0, enc_taps=t2, decide=0, trial_begin=17, trial_end=17 + 3000)]
        det.no_pair(1)
        _, ref = det.detect(segs, seed=7, engine="fsm", want_logp=True)
        for anti in (1, 0):
            det.no_pair(2)
            det.no_antipodal(not anti)
            _, lp = det.detect(segs, seed=7, engine="acs", want_logp=True)
            kind = det.last_kernel_kind()
            det.no_antipodal(False)
            bad = np.flatnonzero((lp != ref).any(axis=1))
            print(f"m={m} N={N} anti={anti} kind={kind} bad={len(bad)} first={bad[:4].tolist()} "
                  f"{lp[bad[0]].tolist() if len(bad) else ''} {ref[bad[0]].tolist() if len(bad) else ''}", flush=True)
    det.close()

run([[[1,1,1,1]],[[1,0,1,1]]], [[[1,0,1,1]],[[1,1,1,1]]], 3)
if "m4" in sys.argv:
    run([[[1,1,0,0,1]],[[1,1,0,1,1]]], [[[1,1,0,1,1]],[[1,1,0,0,1]]], 4, learn_len=200000, enum="lib")
