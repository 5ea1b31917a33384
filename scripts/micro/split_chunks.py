"""Split path: kernel time per call for forced chunk sizes (256 / 512 / 1024 steps per walker thread) against the automatic choice.
usage: python scripts/micro/split_chunks.py"""
import json, sys
import numpy as np
sys.path.insert(0, "detecting-convolutional-codes-via-markovian-statistics_b200")
import Pd_plotter as pdp
from mvd import bitsource
from mvd.engine import Detector, Seg
g1, g2 = [[[1, 1, 1]], [[1, 0, 1]]], [[[1, 1, 0]], [[1, 0, 1]]]
det = Detector(g1, 1, 2, 2)
counts, tables = pdp._learn_edge_tables(det, [0.1], None, 200, 1.0, 12345)
det.set_models(tables)
T = bitsource.bsc_threshold(0.1)
det.split_trials(1)
for N, trials in ((100000, 125), (100000, 250), (100000, 500), (100000, 1250), (100000, 2000), (10000, 1250), (10000, 10000), (20000, 5000)):
    segs = [Seg(N=N, threshold=T, stream=d, table=0, enc_taps=det.taps_of((g1, g2)[d]), decide=d, trial_begin=0, trial_end=trials) for d in (0, 1)]
    out = dict(N=N, trials=trials)
    for chunk in (0, 256, 512, 1024):
        det.split_chunk(chunk)
        det.detect(segs, seed=1, engine="fsm")
        ms = []
        for _ in range(5):
            det.detect(segs, seed=1, engine="fsm")
            ms.append(det.last_kernel_ms())
        out["auto" if chunk == 0 else str(chunk)] = round(float(np.median(ms)), 4)
    print(json.dumps(out), flush=True)
