"""Split path (BASELINE config 3): warm-up length of the speculated chunk starts against repaired chunks and kernel time,
over p (m = 2 paper pair and the m = 3 demo pair).  usage: python scripts/micro/split_warm.py"""
import json, sys
import numpy as np
sys.path.insert(0, "detecting-convolutional-codes-via-markovian-statistics_b200")
import Pd_plotter as pdp
from mvd import bitsource
from mvd.engine import Detector, Seg

for name, g1, g2, m in (("m2", [[[1, 1, 1]], [[1, 0, 1]]], [[[1, 1, 0]], [[1, 0, 1]]], 2),
                        ("m3", [[[1, 1, 1, 1]], [[1, 0, 1, 1]]], [[[1, 0, 1, 1]], [[1, 1, 1, 1]]], 3)):
    det = Detector(g1, 1, 2, m)
    ps = [0.001, 0.1, 0.3, 0.5]
    counts, tables = pdp._learn_edge_tables(det, ps, None, 200, 1.0, 12345)
    det.set_models(tables)
    t1, t2 = det.taps_of(g1), det.taps_of(g2)
    det.split_trials(1)
    for q, p in enumerate(ps):
        T = bitsource.bsc_threshold(p)
        segs = [Seg(N=100000, threshold=T, stream=0, table=q, enc_taps=t1, decide=0, trial_begin=0, trial_end=2000),
                Seg(N=100000, threshold=T, stream=1, table=q, enc_taps=t2, decide=1, trial_begin=0, trial_end=2000)]
        ref = None
        for warm in (128, 96, 64, 32):
            det.learn_warm(warm)
            det.detect(segs, seed=1, engine="fsm")
            ms = []
            for _ in range(3):
                tal = det.detect(segs, seed=1, engine="fsm")
                ms.append(det.last_kernel_ms())
            ref = ref or tal.tolist()
            assert tal.tolist() == ref
            print(json.dumps(dict(code=name, p=p, warm=warm, kernel_ms=round(float(np.median(ms)), 4), kind=det.last_kernel_kind(),
                                  dirty_chunks=det.learn_dirty_chunks())), flush=True)
    det.learn_warm(128)
    det.close()
