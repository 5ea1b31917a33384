"""Learning walk at S = 150 743, one chain of 200 S steps per call: kernel time against p (hot edges at low p).  usage: python scripts/micro/learn_big_p.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
import numpy as np
from mvd import bitsource
from mvd.engine import Detector, Seg
det = Detector([[[1, 0, 0, 1, 1]], [[1, 1, 1, 0, 1]]], 1, 2, 4, enumerate_with="lib")
L = 200 * det.S
for p in (0.001, 0.01, 0.1, 0.3, 0.5):
    seg = [Seg(N=L, threshold=bitsource.bsc_threshold(p), stream=bitsource.LEARN_STREAM, enc_taps=det.dec_taps, trial_begin=0, trial_end=1)]
    ms = []
    for _ in range(3):
        c = det.learn_counts(seg, burn=200, seed=12345)
        ms.append(det.last_kernel_ms())
    c = c[0].reshape(-1)
    top = np.sort(c)[::-1]
    print(json.dumps(dict(p=p, kernel_ms=round(float(np.median(ms)), 3), edges_visited=int((c > 0).sum()), top_edge_share=round(float(top[0] / c.sum()), 4),
                          top16_share=round(float(top[:16].sum() / c.sum()), 4))), flush=True)
