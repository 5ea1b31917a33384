"""Learning chains at config-4 scale: S = 150 743 ((23,35)), learn_len = 200 S per p, 7 p's -- kernel time of mvd_learn_counts.
usage: python scripts/micro/learn_big.py [warm]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
import numpy as np
import Pd_plotter as pdp
from mvd import bitsource
from mvd.engine import Detector, Seg
P7 = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
for name, g in (("(31,33)", [[[1, 1, 0, 0, 1]], [[1, 1, 0, 1, 1]]]), ("(23,35)", [[[1, 0, 0, 1, 1]], [[1, 1, 1, 0, 1]]])):
    det = Detector(g, 1, 2, 4, enumerate_with="lib")
    if len(sys.argv) > 1:
        det.learn_warm(int(sys.argv[1]))
    L = 200 * det.S
    segs = [Seg(N=L, threshold=bitsource.bsc_threshold(p), stream=bitsource.LEARN_STREAM, enc_taps=det.dec_taps, trial_begin=0, trial_end=1) for p in P7]
    ref = None
    for it in range(3):
        t0 = time.perf_counter()
        c = det.learn_counts(segs, burn=200, seed=12345)
        w = time.perf_counter() - t0
        ref = c if ref is None else ref
        assert np.array_equal(c, ref)
        print(json.dumps(dict(code=name, S=det.S, learn_len=L, steps=7 * L, kernel_ms=round(det.last_kernel_ms(), 3), wall_ms=round(1e3 * w, 3),
                              dirty=det.learn_dirty_chunks(), sha=hex(int(np.bitwise_xor.reduce(c.reshape(-1) * np.arange(1, c.size + 1, dtype=np.uint64))))[:14])), flush=True)
    det.close()
