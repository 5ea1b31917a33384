"""One config-3 point through the split path (for ncu): N = 1e5, 2 000 trials x 2; argv[1] = seq: every term added in order"""
import sys
sys.path.insert(0, "detecting-convolutional-codes-via-markovian-statistics_b200")
import Pd_plotter as pdp
from mvd import bitsource
from mvd.engine import Detector, Seg
g1, g2 = [[[1, 1, 1]], [[1, 0, 1]]], [[[1, 1, 0]], [[1, 0, 1]]]
det = Detector(g1, 1, 2, 2)
counts, tables = pdp._learn_edge_tables(det, [0.1], None, 200, 1.0, 12345)
det.set_models(tables)
T = bitsource.bsc_threshold(0.1)
segs = [Seg(N=100000, threshold=T, stream=d, table=0, enc_taps=det.taps_of((g1, g2)[d]), decide=d, trial_begin=0, trial_end=2000) for d in (0, 1)]
det.split_trials(1)
det.split_sequential(len(sys.argv) > 1 and sys.argv[1] == "seq")
for _ in range(3):
    print(det.detect(segs, seed=1, engine="fsm"), det.last_kernel_ms(), det.last_kernel_kind(), det.split_stats())
