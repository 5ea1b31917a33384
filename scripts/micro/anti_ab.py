#!/usr/bin/env python3
"""A/B of the pair kernel's branch-metric tables on the paper sweep: complement-label short cut (default for (7,5))
against the general table (MVD_OPT_NO_ANTIPODAL).  Prints kernel ms (median of 5) for both and checks the tallies."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
import numpy as np
import Pd_plotter as pdp
from mvd import bitsource
from mvd.engine import Detector, Seg

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
P7 = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
g1, g2 = [[[1, 1, 1]], [[1, 0, 1]]], [[[1, 1, 0]], [[1, 0, 1]]]
det = Detector(g1, 1, 2, 2)
_, tables = pdp._learn_edge_tables(det, P7, None, 200, 1.0, 12345)
det.set_models(tables)
t1, t2 = det.taps_of(g1), det.taps_of(g2)
segs = []
for q, p in enumerate(P7):
    T = bitsource.bsc_threshold(p)
    segs.append(Seg(N=500, threshold=T, stream=2 * q, table=q, enc_taps=t1, decide=0, trial_begin=0, trial_end=trials))
    segs.append(Seg(N=500, threshold=T, stream=2 * q + 1, table=q, enc_taps=t2, decide=1, trial_begin=0, trial_end=trials))
out = {}
for rnd in range(2):
    for tag, off in (("complement_label", False), ("general_table", True)):
        det.no_antipodal(off)
        det.detect(segs, seed=1, engine="acs")
        ms = []
        for _ in range(5):
            t = det.detect(segs, seed=1, engine="acs")
            ms.append(det.last_kernel_ms())
        out.setdefault(tag, []).append(round(float(np.median(ms)), 3))
        out[tag + "_tallies"] = t.tolist()
det.no_antipodal(False)
steps = 2 * 500 * trials * len(P7)
print(json.dumps(dict(steps=steps, kernel_ms=dict(complement_label=out["complement_label"], general_table=out["general_table"]),
                      steps_per_s={k: steps / (1e-3 * min(out[k])) for k in ("complement_label", "general_table")},
                      tallies_equal=out["complement_label_tallies"] == out["general_table_tallies"])))
