#!/usr/bin/env python3
"""Soak check: every detection path of libmvd gives the same tallies on the paper sweep at a few million trials per point
(two-trials-per-thread ACS, one-trial ACS, one-load NEXT walk, two-load NEXT walk, generic checked kernels, split path),
for m = 2, 3 and 4.  Prints one JSON line per code."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
import numpy as np
import Pd_plotter as pdp
from mvd import bitsource
from mvd.engine import Detector, Seg

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
P7 = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
CODES = [("m2 (7,5)/(6,5)", 2, [[[1,1,1]],[[1,0,1]]], [[[1,1,0]],[[1,0,1]]], trials, "python"),
         ("m3 demo pair", 3, [[[1,1,1,1]],[[1,0,1,1]]], [[[1,0,1,1]],[[1,1,1,1]]], trials // 2, "python"),
         ("m4 (31,33)", 4, [[[1,1,0,0,1]],[[1,1,0,1,1]]], [[[1,1,0,1,1]],[[1,1,0,0,1]]], trials // 8, "gpu")]
for name, m, g1, g2, ntr, how in CODES:
    det = Detector(g1, 1, 2, m, enumerate_with=how, max_states=1 << 20)
    _, tables = pdp._learn_edge_tables(det, P7, None, 200, 1.0, 12345)
    det.set_models(tables)
    t1, t2 = det.taps_of(g1), det.taps_of(g2)
    segs = []
    for q, p in enumerate(P7):
        T = bitsource.bsc_threshold(p)
        segs.append(Seg(N=500, threshold=T, stream=2 * q, table=q, enc_taps=t1, decide=0, trial_begin=0, trial_end=ntr))
        segs.append(Seg(N=500, threshold=T, stream=2 * q + 1, table=q, enc_taps=t2, decide=1, trial_begin=0, trial_end=ntr))
    results = {}

    def run(tag, engine, **opt):
        det.force_generic(opt.get("generic", False))
        det.no_pair(opt.get("pair", 0))
        det.no_fsm1(opt.get("no_fsm1", False))
        det.split_trials(opt.get("split", 2))
        det.no_antipodal(opt.get("no_antipodal", False))
        t0 = time.perf_counter()
        t = det.detect(segs, seed=2026, engine=engine)
        results[tag] = (t.copy(), det.last_kernel_kind(), round(det.last_kernel_ms(), 3))
        det.force_generic(False); det.no_pair(0); det.no_fsm1(False); det.split_trials(0); det.no_antipodal(False)

    run("acs pair", "acs", pair=2)
    run("acs pair general table", "acs", pair=2, no_antipodal=True)
    run("acs one trial", "acs", pair=1)
    run("next walk one load", "fsm")
    run("next walk two loads", "fsm", no_fsm1=True)
    run("generic acs", "acs", generic=True)
    run("generic next walk", "fsm", generic=True)
    run("split", "fsm", split=1)
    ref = results["acs pair"][0]
    same = {k: bool(np.array_equal(v[0], ref)) for k, v in results.items()}
    print(json.dumps(dict(code=name, trials_per_segment=ntr, segments=len(segs), steps=2 * 500 * ntr * len(P7),
                          identical=same, kinds={k: v[1] for k, v in results.items()}, kernel_ms={k: v[2] for k, v in results.items()},
                          tallies=ref.tolist())), flush=True)
    assert all(same.values()), same
    det.close()
