"""Pd-vs-N sweep at the reference's 10^4 iterations (BASELINE config 3 on one GPU): kernel time of one detection call for the
automatic path choice, the split path and one thread per trial.  usage: python scripts/micro/nsweep_paths.py"""
import json, sys
import numpy as np
sys.path.insert(0, "detecting-convolutional-codes-via-markovian-statistics_b200")
import Pd_plotter as pdp
import viterbi_markov as vm
from mvd import codes
PAIRS = {"m2": (2, [[[1, 1, 1]], [[1, 0, 1]]], [[[1, 1, 0]], [[1, 0, 1]]]), "m3": (3, [[[1, 1, 1, 1]], [[1, 0, 1, 1]]], [[[1, 0, 1, 1]], [[1, 1, 1, 1]]])}
Nv = [100, 200, 500, 1000, 2000, 5000, 10000, 100000]
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
for name, (m, g1, g2) in PAIRS.items():
    det = vm._detector(codes.freeze_generator(g1), 1, 2, m, 0)
    out = dict(case=name, trials=trials)
    for label, mode in (("auto", 0), ("split", 1), ("one_thread", 2)):
        det.split_trials(mode)
        d = {}
        pdp.run_experiment(1, 2, m, g1, g2, trials, [0.05, 0.1], None, 200, 1.0, 12345, N_spectrum=Nv, details=d)
        ms = []
        for _ in range(3):
            pdp.run_experiment(1, 2, m, g1, g2, trials, [0.05, 0.1], None, 200, 1.0, 12345, N_spectrum=Nv, details=d)
            ms.append(d["detect_kernel_ms"])
        out[label] = dict(kernel_ms=round(float(np.median(ms)), 3), kind=d["kernel_kind"], sha=hash(tuple(np.asarray(d["tallies"]).tolist())) & 0xFFFF)
    det.split_trials(0)
    print(json.dumps(out), flush=True)
