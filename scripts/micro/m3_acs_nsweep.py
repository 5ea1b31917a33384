"""m = 3 demo pair, ACS engine, Pd-vs-N sweep at few trials per point: the pair kernel (blocks of 768 threads) against the one-trial
hash kernel.  usage: python scripts/micro/m3_acs_nsweep.py [trials]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
import numpy as np
import Pd_plotter as pdp
import viterbi_markov as vm
from mvd import codes
g1, g2 = [[[1, 1, 1, 1]], [[1, 0, 1, 1]]], [[[1, 0, 1, 1]], [[1, 1, 1, 1]]]
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
Ns = [100, 200, 500, 1000, 2000, 5000, 10000, 20000, 50000, 100000]
det = vm._detector(codes.freeze_generator(g1), 1, 2, 3)
ref = None
for label, mode in (("dispatcher", 0), ("one trial per thread", 1), ("pair kernel forced", 2)):
    det.no_pair(mode)
    ms = []
    for _ in range(3):
        d = {}
        df = pdp.run_experiment(1, 2, 3, g1, g2, trials, [0.05, 0.1], None, 200, 1.0, 12345, N_spectrum=Ns, engine="acs", details=d)
        ms.append(d["detect_kernel_ms"])
    t = np.asarray(d["tallies"]).tolist()
    ref = ref or t
    assert t == ref
    print(json.dumps(dict(mode=label, trials=trials, kernel_ms=round(float(np.median(ms)), 3), kind=d["kernel_kind"], steps=d["steps"])), flush=True)
det.no_pair(0)
