#!/usr/bin/env python3
"""Throughput / wall time of the other BASELINE configs (not the bench line): m = 3 demo pair, m = 4,
the paper sweep at the reference's own num_iter, and the N sweep.  Prints one JSON per case."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
import numpy as np
import Pd_plotter as pdp
import viterbi_markov as vm
from mvd import bitsource, codes
from mvd.engine import Detector, Seg

def timed_detect(det, segs, engine, reps=3):
    det.detect(segs, seed=1, engine=engine)
    ms = []
    for _ in range(reps):
        det.detect(segs, seed=1, engine=engine)
        ms.append(det.last_kernel_ms())
    return float(np.median(ms)), det.last_kernel_kind()

def case_code(name, gen1, gen2, m, N, trials, ps, learn_len=None, enumerate_with="python"):
    t0 = time.perf_counter()
    det = Detector(gen1, 1, 2, m, enumerate_with=enumerate_with)
    t_enum = time.perf_counter() - t0
    t0 = time.perf_counter()
    counts, tables = pdp._learn_edge_tables(det, ps, learn_len, 200, 1.0, 12345)
    t_learn = time.perf_counter() - t0
    learn_ms = det.last_kernel_ms()
    det.set_models(tables)
    t1, t2 = det.taps_of(gen1), det.taps_of(gen2)
    segs = []
    for q, p in enumerate(ps):
        T = bitsource.bsc_threshold(p)
        segs.append(Seg(N=N, threshold=T, stream=2 * q, table=q, enc_taps=t1, decide=0, trial_begin=0, trial_end=trials))
        segs.append(Seg(N=N, threshold=T, stream=2 * q + 1, table=q, enc_taps=t2, decide=1, trial_begin=0, trial_end=trials))
    steps = 2 * N * trials * len(ps)
    out = dict(case=name, S=det.S, m=m, N=N, trials=trials, points=len(ps), steps=steps, enum_s=round(t_enum, 3),
               learn_wall_s=round(t_learn, 4), learn_kernel_ms=round(learn_ms, 3), learn_len=pdp._learn_len(det.S, learn_len))
    for eng in ("acs", "fsm"):
        try:
            ms, kind = timed_detect(det, segs, eng)
            out[eng] = dict(kernel_ms=round(ms, 3), steps_per_s=steps / (ms * 1e-3), kind=kind)
        except Exception as exc:
            out[eng] = dict(error=str(exc)[:200])
    det.close()
    print(json.dumps(out), flush=True)

P7 = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
which = sys.argv[1:] or ["m2", "m3", "m4", "paper", "nsweep"]
if "m2" in which:
    case_code("m2 (7,5)/(6,5) paper", [[[1,1,1]],[[1,0,1]]], [[[1,1,0]],[[1,0,1]]], 2, 500, 200000, P7)
if "m3" in which:
    case_code("m3 demo pair", [[[1,1,1,1]],[[1,0,1,1]]], [[[1,0,1,1]],[[1,1,1,1]]], 3, 500, 200000, [0.01, 0.05, 0.1, 0.2, 0.3])
if "m4" in which:
    case_code("m4 (31,33)", [[[1,1,0,0,1]],[[1,1,0,1,1]]], [[[1,1,0,1,1]],[[1,1,0,0,1]]], 4, 500, 100000, [0.05, 0.1], learn_len=200000, enumerate_with="lib")
if "nsweep" in which:
    for N in (100, 1000, 10000, 100000):
        case_code(f"N sweep N={N}", [[[1,1,1]],[[1,0,1]]], [[[1,1,0]],[[1,0,1]]], 2, N, max(2000, 20000000 // N), [0.1])
if "m56" in which:
    from mvd.engine import HashOnlyDetector
    for name, gen, m in (("m4 (23,35) recursion only", [[[1,0,0,1,1]],[[1,1,1,0,1]]], 4),
                         ("m5 (53,75) recursion only", [[[1,0,1,0,1,1]],[[1,1,1,1,0,1]]], 5),
                         ("m6 (133,171) recursion only", [[[1,0,1,1,0,1,1]],[[1,1,1,1,0,0,1]]], 6)):
        det = HashOnlyDetector(gen, 1, 2, m)
        taps = det.dec_taps
        N, trials = 500, 400000
        seg = Seg(N=N, threshold=bitsource.bsc_threshold(0.05), stream=1, enc_taps=taps, trial_begin=0, trial_end=trials)
        det.acs_hash(seg, seed=3, want_final=False)
        ms = []
        for _ in range(3):
            det.acs_hash(seg, seed=3, want_final=False)
            ms.append(det.last_kernel_ms())
        ms = float(np.median(ms))
        big = Seg(N=N, threshold=bitsource.bsc_threshold(0.05), stream=1, enc_taps=taps, trial_begin=0, trial_end=4 * trials)
        det.acs_final(big, seed=3)
        ms2 = []
        for _ in range(3):
            det.acs_final(big, seed=3)
            ms2.append(det.last_kernel_ms())
        ms2 = float(np.median(ms2))
        print(json.dumps(dict(case=name, m=m, N=N, trials=trials, kernel_ms=round(ms, 3), steps_per_s=N * trials / (ms * 1e-3),
                              int_ops_per_s=(5 * 2 ** m + 11) * N * trials / (ms * 1e-3),
                              pair_final_only=dict(trials=4 * trials, kernel_ms=round(ms2, 3), steps_per_s=N * 4 * trials / (ms2 * 1e-3),
                                                   int_ops_per_s=(5 * 2 ** m + 11) * N * 4 * trials / (ms2 * 1e-3)))), flush=True)
        det.close()
if "m4full" in which:
    # BASELINE config 4 at the reference's own defaults (N_SPECTRUM_BY_M[4], learn_len = 200 S, 7 p's) through the public
    # API; the reference cannot run this (dense S x S counts: 182 GB at S = 150 743)
    for name, g1, g2, it in (("m4 (31,33) S=25751", [[[1,1,0,0,1]],[[1,1,0,1,1]]], [[[1,1,0,1,1]],[[1,1,0,0,1]]], 100000),
                             ("m4 (23,35) S=150743", [[[1,0,0,1,1]],[[1,1,1,0,1]]], [[[1,1,1,0,1]],[[1,0,0,1,1]]], 100000)):
        t0 = time.perf_counter()
        d = {}
        df = pdp.run_experiment(1, 2, 4, g1, g2, it, P7, None, 200, 1.0, 12345, details=d)
        cold = time.perf_counter() - t0
        t0 = time.perf_counter()
        d = {}
        df = pdp.run_experiment(1, 2, 4, g1, g2, it, P7, None, 200, 1.0, 12345, details=d)
        wall = time.perf_counter() - t0
        det = vm._detector(codes.freeze_generator(g1), 1, 2, 4)
        print(json.dumps(dict(case=f"run_experiment defaults {name} num_iter={it}", S=d["S"], learn_len=d["learn_len"],
                              bfs=getattr(det, "bfs_stats", None) and {k: det.bfs_stats[k] for k in ("S", "ms", "iterations")},
                              cold_wall_s=round(cold, 3), wall_s=round(wall, 3), detect_kernel_ms=round(d["detect_kernel_ms"], 3),
                              learn_kernel_ms=round(d["learn_kernel_ms"], 3), wall_breakdown_s={k: round(v, 4) for k, v in d["wall_s"].items()},
                              steps=d["steps"], learn_steps=7 * d["learn_len"], steps_per_s=(d["steps"] + 7 * d["learn_len"]) / wall,
                              Pd=df["Pd"].tolist(), Pc=df["Pc"].tolist())), flush=True)
if "paper" in which:
    # python Pd_plotter.py as shipped: num_iter = 10^4 (reference default), wall time through the public API
    for it in (10000, 1000000):
        pdp.run_experiment(1, 2, 2, [[[1,1,1]],[[1,0,1]]], [[[1,1,0]],[[1,0,1]]], it, P7, None, 200, 1.0, 12345)
        t0 = time.perf_counter()
        d = {}
        df = pdp.run_experiment(1, 2, 2, [[[1,1,1]],[[1,0,1]]], [[[1,1,0]],[[1,0,1]]], it, P7, None, 200, 1.0, 12345, details=d)
        wall = time.perf_counter() - t0
        print(json.dumps(dict(case=f"paper sweep run_experiment num_iter={it}", wall_ms=round(1e3 * wall, 3), detect_kernel_ms=round(d["detect_kernel_ms"], 3),
                              steps=d["steps"], steps_per_s=d["steps"] / wall, Pd=df["Pd"].tolist())), flush=True)
