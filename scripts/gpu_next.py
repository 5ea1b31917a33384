#!/usr/bin/env python3
"""Measurements of the SURVEY 8(f) "next" rows beside the oracle on host cores: N3 (alpha_exponent: joint-tensor
chains + Chernoff spectral radius) and N4 (parity-template baseline Monte-Carlo).  One JSON line per case.
usage: scripts/gpu_next.py [alpha] [parity] [--no-cpu]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import alpha_exponent as ae
import comp_parity as cp
import parity_eqn_check as pec
import viterbi_markov as vm

args = sys.argv[1:]
which = [a for a in args if not a.startswith("--")] or ["alpha", "parity"]
CPU = "--no-cpu" not in args


def emit(**kw):
    print(json.dumps(kw), flush=True)


if "alpha" in which:
    cases = [("(7,5) vs (6,5) m2", [[1, 1, 1], [1, 0, 1]], [[1, 1, 0], [1, 0, 1]], 2, True),
             ("demo m3 pair", [[1, 1, 1, 1], [1, 0, 1, 1]], [[1, 0, 1, 1], [1, 1, 1, 1]], 3, True),
             ("(31,33) m4", [[1, 1, 0, 0, 1], [1, 1, 0, 1, 1]], [[1, 1, 0, 1, 1], [1, 1, 0, 0, 1]], 4, False),
             ("(37,21) m4", [[1, 1, 1, 1, 1], [1, 0, 0, 0, 1]], [[1, 0, 0, 0, 1], [1, 1, 1, 1, 1]], 4, False)]
    for name, dec, enc2, m, dense in cases:
        ae.learn_transition_edges(dec, dec, m, 0.1, 1000, 100, 1)                     # warm-up: enumeration + module load
        det = vm._detector(vm.codes.freeze_generator([[g] for g in dec]), 1, 2, m)
        t0 = time.perf_counter()
        c1, table = ae.learn_transition_edges(dec, dec, m, 0.1, 300_000, 5_000, 1, trial=0)      # the reference's defaults
        learn_wall = time.perf_counter() - t0
        learn_ms = det.last_kernel_ms()
        c2, _ = ae.learn_transition_edges(enc2, dec, m, 0.1, 300_000, 5_000, 1, trial=1)
        out = dict(case="alpha " + name, K=table.S, chain_steps=305_000, learn_kernel_ms=round(learn_ms, 3),
                   learn_wall_ms=round(1e3 * learn_wall, 3), chain_steps_per_s=305_000 / (learn_ms * 1e-3))
        d = {}
        t0 = time.perf_counter()
        I, u = ae.error_exponent_from_edges(table, c1, c2, 1.0, 401, details=d)
        out.update(edges=dict(I_err=I, best_u=u, kernel_ms=round(d["kernel_ms"], 3), wall_ms=round(1e3 * (time.perf_counter() - t0), 2),
                              iters_mean=float(d["iters"].mean()), iters_max=int(d["iters"].max()),
                              matvec_per_s=float(d["iters"].sum()) / (d["kernel_ms"] * 1e-3),
                              # bytes of one product: w - bg and NEXT per edge, x gathered per edge, bgR / y per row
                              gbs=float(d["iters"].sum()) * (table.S * table.R * (8 + 4 + 8) + table.S * 24) / (d["kernel_ms"] * 1e-3) / 1e9))
        if dense:
            C1, C2 = ae.edges_to_tensor(table, c1, 1.0), ae.edges_to_tensor(table, c2, 1.0)
            d = {}
            t0 = time.perf_counter()
            Id, ud = ae.compute_error_exponent(C1, C2, 401, details=d)
            out.update(dense=dict(I_err=Id, best_u=ud, kernel_ms=round(d["kernel_ms"], 3), wall_ms=round(1e3 * (time.perf_counter() - t0), 2),
                                  iters_mean=float(d["iters"].mean())))
            if CPU:
                import ref_port
                grid = 401 if table.S <= 64 else 5
                t0 = time.perf_counter()
                Ic, uc, _ = ref_port.compute_error_exponent(C1, C2, grid)
                out.update(cpu_eigvals=dict(u_grid=grid, seconds=round(time.perf_counter() - t0, 3), I_err=Ic, best_u=uc,
                                            seconds_per_u=round((time.perf_counter() - t0) / grid, 5)))
        if CPU and m <= 3:
            import ref_port
            L = 20000 if m == 2 else 8000
            t0 = time.perf_counter()
            ref_port.learn_transition_tensor(dec, dec, m, 0.1, L, 500, 1.0, 1)
            sec = time.perf_counter() - t0
            out.update(cpu_chain=dict(steps=L + 500, seconds=round(sec, 2), steps_per_s=(L + 500) / sec, cores=1, kind="port"))
        emit(**out)

if "parity" in which:
    g1 = [[pec.parse_poly_token("7")], [pec.parse_poly_token("5")]]
    g2 = [[pec.parse_poly_token("6")], [pec.parse_poly_token("5")]]
    P7 = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
    cp.run_parity_experiment(g1, g2, 2, [200], [0.1], 0.6, 1000, 1)
    for N_list, trials in (([200], 1000), ([200], 1_000_000), ([200, 500], 1_000_000), ([100000], 20_000)):
        d = {}
        t0 = time.perf_counter()
        df = cp.run_parity_experiment(g1, g2, 2, N_list, P7, 0.6, trials, 12345, details=d)
        wall = time.perf_counter() - t0
        emit(case=f"parity (7,5) vs (6,5) N={N_list} trials={trials}", steps=d["steps"], kernel_ms=round(d["kernel_ms"], 3),
             wall_ms=round(1e3 * wall, 3), steps_per_s=d["steps"] / (d["kernel_ms"] * 1e-3), e2e_steps_per_s=d["steps"] / wall,
             Pd=df["Pd"].tolist()[:7], Pc=df["Pc"].tolist()[:7])
    if CPU:
        import ref_port
        template, _ = cp.template_from_generators(g1, 2)
        t0 = time.perf_counter()
        ntr = 150
        for trial in range(ntr):
            ref_port.parity_trial(g1, 2, template, 0.6, 200, 0.1, 12345, ref_port.PARITY_STREAM_BASE, trial)
        sec = time.perf_counter() - t0
        emit(case="parity cpu oracle (pure Python port of comp_parity.py:165-176)", trials=ntr, steps=ntr * 202, seconds=round(sec, 2),
             steps_per_s=ntr * 202 / sec, cores=1, kind="port")
