#!/usr/bin/env python3
"""make_golden.py -- freeze golden vectors by EXECUTING THE REFERENCE'S OWN FUNCTIONS.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box);
its output, tests/golden/*.json, is committed and is what pins the oracle and the CUDA path.

What is executed unmodified from /root/reference:
  viterbi_markov.{branch_output_and_next_state, build_trellis, viterbi_metric_step,
                  enumerate_markov_states_allzero, build_symbolic_T}
  Pd_plotter.{evaluate_symbolic_T, log_prob_sequence, learn_P1_empirical, run_experiment}
What is injected: ``viterbi_markov.simulate_markov_sequence`` (absent from the reference, SURVEY
F2) -- the driver in oracle/ref_port.py running on the reference's own branch/step/trellis
functions, fed by the MVD-PHILOX-2 bit source, decoder fixed to gen1.
matplotlib (not installed; imported but unused by Pd_plotter.py:58) is stubbed.

Usage:  python oracle/make_golden.py [--ref /root/reference] [--out tests/golden]
"""
from __future__ import annotations

import argparse
import itertools
import hashlib
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_port  # noqa: E402  (driver + bit source only; arithmetic comes from the reference)

CODES = {
    "c75": dict(k=1, n=2, m=2, gen=[[[1, 1, 1]], [[1, 0, 1]]]),               # Pd_plotter.py:247
    "c65": dict(k=1, n=2, m=2, gen=[[[1, 1, 0]], [[1, 0, 1]]]),               # Pd_plotter.py:248
    "m3a": dict(k=1, n=2, m=3, gen=[[[1, 1, 1, 1]], [[1, 0, 1, 1]]]),         # demo_script.py:49
    "m3b": dict(k=1, n=2, m=3, gen=[[[1, 0, 1, 1]], [[1, 1, 1, 1]]]),         # demo_script.py:50
    "r13": dict(k=1, n=3, m=2, gen=[[[1, 1, 1]], [[1, 0, 1]], [[1, 1, 0]]]),  # rate-1/3 edge case
    "m1": dict(k=1, n=2, m=1, gen=[[[1, 1]], [[1, 0]]]),                      # smallest memory
}


M4_CODES = {
    "m4a": dict(k=1, n=2, m=4, gen=[[[1, 1, 0, 0, 1]], [[1, 1, 0, 1, 1]]]),   # (31,33)
    "m4c": dict(k=1, n=2, m=4, gen=[[[1, 0, 0, 1, 1]], [[1, 1, 1, 0, 1]]]),   # (23,35)
}


# k = 2 inputs per step (viterbi_markov.py:82-132 is generic in k; every input sees the same register [u_i, s_0 ..]):
# the table-driven device path (mvd_set_code_tables) is pinned by these
K2_CODES = {
    "k2b": dict(k=2, n=3, m=2, gen=[[[1, 0, 1], [0, 1, 1]], [[1, 1, 0], [1, 0, 1]], [[0, 1, 1], [1, 1, 1]]]),             # S = 5
    "k2c": dict(k=2, n=3, m=3, gen=[[[1, 1, 0, 1], [0, 1, 1, 0]], [[1, 0, 1, 1], [1, 1, 0, 0]], [[0, 1, 1, 1], [1, 0, 1, 0]]]),   # S = 235
    "k2d": dict(k=2, n=3, m=3, gen=[[[1, 0, 1, 1], [0, 1, 1, 0]], [[1, 1, 0, 1], [1, 0, 1, 0]], [[0, 1, 1, 1], [1, 1, 0, 0]]]),
    # k = 3 inputs, n = 4 outputs, m = 2 < k (the next state is the first two inputs): S = 89 and S = 15
    "k3a": dict(k=3, n=4, m=2, gen=[[[1, 1, 1], [1, 1, 1], [0, 1, 0]], [[1, 0, 1], [1, 0, 1], [0, 0, 0]],
                                    [[1, 1, 1], [0, 0, 1], [0, 0, 1]], [[0, 1, 1], [1, 0, 1], [0, 1, 1]]]),
    "k3b": dict(k=3, n=4, m=2, gen=[[[0, 0, 1], [1, 1, 0], [0, 1, 0]], [[0, 0, 0], [0, 1, 0], [0, 1, 0]],
                                    [[0, 1, 0], [0, 1, 0], [0, 1, 0]], [[1, 1, 0], [1, 0, 0], [0, 0, 0]]]),
}


def sha16(obj) -> str:
    return hashlib.sha256(repr(obj).encode()).hexdigest()[:16]


def load_reference(path):
    sys.path.insert(0, path)
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import warnings
    warnings.simplefilter("ignore")
    import viterbi_markov as vm
    import Pd_plotter as pdp
    assert not hasattr(vm, "simulate_markov_sequence"), "reference now ships a simulator: re-survey"
    return vm, pdp


class Injected:
    """The simulator injected into the reference; tracks which (point, hyp, trial) a call is."""

    def __init__(self, vm, gen1, num_iter, trial_offset=0):
        self.vm, self.gen1, self.num_iter, self.trial_offset = vm, gen1, num_iter, trial_offset
        self.calls = 0
        self.learn_calls = 0

    def __call__(self, generator_matrix, m, k, n, length, p_val, random_input=True, seed=None):
        vm = self.vm
        kw = dict(step=vm.viterbi_metric_step, branch_fn=vm.branch_output_and_next_state,
                  trellis_fn=vm.build_trellis, decoder_matrix=self.gen1)
        if seed is not None:                                    # learn chain, Pd_plotter.py:149-155
            self.learn_calls += 1
            self.seed = seed
            return ref_port.simulate_markov_sequence(generator_matrix, m, k, n, length, p_val, random_input,
                                                     seed, stream=ref_port.LEARN_STREAM, trial=0, **kw)
        c = self.calls                                          # trial loop, Pd_plotter.py:212,219
        self.calls += 1
        point, within = divmod(c, 2 * self.num_iter)
        trial, hyp = divmod(within, 2)
        return ref_port.simulate_markov_sequence(generator_matrix, m, k, n, length, p_val, random_input,
                                                 self.seed, stream=2 * point + hyp,
                                                 trial=self.trial_offset + trial, **kw)


def kat_for_code(vm, pdp, name, spec, symbolic=True):
    k, n, m, gen = spec["k"], spec["n"], spec["m"], spec["gen"]
    states, trans, all_r = vm.enumerate_markov_states_allzero(gen, m, k, n)
    index = {s: i for i, s in enumerate(states)}
    trellis = vm.build_trellis(gen, m, k)
    S = len(states)
    nxt = [[index[vm.viterbi_metric_step(list(states[i]), trellis, r)] for r in all_r] for i in range(S)]
    out = dict(k=k, n=n, m=m, gen=gen, S=S, max_metric=max(max(s) for s in states),
               nnz=sum(len(trans[i]) for i in range(S)),
               states_sha=sha16(states), next_sha=sha16(nxt),
               all_r=[list(r) for r in all_r],
               trellis={str(ns): [[ps, list(u), list(o)] for ps, u, o in lst] for ns, lst in trellis.items()},
               branches=[[s, list(u), list(vm.branch_output_and_next_state(s, u, gen, m, k)[0]),
                          vm.branch_output_and_next_state(s, u, gen, m, k)[1]]
                         for s in range(1 << m) for u in itertools.product([0, 1], repeat=k)])
    if S <= 500:
        out["states"] = [list(s) for s in states]
        out["next"] = nxt
        out["mult"] = [[len(trans[i][nxt[i][r]]) for r in range(len(all_r))] for i in range(S)]
    # trajectory KAT: received words from a 32-bit LCG
    x = 12345
    cur = tuple([0] * (1 << m))
    traj = []
    for _ in range(10000):
        x = (1664525 * x + 1013904223) % (1 << 32)
        cur = vm.viterbi_metric_step(list(cur), trellis, all_r[x >> (32 - n)])
        traj.append(index[cur])
    out["lcg_traj_first"] = traj[:32]
    out["lcg_traj_sum"] = sum(traj)
    out["lcg_traj_sha"] = sha16(traj)
    if symbolic:
        t0 = time.time()
        p_sym, T_sym = vm.build_symbolic_T(states, trans, all_r)
        out["symbolic_seconds"] = round(time.time() - t0, 2)
        tvals = {}
        for pv in (0.5, 0.1, 0.3):
            Tn = pdp.evaluate_symbolic_T(T_sym, p_sym, pv)
            # edge form: T[i, next[i][r]]
            tvals[repr(pv)] = [[float(Tn[i, nxt[i][r]]) for r in range(len(all_r))] for i in range(S)]
            assert abs(Tn.sum(axis=1) - 1).max() < 1e-12
        out["T_edge"] = tvals
        if S <= 40:
            out["T_sym_str"] = {f"{i},{j}": str(T_sym[i, j]) for i in range(S) for j in trans[i]}
    return out


def sim_kat(vm, name, spec, enc_spec, N, p, seed, stream, trial):
    """Metric trajectory from the reference's branch + step functions under MVD-PHILOX-2."""
    k, n, m = spec["k"], spec["n"], spec["m"]
    sim = ref_port.simulate_markov_sequence(enc_spec["gen"], m, k, n, N, p, True, seed,
                                            decoder_matrix=spec["gen"], stream=stream, trial=trial,
                                            step=vm.viterbi_metric_step,
                                            branch_fn=vm.branch_output_and_next_state,
                                            trellis_fn=vm.build_trellis)
    u, e = ref_port.philox_bits(seed, stream, trial, N, n, ref_port.threshold_of(p), k)
    return dict(N=N, p=p, seed=seed, stream=stream, trial=trial, u_bits=u, e_bits=e,
                metrics=[list(d) for d in sim["metrics"]], received=[list(r) for r in sim["received"]])


def experiment_golden(vm, pdp, dec, enc2, num_iter, p_vec, N_list, seed, laplace=1.0, learn_len=None,
                      learn_burn=200):
    """Reference run_experiment + learn_P1_empirical, unmodified, with the injected simulator."""
    k, n, m = dec["k"], dec["n"], dec["m"]
    gen1, gen2 = dec["gen"], enc2["gen"]
    inj = Injected(vm, gen1, num_iter)
    vm.simulate_markov_sequence = inj
    pdp.learn_P1_empirical.cache_clear()
    saved = dict(pdp.N_SPECTRUM_BY_M)
    pdp.N_SPECTRUM_BY_M[m] = list(N_list)            # configuration dict, Pd_plotter.py:78-83
    logs = []
    orig_lps = pdp.log_prob_sequence

    def recording_lps(metrics, state_index, T):
        v = orig_lps(metrics, state_index, T)
        logs.append(v)
        return v

    pdp.log_prob_sequence = recording_lps
    orig_tqdm = pdp.tqdm
    pdp.tqdm = lambda it, **kw: it
    try:
        t0 = time.time()
        df = pdp.run_experiment(k, n, m, gen1, gen2, num_iter, p_vec, learn_len, learn_burn, laplace, seed)
        secs = time.time() - t0
        # the learned matrices, straight from the reference's (cached) learner
        states, trans, all_r = vm.enumerate_markov_states_allzero(gen1, m, k, n)
        index = {s: i for i, s in enumerate(states)}
        trellis = vm.build_trellis(gen1, m, k)
        nxt = [[index[vm.viterbi_metric_step(list(s), trellis, r)] for r in all_r] for s in states]
        P_edge = {}
        for p in p_vec:
            _, _, P = pdp.learn_P1_empirical(tuple(tuple(tuple(x) for x in row) for row in gen1),
                                             k, n, m, p, learn_len, learn_burn, laplace, seed)
            P_edge[repr(p)] = dict(
                edge=[[float(P[i, nxt[i][r]]) for r in range(len(all_r))] for i in range(len(states))],
                off_edge_min=float(P.min()), sha=hashlib.sha256(P.tobytes()).hexdigest()[:16])
    finally:
        pdp.log_prob_sequence = orig_lps
        pdp.tqdm = orig_tqdm
        pdp.N_SPECTRUM_BY_M.clear()
        pdp.N_SPECTRUM_BY_M.update(saved)
        del vm.simulate_markov_sequence
    assert len(logs) == 4 * num_iter * len(p_vec) * len(N_list)
    return dict(k=k, n=n, m=m, gen1=gen1, gen2=gen2, num_iter=num_iter, p_vec=p_vec, N_list=list(N_list),
                seed=seed, laplace=laplace, learn_len=learn_len, learn_burn=learn_burn,
                rows=df.to_dict(orient="records"), csv=df.to_csv(index=False),
                logps=logs, P1_edge=P_edge, reference_seconds=round(secs, 2),
                learn_calls=inj.learn_calls)


class _RandomFromBits:
    """Stands in for ``np.random`` inside the reference's alpha_exponent module: hands out the MVD-PHILOX-2
    bits in the order the module draws them (one randint per step, then one binomial per output bit)."""

    def __init__(self, u_bits, e_bits):
        self.u, self.e, self.t, self.j = u_bits, e_bits, -1, 0

    def seed(self, _):
        pass

    def randint(self, lo, hi):
        assert (lo, hi) == (0, 2)
        self.t += 1
        self.j = 0
        return int(self.u[self.t])

    def binomial(self, one, p):
        assert one == 1
        v = int(self.e[self.t][self.j])
        self.j += 1
        return v


class _NumpyProxy:
    def __init__(self, real, rnd):
        self._real, self.random = real, rnd

    def __getattr__(self, name):
        return getattr(self._real, name)


def load_alpha_exponent(vm, ref_path):
    """Import the reference's alpha_exponent.py.  Two names it imports from viterbi_markov do not exist
    (alpha_exponent.py:58,62): they are stubbed for the import only (never called by the functions used here)."""
    import importlib
    vm.octal_to_taps = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("absent from the reference"))
    had_sim = hasattr(vm, "simulate_markov_sequence")
    if not had_sim:
        vm.simulate_markov_sequence = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("absent"))
    try:
        ae = importlib.import_module("alpha_exponent")
    finally:
        del vm.octal_to_taps
        if not had_sim:
            del vm.simulate_markov_sequence
    # the two stale call sites (:109, :116) adapted to the shipped signatures (k = 1, n = number of tap lists)
    ae.enumerate_markov_states_allzero = lambda taps, m: vm.enumerate_markov_states_allzero([[list(g)] for g in taps], m, 1, len(taps))
    ae.build_trellis = lambda taps, m: vm.build_trellis([[list(g)] for g in taps], m, 1)
    return ae


def alpha_golden(vm, ae, dec, enc1, enc2, m, p, length, burn_in, laplace, seed, u_grid):
    """learn_transition_tensor + compute_error_exponent of the reference, run on MVD-PHILOX-2 bits."""
    import numpy as real_np
    n = len(dec)
    out = dict(dec=dec, enc1=enc1, enc2=enc2, m=m, p=p, length=length, burn_in=burn_in, laplace=laplace, seed=seed,
               u_grid=u_grid, tensors=[])
    Cs = []
    for hyp, enc in enumerate((enc1, enc2)):
        trial = hyp
        u, e = ref_port.philox_bits(seed, ref_port.ALPHA_STREAM, trial, burn_in + length, n, ref_port.threshold_of(p))
        ae.np = _NumpyProxy(real_np, _RandomFromBits(u, e))
        try:
            t0 = time.time()
            C, states, sidx, all_r = ae.learn_transition_tensor(enc, dec, m, p, length=length, burn_in=burn_in,
                                                                laplace=laplace, seed=seed)
            secs = time.time() - t0
        finally:
            ae.np = real_np
        K, R = len(states), len(all_r)
        trellis = vm.build_trellis([[list(g)] for g in dec], m, 1)
        nxt = [[sidx[vm.viterbi_metric_step(list(s), trellis, r)] for r in all_r] for s in states]
        edge = [[float(C[i, nxt[i][r], r]) for r in range(R)] for i in range(K)]
        mask = real_np.ones((K, K, R), dtype=bool)
        for i in range(K):
            for r in range(R):
                mask[i, nxt[i][r], r] = False
        bg = [float(C[i][mask[i]].max()) for i in range(K)]
        assert all(float(C[i][mask[i]].min()) == bg[i] for i in range(K))
        out["tensors"].append(dict(trial=trial, K=K, R=R, edge=edge, background=bg,
                                   sha=hashlib.sha256(real_np.ascontiguousarray(C).tobytes()).hexdigest()[:16],
                                   reference_seconds=round(secs, 2)))
        Cs.append(C)
    t0 = time.time()
    I_err, best_u = ae.compute_error_exponent(Cs[0], Cs[1], u_grid=u_grid)
    out["I_err"], out["best_u"] = I_err, best_u
    out["exponent_seconds"] = round(time.time() - t0, 2)
    P1 = real_np.clip(Cs[0], 1e-300, 1.0)
    P2 = real_np.clip(Cs[1], 1e-300, 1.0)
    out["rho"] = {repr(u): ae.spectral_radius(real_np.sum((P1 ** u) * (P2 ** (1.0 - u)), axis=2))
                  for u in (0.0, 0.25, 0.5, 0.75, 1.0)}
    return out


def parity_golden(ref_path):
    """parity_eqn_check.py + comp_parity.py of the reference, executed unmodified: parity-check bases and
    per-trial parity_detector outputs on MVD-PHILOX-2 bits (the reference draws from Python's `random`)."""
    import importlib
    pec = importlib.import_module("parity_eqn_check")
    cp = importlib.import_module("comp_parity")
    out = dict(bases={}, trials={})
    for name, toks, deg in (("c75", ("7", "5"), 5), ("m3", ("17", "13"), 6), ("m4", ("23", "35"), 7), ("r13", ("7", "5", "6"), 4)):
        gens = [[pec.parse_poly_token(t)] for t in toks]
        A = pec.build_parity_system(gens, deg)
        basis = pec.nullspace_mod2(A)
        eqs = [pec.parity_vector_to_equation([row[j * (deg + 1):(j + 1) * (deg + 1)].tolist() for j in range(len(gens))])
               for row in basis]
        out["bases"][name] = dict(tokens=list(toks), deg_h=deg, gens=gens, A=A.tolist(), basis=basis.tolist(), equations=eqs)
    cases = {
        "main": dict(h1=("7", "5"), h2=("6", "5"), m=2, deg=5, N=200, p=0.1, gamma=0.6, seed=12345, ntr=40),
        "short": dict(h1=("7", "5"), h2=("6", "5"), m=2, deg=5, N=31, p=0.3, gamma=0.55, seed=7, ntr=40),
        "m3": dict(h1=("17", "13"), h2=("13", "17"), m=3, deg=6, N=130, p=0.05, gamma=0.7, seed=99, ntr=30),
        "r13": dict(h1=("7", "5", "6"), h2=("7", "6", "5"), m=2, deg=4, N=97, p=0.2, gamma=0.5, seed=3, ntr=30),
    }
    for name, c in cases.items():
        g1 = [[pec.parse_poly_token(t)] for t in c["h1"]]
        g2 = [[pec.parse_poly_token(t)] for t in c["h2"]]
        n, m, deg = len(g1), c["m"], c["deg"]
        basis = pec.nullspace_mod2(pec.build_parity_system(g1, deg))
        h_vec = [basis[0][j * (deg + 1):(j + 1) * (deg + 1)].tolist() for j in range(n)]       # comp_parity.py:148-150
        template = [(j, s) for j, poly in enumerate(h_vec) for s, bit in enumerate(poly) if bit]   # :156-160
        rec = dict(c, gens1=g1, gens2=g2, template=[list(t) for t in template], hyp=[])
        for h, gens in enumerate((g1, g2)):
            stream = ref_port.PARITY_STREAM_BASE + h
            rows = []
            for trial in range(c["ntr"]):
                u, e = ref_port.philox_bits(c["seed"], stream, trial, c["N"] + m, n, ref_port.threshold_of(c["p"]))
                v = cp.encode_convolutional(u[:c["N"]], gens, m)
                y = [[v[j][t] ^ e[t][j] for t in range(c["N"] + m)] for j in range(n)]
                decision, p_hat = cp.parity_detector(y, template, c["gamma"])
                rows.append([bool(decision), p_hat])
            rec["hyp"].append(dict(stream=stream, rows=rows))
        out["trials"][name] = rec
    return out


# The reference's run_experiment, unmodified, on the injected simulator.  c75_c65_n1000 is BASELINE config 1
# (demo_script.py predefined (2,1,2) pair at N = 1000, the demo's p grid, reduced Monte-Carlo trials).
EXPERIMENTS = {
    "c75_c65_small": lambda vm, pdp: experiment_golden(vm, pdp, CODES["c75"], CODES["c65"], num_iter=40,
                                                       p_vec=[0.01, 0.05, 0.1, 0.2, 0.3], N_list=[100, 500], seed=123),
    "c75_c65_lap": lambda vm, pdp: experiment_golden(vm, pdp, CODES["c75"], CODES["c65"], num_iter=30, p_vec=[0.1, 0.5],
                                                     N_list=[64], seed=12345, laplace=0.1, learn_len=3000, learn_burn=10),
    "m3_small": lambda vm, pdp: experiment_golden(vm, pdp, CODES["m3a"], CODES["m3b"], num_iter=12, p_vec=[0.05, 0.2],
                                                  N_list=[200], seed=123),
    "c75_c65_n1000": lambda vm, pdp: experiment_golden(vm, pdp, CODES["c75"], CODES["c65"], num_iter=24,
                                                       p_vec=[0.01, 0.05, 0.1, 0.2, 0.3], N_list=[1000], seed=123),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "tests", "golden"))
    ap.add_argument("--skip-m3-symbolic", action="store_true")
    ap.add_argument("--m4-only", action="store_true", help="only (re)write m4_kats.json")
    ap.add_argument("--alpha-only", action="store_true", help="only (re)write alpha_kats.json")
    ap.add_argument("--parity-only", action="store_true", help="only (re)write parity_kats.json")
    ap.add_argument("--k2-only", action="store_true", help="only (re)write k2_kats.json (k = 2 codes)")
    ap.add_argument("--add-experiments", action="store_true",
                    help="only add the experiments missing from experiments.json (existing entries are kept byte for byte)")
    args = ap.parse_args()
    vm, pdp = load_reference(args.ref)
    os.makedirs(args.out, exist_ok=True)

    if args.add_experiments:
        path = os.path.join(args.out, "experiments.json")
        with open(path) as f:
            exps = json.load(f)
        for name, make in EXPERIMENTS.items():
            if name not in exps:
                exps[name] = make(vm, pdp)
                print("[exp]", name, exps[name]["reference_seconds"], "s", flush=True)
        with open(path, "w") as f:
            json.dump(exps, f, separators=(",", ":"))
        return

    if args.k2_only:
        k2 = dict(codes={}, sims={}, experiments={})
        for name, spec in K2_CODES.items():
            k2["codes"][name] = kat_for_code(vm, pdp, name, spec, symbolic=(name in ("k2b", "k3b")))
            print(f"[k2] {name}: S={k2['codes'][name]['S']}", flush=True)
        k2["sims"] = {
            "k2c_self": sim_kat(vm, "k2c", K2_CODES["k2c"], K2_CODES["k2c"], 300, 0.1, 12345, 0, 7),
            "k2c_vs_d": sim_kat(vm, "k2c", K2_CODES["k2c"], K2_CODES["k2d"], 260, 0.05, 99, 1, (1 << 33) + 5),
            "k2b_self": sim_kat(vm, "k2b", K2_CODES["k2b"], K2_CODES["k2b"], 129, 0.3, 5, 2, 3),
            "k3a_self": sim_kat(vm, "k3a", K2_CODES["k3a"], K2_CODES["k3a"], 200, 0.1, 77, 0, 1),
            "k3a_vs_b": sim_kat(vm, "k3a", K2_CODES["k3a"], K2_CODES["k3b"], 161, 0.2, 77, 1, 2),
        }
        k2["experiments"]["k2c_k2d"] = experiment_golden(vm, pdp, K2_CODES["k2c"], K2_CODES["k2d"], num_iter=10, p_vec=[0.05, 0.2],
                                                         N_list=[64, 150], seed=7)
        print("[k2] experiment", k2["experiments"]["k2c_k2d"]["reference_seconds"], "s", flush=True)
        with open(os.path.join(args.out, "k2_kats.json"), "w") as f:
            json.dump(k2, f, separators=(",", ":"))
        return

    with open(os.path.join(args.out, "parity_kats.json"), "w") as f:
        json.dump(parity_golden(args.ref), f, separators=(",", ":"))
    print("[parity] done", flush=True)
    if args.parity_only:
        return

    # alpha_exponent.py (Eq. 7): the reference's learn_transition_tensor / compute_error_exponent / fit_error_exponent
    if not args.m4_only:
        ae = load_alpha_exponent(vm, args.ref)
        alpha = {}
        alpha["c75_vs_c65"] = alpha_golden(vm, ae, [[1, 1, 1], [1, 0, 1]], [[1, 1, 1], [1, 0, 1]], [[1, 1, 0], [1, 0, 1]],
                                           2, 0.1, 20000, 1000, 1.0, 5, 101)
        print("[alpha] c75_vs_c65", alpha["c75_vs_c65"]["I_err"], alpha["c75_vs_c65"]["best_u"], flush=True)
        alpha["c75_lap"] = alpha_golden(vm, ae, [[1, 1, 1], [1, 0, 1]], [[1, 1, 1], [1, 0, 1]], [[1, 0, 1], [1, 1, 1]],
                                        2, 0.05, 4000, 100, 0.25, 77, 41)
        print("[alpha] c75_lap", alpha["c75_lap"]["I_err"], flush=True)
        alpha["m3"] = alpha_golden(vm, ae, [[1, 1, 1, 1], [1, 0, 1, 1]], [[1, 1, 1, 1], [1, 0, 1, 1]],
                                   [[1, 0, 1, 1], [1, 1, 1, 1]], 3, 0.05, 30000, 1000, 1.0, 9, 11)
        print("[alpha] m3", alpha["m3"]["I_err"], alpha["m3"]["exponent_seconds"], "s", flush=True)
        Nv = [50, 100, 150, 200, 250, 300, 400]
        Pe = [0.31, 0.12, 0.05, 0.021, 0.0085, 0.0036, 0.0]
        alpha["fit"] = dict(N=Nv, Pe=Pe, result=list(ae.fit_error_exponent(Nv, Pe)),
                            few=list(ae.fit_error_exponent([10, 20], [0.1, 0.05])))
        alpha["encoder_steps"] = [[s, u, list(ae._encoder_step(s, u, [[1, 1, 0, 1], [1, 0, 1, 1]], 3)[0]),
                                   ae._encoder_step(s, u, [[1, 1, 0, 1], [1, 0, 1, 1]], 3)[1]] for s in range(8) for u in (0, 1)]
        with open(os.path.join(args.out, "alpha_kats.json"), "w") as f:
            json.dump(alpha, f, separators=(",", ":"))
        if args.alpha_only:
            return

    # m = 4: the reference's own BFS (5 s and 28 s of CPython); hashes of the state list / NEXT table and
    # the LCG trajectory pin the oracle's and the GPU's enumeration at the largest fully checkable memory
    m4 = {}
    for name, spec in (M4_CODES.items() if not args.alpha_only else []):
        t0 = time.time()
        m4[name] = kat_for_code(vm, pdp, name, spec, symbolic=False)
        m4[name]["reference_bfs_seconds"] = round(time.time() - t0, 1)
        print(f"[kat] {name}: S={m4[name]['S']} ({time.time() - t0:.1f}s)", flush=True)
    with open(os.path.join(args.out, "m4_kats.json"), "w") as f:
        json.dump(m4, f, separators=(",", ":"))
    if args.m4_only:
        return

    kats = {}
    for name, spec in CODES.items():
        sym = not (spec["m"] >= 3 and args.skip_m3_symbolic)
        if name == "m3b":
            sym = False                                   # 50 s of sympy for a mirrored pair: skip
        t0 = time.time()
        kats[name] = kat_for_code(vm, pdp, name, spec, symbolic=sym)
        print(f"[kat] {name}: S={kats[name]['S']} ({time.time() - t0:.1f}s)", flush=True)
    # step KATs quoted in SURVEY section 4 for (7,5)
    tr = vm.build_trellis(CODES["c75"]["gen"], 2, 1)
    kats["c75"]["step_kats"] = [[list(d), list(r), list(vm.viterbi_metric_step(list(d), tr, r))] for d, r in
                                [((0, 0, 0, 0), (0, 0)), ((0, 0, 0, 0), (0, 1)), ((0, 0, 1, 1), (1, 1)),
                                 ((2, 0, 1, 1), (1, 0)), ((0, 3, 2, 1), (0, 1))]]
    with open(os.path.join(args.out, "code_kats.json"), "w") as f:
        json.dump(kats, f, separators=(",", ":"))

    sims = {
        "c75_self": sim_kat(vm, "c75", CODES["c75"], CODES["c75"], 300, 0.1, 12345, 0, 7),
        "c75_vs65": sim_kat(vm, "c75", CODES["c75"], CODES["c65"], 300, 0.1, 12345, 1, 7),
        "c75_p001": sim_kat(vm, "c75", CODES["c75"], CODES["c75"], 200, 0.001, 99, 4, 0),
        "c75_p05": sim_kat(vm, "c75", CODES["c75"], CODES["c75"], 129, 0.5, 1 << 40, 6, (1 << 33) + 5),
        "m3_self": sim_kat(vm, "m3a", CODES["m3a"], CODES["m3a"], 260, 0.05, 123, 2, 1),
        "m3_vs": sim_kat(vm, "m3a", CODES["m3a"], CODES["m3b"], 260, 0.2, 123, 3, 1),
        "r13_self": sim_kat(vm, "r13", CODES["r13"], CODES["r13"], 100, 0.15, 5, 0, 3),
        "m1_self": sim_kat(vm, "m1", CODES["m1"], CODES["m1"], 33, 0.25, 5, 0, 3),
    }
    with open(os.path.join(args.out, "sim_kats.json"), "w") as f:
        json.dump(sims, f, separators=(",", ":"))
    print("[sim] done", flush=True)

    exps = {}
    for name, make in EXPERIMENTS.items():
        exps[name] = make(vm, pdp)
        print("[exp]", name, exps[name]["reference_seconds"], "s", flush=True)
    with open(os.path.join(args.out, "experiments.json"), "w") as f:
        json.dump(exps, f, separators=(",", ":"))
    print("golden vectors written to", os.path.abspath(args.out))


if __name__ == "__main__":
    main()
