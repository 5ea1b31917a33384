"""CPU tests: pin the oracle (C restatement oracle/mvd_oracle.c and Python port oracle/ref_port.py)
against the golden vectors that oracle/make_golden.py froze by executing the reference's own
functions (viterbi_markov.py / Pd_plotter.py, unmodified).  No GPU, no /root/reference needed."""
import hashlib
import math

import numpy as np
import pytest


def sha16(obj):
    return hashlib.sha256(repr(obj).encode()).hexdigest()[:16]


def _taps(spec):
    # bit t of taps[j] = generator_matrix[j][0][t]  (k = 1)
    return [sum(int(b) << t for t, b in enumerate(row[0])) for row in spec["gen"]]


SMALL = ["c75", "c65", "m3a", "m3b", "r13", "m1"]


# ------------------------------------------------------------------ trellis / step / enumeration
@pytest.mark.parametrize("name", SMALL)
def test_c_oracle_trellis_and_branches(golden, name):
    """mvdo_build_trellis / mvdo_encoder_branch == build_trellis / branch_output_and_next_state
    (viterbi_markov.py:118-132, :82-106)."""
    import c_oracle as co
    g = golden["code_kats"][name]
    n, m = g["n"], g["m"]
    prev, blab = co.build_trellis(_taps(g), n, m)
    for ns in range(1 << m):
        want = g["trellis"][str(ns)]
        assert len(want) == 2
        for b, (ps, u, out) in enumerate(want):
            assert int(prev[ns, b]) == ps
            assert int(blab[ns, b]) == sum(o << (n - 1 - j) for j, o in enumerate(out))
            assert u == [ns & 1]


@pytest.mark.parametrize("name", SMALL)
def test_ref_port_trellis_and_branches(golden, name):
    import ref_port
    g = golden["code_kats"][name]
    tr = ref_port.trellis_of(g["gen"], g["m"], g["k"])
    assert {str(ns): [[ps, list(u), list(o)] for ps, u, o in lst] for ns, lst in tr.items()} == g["trellis"]
    for s, u, out, ns in g["branches"]:
        assert ref_port.branch(s, tuple(u), g["gen"], g["m"], g["k"]) == (tuple(out), ns)


def test_step_kats(golden):
    """Eq. 4-5 known answers for (7,5) (viterbi_markov.py:139-159)."""
    import c_oracle as co
    import ref_port
    g = golden["code_kats"]["c75"]
    prev, blab = co.build_trellis(_taps(g), 2, 2)
    tr = ref_port.trellis_of(g["gen"], 2, 1)
    assert len(g["step_kats"]) == 5
    for d, r, want in g["step_kats"]:
        assert co.metric_step(prev, blab, 2, d, 2 * r[0] + r[1]).tolist() == want
        assert list(ref_port.metric_step(d, tr, tuple(r))) == want


@pytest.mark.parametrize("name", SMALL)
def test_c_oracle_enumeration(golden, name):
    """State set, BFS index order and NEXT table == enumerate_markov_states_allzero
    (viterbi_markov.py:166-195)."""
    import c_oracle as co
    g = golden["code_kats"][name]
    met, nxt = co.enumerate_states(_taps(g), g["n"], g["m"], max_states=4096)
    assert met.shape[0] == g["S"]
    assert int(met.max()) == g["max_metric"]
    assert sha16([tuple(int(v) for v in row) for row in met]) == g["states_sha"]
    assert sha16([[int(v) for v in row] for row in nxt]) == g["next_sha"]
    assert met.tolist() == g["states"]
    assert nxt.tolist() == g["next"]
    # number of distinct (i, j) pairs
    assert sum(len(set(row)) for row in nxt.tolist()) == g["nnz"]


@pytest.mark.parametrize("name", ["m4a", "m4c"])
def test_c_oracle_enumeration_m4(golden, name):
    """m = 4 (S = 25 751 / 150 743), the largest memory the reference's own BFS finishes: the oracle's
    state list, index order, NEXT table and LCG trajectory == the reference's (hashes in m4_kats.json)."""
    import c_oracle as co
    g = golden["m4_kats"][name]
    met, nxt = co.enumerate_states(_taps(g), g["n"], g["m"], max_states=1 << 18)
    assert met.shape[0] == g["S"] and int(met.max()) == g["max_metric"]
    assert sha16([tuple(r) for r in met.tolist()]) == g["states_sha"]
    assert sha16(nxt.tolist()) == g["next_sha"]
    x, cur, traj = 12345, 0, []
    for _ in range(10000):
        x = (1664525 * x + 1013904223) % (1 << 32)
        cur = int(nxt[cur, x >> 30])
        traj.append(cur)
    assert traj[:32] == g["lcg_traj_first"] and sum(traj) == g["lcg_traj_sum"] and sha16(traj) == g["lcg_traj_sha"]


@pytest.mark.parametrize("name", ["c75", "c65", "m1", "r13"])
def test_ref_port_enumeration(golden, name):
    import ref_port
    g = golden["code_kats"][name]
    states, trans, all_r = ref_port.enumerate_states(g["gen"], g["m"], g["k"], g["n"])
    assert [list(s) for s in states] == g["states"]
    assert [list(r) for r in all_r] == g["all_r"]
    assert [[len(trans[i][g["next"][i][r]]) for r in range(len(all_r))] for i in range(g["S"])] == g["mult"]


@pytest.mark.parametrize("name", SMALL)
def test_lcg_trajectory(golden, name):
    """10 000-step trajectory driven by the LCG of SURVEY section 4, via the oracle's NEXT table."""
    import c_oracle as co
    g = golden["code_kats"][name]
    n = g["n"]
    _, nxt = co.enumerate_states(_taps(g), n, g["m"], max_states=4096)
    x, cur, traj = 12345, 0, []
    for _ in range(10000):
        x = (1664525 * x + 1013904223) % (1 << 32)
        cur = int(nxt[cur, x >> (32 - n)])
        traj.append(cur)
    assert traj[:32] == g["lcg_traj_first"]
    assert sum(traj) == g["lcg_traj_sum"]
    assert sha16(traj) == g["lcg_traj_sha"]


@pytest.mark.parametrize("name", ["c75", "c65", "m3a", "r13", "m1"])
def test_T_half_is_multiplicity(golden, name):
    """evaluate_symbolic_T(build_symbolic_T(...), 1/2) == len(transitions[i][j]) / 2^n, bit-exact
    (viterbi_markov.py:202-230, Pd_plotter.py:89-99): the identity that lets T_ref skip sympy."""
    g = golden["code_kats"][name]
    R = 1 << g["n"]
    want = np.array(g["T_edge"]["0.5"])
    assert np.array_equal(want, np.array(g["mult"], dtype=np.float64) / R)


@pytest.mark.parametrize("name", ["c75", "c65", "r13", "m1"])
@pytest.mark.parametrize("p", [0.1, 0.3])
def test_ref_port_numeric_T(golden, name, p):
    """Sparse numeric T(p) == the reference's sympy T(p) evaluated (1e-12: sympy evaluates the
    simplified polynomial, the port sums p^w (1-p)^(n-w))."""
    import ref_port
    g = golden["code_kats"][name]
    states, trans, all_r = ref_port.enumerate_states(g["gen"], g["m"], g["k"], g["n"])
    T = ref_port.numeric_T(states, trans, all_r, p)
    want = np.array(g["T_edge"][repr(p)])
    got = np.array([[T[i, g["next"][i][r]] for r in range(len(all_r))] for i in range(g["S"])])
    np.testing.assert_allclose(got, want, rtol=1e-12)


# ------------------------------------------------------------------ simulator (bits -> metrics)
SIM_CASES = [("c75_self", "c75", "c75"), ("c75_vs65", "c75", "c65"), ("c75_p001", "c75", "c75"),
             ("c75_p05", "c75", "c75"), ("m3_self", "m3a", "m3a"), ("m3_vs", "m3a", "m3b"),
             ("r13_self", "r13", "r13"), ("m1_self", "m1", "m1")]


@pytest.mark.parametrize("case,dec,enc", SIM_CASES)
def test_c_oracle_simulate_matches_reference(golden, codes_spec, case, dec, enc):
    """Philox words -> bits, encoder, BSC, Eq. 4-5 recursion == trajectory produced by the
    reference's branch/step functions (golden)."""
    import c_oracle as co
    g = golden["sim_kats"][case]
    d, e = codes_spec[dec], codes_spec[enc]
    n, m, N = d["n"], d["m"], g["N"]
    T = int(math.floor(g["p"] * 4294967296.0 + 0.5))
    U, E = co.trial_words(g["seed"], g["stream"], g["trial"], N, n, min(T, 0xFFFFFFFF))
    ubits = [(int(U[t >> 5]) >> (t & 31)) & 1 for t in range(N)]
    ebits = [[(int(E[j, t >> 5]) >> (t & 31)) & 1 for j in range(n)] for t in range(N)]
    assert ubits == g["u_bits"] and ebits == g["e_bits"]
    met, _ = co.enumerate_states(_taps(d), n, m, max_states=4096)
    tab = co.Table(met, m)
    idx, rseq, om = co.simulate(_taps(d), _taps(e), n, m, N, U, E, tab, want_metrics=True)
    assert om.tolist() == g["metrics"]
    assert [[(int(r) >> (n - 1 - j)) & 1 for j in range(n)] for r in rseq] == g["received"]
    assert [met[i].tolist() for i in idx] == g["metrics"]


@pytest.mark.parametrize("case,dec,enc", [c for c in SIM_CASES if c[1] != "m3a"])
def test_ref_port_simulate_matches_reference(golden, codes_spec, case, dec, enc):
    import ref_port
    g = golden["sim_kats"][case]
    d, e = codes_spec[dec], codes_spec[enc]
    sim = ref_port.simulate_markov_sequence(e["gen"], d["m"], 1, d["n"], g["N"], g["p"], True, g["seed"],
                                            decoder_matrix=d["gen"], stream=g["stream"], trial=g["trial"])
    assert [list(x) for x in sim["metrics"]] == g["metrics"]


def test_philox_known_answers():
    """Philox4x32-10 KATs from the Random123 distribution (kat_vectors): the bit source both the
    oracle and the kernels implement."""
    import c_oracle as co
    assert co.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert co.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert co.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


# ------------------------------------------------------------------ experiments (Pd / Pc / P1 / logp)
@pytest.mark.parametrize("exp", ["c75_c65_small", "c75_c65_lap", "m3_small", "c75_c65_n1000"])
def test_c_oracle_experiment_matches_reference(golden, exp):
    """The whole path on the CPU oracle == the reference's run_experiment (unmodified, with the
    injected simulator): P1 (edge form), every per-trial log-likelihood in call order, Pd, Pc."""
    import c_oracle as co
    g = golden["experiments"][exp]
    n, m = g["n"], g["m"]
    d = dict(gen=g["gen1"])
    t1, t2 = _taps(dict(gen=g["gen1"])), _taps(dict(gen=g["gen2"]))
    met, nxt = co.enumerate_states(t1, n, m, max_states=4096)
    S, R = met.shape[0], 1 << n
    tab = co.Table(met, m)
    mult = (nxt[:, :, None] == nxt[:, None, :]).sum(axis=2)
    Tref = mult / float(R)
    L = max(5000, 200 * S) if g["learn_len"] is None else g["learn_len"]
    lam = g["laplace"]
    P1 = {}
    for p in g["p_vec"]:
        T = min(int(math.floor(p * 4294967296.0 + 0.5)), 0xFFFFFFFF)
        edge, dense = co.learn_chain(t1, t1, n, m, L, g["learn_burn"], T, g["seed"], 0xFFFFFFFF, 0, tab, dense=True)
        assert int(edge.sum()) == L - g["learn_burn"]
        P = dense + lam                                        # Pd_plotter.py:166-167
        P /= P.sum(axis=1, keepdims=True)
        pe = np.array([[P[i, nxt[i, r]] for r in range(R)] for i in range(S)])
        want = np.array(g["P1_edge"][repr(p)]["edge"])
        assert np.array_equal(pe, want)
        assert hashlib.sha256(P.tobytes()).hexdigest()[:16] == g["P1_edge"][repr(p)]["sha"]
        P1[p] = pe
    logs = np.array(g["logps"]).reshape(len(g["N_list"]) * len(g["p_vec"]), g["num_iter"], 2, 2)
    q, rows = 0, []
    for N in g["N_list"]:
        for p in g["p_vec"]:
            T = min(int(math.floor(p * 4294967296.0 + 0.5)), 0xFFFFFFFF)
            s = []
            for h, enc in enumerate((t1, t2)):
                tally, lp = co.run_trials(t1, enc, n, m, N, T, g["seed"], 2 * q + h, 0, g["num_iter"], tab, P1[p],
                                          Tref, h, want_logp=True)
                assert np.array_equal(lp, logs[q, :, h, :])
                s.append(tally)
            rows.append({"N": N, "p": p, "Pd": s[0] / g["num_iter"], "Pc": (s[0] + s[1]) / (2 * g["num_iter"])})
            q += 1
    assert rows == g["rows"]
    del d


def test_ref_port_experiment_matches_reference(golden):
    """Pure-Python port (the CPU baseline bench.py times) == reference, on the smallest golden run."""
    import ref_port
    g = golden["experiments"]["c75_c65_lap"]
    collect = []
    rows = ref_port.run_experiment(g["k"], g["n"], g["m"], g["gen1"], g["gen2"], g["num_iter"], g["p_vec"],
                                   g["learn_len"], g["learn_burn"], g["laplace"], g["seed"], N_spectrum=g["N_list"],
                                   collect=collect)
    assert [{k: r[k] for k in ("N", "p", "Pd", "Pc")} for r in rows] == g["rows"]
    flat = [v for c in collect for v in c[2:]]
    assert flat == g["logps"]
