"""CPU tests of the product's HOST side (mvd.codes / mvd.bitsource / the drop-in modules' host
functions) against the golden vectors from the reference and against the oracle.  The device
hot path is not exercised here (no GPU); see tests/test_gpu_parity.py."""
import hashlib
import math

import numpy as np
import pytest


def sha16(obj):
    return hashlib.sha256(repr(obj).encode()).hexdigest()[:16]


SMALL = ["c75", "c65", "m3a", "m3b", "r13", "m1"]


@pytest.mark.parametrize("name", SMALL)
def test_dropin_trellis_api(golden, name):
    """viterbi_markov.build_trellis / branch_output_and_next_state: reference structures."""
    import viterbi_markov as vm
    g = golden["code_kats"][name]
    tr = vm.build_trellis(g["gen"], g["m"], g["k"])
    assert {str(ns): [[ps, list(u), list(o)] for ps, u, o in lst] for ns, lst in tr.items()} == g["trellis"]
    for s, u, out, ns in g["branches"]:
        assert vm.branch_output_and_next_state(s, tuple(u), g["gen"], g["m"], g["k"]) == (tuple(out), ns)


def test_dropin_small_helpers():
    import viterbi_markov as vm
    assert vm.state_bits_from_int(6, 3) == [0, 1, 1]             # LSB first (viterbi_markov.py:60-66)
    assert vm.bits_to_int([0, 1, 1]) == 6                        # :70-75
    assert vm.hamming_distance((0, 1, 1), (1, 1, 0)) == 2        # :109-111
    assert vm.hamming_distance((0, 1), (1, 1, 0)) == 1           # zip-truncating


def test_dropin_step_kats(golden):
    import viterbi_markov as vm
    g = golden["code_kats"]["c75"]
    tr = vm.build_trellis(g["gen"], 2, 1)
    for d, r, want in g["step_kats"]:
        out = vm.viterbi_metric_step(d, tr, tuple(r))
        assert isinstance(out, tuple) and list(out) == want


@pytest.mark.parametrize("name", SMALL)
def test_dropin_enumeration(golden, name):
    """enumerate_markov_states_allzero: same states, same BFS order, same transitions/all_r."""
    import viterbi_markov as vm
    g = golden["code_kats"][name]
    states, trans, all_r = vm.enumerate_markov_states_allzero(g["gen"], g["m"], g["k"], g["n"])
    assert len(states) == g["S"]
    assert sha16(states) == g["states_sha"]
    assert [list(r) for r in all_r] == g["all_r"]
    assert sum(len(trans[i]) for i in range(len(states))) == g["nnz"]
    tab = vm.state_table(g["gen"], g["m"], g["k"], g["n"])
    assert sha16([[int(v) for v in row] for row in tab.nxt]) == g["next_sha"]
    assert tab.max_metric == g["max_metric"]
    assert tab.mult.tolist() == g["mult"]
    for i in (0, len(states) // 2, len(states) - 1):
        for j, rl in trans[i].items():
            assert all(isinstance(r, tuple) for r in rl)
            assert len(rl) == g["mult"][i][g["next"][i].index(j)]


def test_enumeration_matches_oracle_m4():
    """m = 4 (31,33): S = 25 751 -- host numpy BFS == C oracle BFS."""
    import c_oracle as co
    from mvd import codes
    gen = codes.freeze_generator([[[1, 1, 0, 0, 1]], [[1, 1, 0, 1, 1]]])
    tab = codes.enumerate_states(gen, 4, 1, 2)
    assert tab.S == 25751
    om, on = co.enumerate_states(codes.tap_masks(gen, 4, 1), 2, 4)
    assert np.array_equal(tab.metrics, om) and np.array_equal(tab.nxt, on)


def test_enumeration_max_states_guard():
    from mvd import codes
    gen = codes.freeze_generator([[[1, 1, 1, 1]], [[1, 0, 1, 1]]])
    with pytest.raises(MemoryError):
        codes.enumerate_states(gen, 3, 1, 2, max_states=100)


def test_transitions_roundtrip(golden):
    from mvd import codes
    g = golden["code_kats"]["c75"]
    tab = codes.enumerate_states(codes.freeze_generator(g["gen"]), 2, 1, 2)
    trans, all_r = codes.transitions_from_table(tab)
    back = codes.table_from_transitions(tab.state_tuples(), trans, 2)
    assert np.array_equal(back.nxt, tab.nxt) and np.array_equal(back.mult, tab.mult)


@pytest.mark.parametrize("name", ["c75", "c65", "m1"])
def test_symbolic_T(golden, name):
    """build_symbolic_T entries == the reference's (as strings, after sympy simplification) and
    evaluate_symbolic_T at 0.5 / 0.1 / 0.3 == golden."""
    import sympy as sp
    import Pd_plotter as pdp
    import viterbi_markov as vm
    g = golden["code_kats"][name]
    states, trans, all_r = vm.enumerate_markov_states_allzero(g["gen"], g["m"], g["k"], g["n"])
    p, T = vm.build_symbolic_T(states, trans, all_r)
    assert T.shape == (g["S"], g["S"])
    for key, want in g["T_sym_str"].items():
        i, j = map(int, key.split(","))
        assert sp.simplify(T[i, j] - sp.sympify(want, locals={"p": p})) == 0
    for pv in (0.5, 0.1, 0.3):
        Tn = pdp.evaluate_symbolic_T(T, p, pv)
        got = np.array([[Tn[i, g["next"][i][r]] for r in range(len(all_r))] for i in range(g["S"])])
        np.testing.assert_allclose(got, np.array(g["T_edge"][repr(pv)]), rtol=1e-12)
        np.testing.assert_allclose(Tn.sum(axis=1), 1.0, rtol=1e-12)


@pytest.mark.parametrize("name", ["c75", "c65", "m3a", "r13", "m1"])
def test_tref_half_table(golden, name):
    """T_ref shipped to the GPU == evaluate_symbolic_T(T, p, 0.5) on the edges, bit-exact."""
    import viterbi_markov as vm
    from mvd import codes
    g = golden["code_kats"][name]
    tab = vm.state_table(g["gen"], g["m"], g["k"], g["n"])
    assert np.array_equal(codes.tref_half_table(tab), np.array(g["T_edge"]["0.5"]))


@pytest.mark.parametrize("name", ["c75", "c65", "m3a", "r13", "m1"])
@pytest.mark.parametrize("pv", [0.5, 0.1, 0.3])
def test_numeric_T_edges_match_reference_symbolic(golden, name, pv):
    """the sympy-free T(p) (codes.t_edge_table / vm.numeric_T) equals the reference's
    evaluate_symbolic_T(build_symbolic_T(...), p) on every edge (golden, generated by the reference)."""
    import viterbi_markov as vm
    from mvd import codes
    g = golden["code_kats"][name]
    tab = vm.state_table(g["gen"], g["m"], g["k"], g["n"])
    got = codes.t_edge_table(tab, pv)
    np.testing.assert_allclose(got, np.array(g["T_edge"][repr(pv)]), rtol=1e-13, atol=0)
    if pv == 0.5:
        assert np.array_equal(got, codes.tref_half_table(tab))
    if g["S"] <= 100:
        states, trans, all_r = vm.enumerate_markov_states_allzero(g["gen"], g["m"], g["k"], g["n"])
        dense = vm.numeric_T(states, trans, all_r, pv)
        assert dense.shape == (g["S"], g["S"])
        np.testing.assert_allclose(dense.sum(axis=1), 1.0, rtol=1e-13)
        assert np.count_nonzero(dense) == g["nnz"] or pv in (0.0, 1.0)
        np.testing.assert_allclose(dense[np.arange(g["S"])[:, None], tab.nxt], got, rtol=0, atol=0)


def test_host_p1_edge_tables_equal_the_numpy_closed_form():
    """mvd_host_p1_edge_tables (C, host threads) == the numpy statements of the closed form, bit for bit, on a state
    table above DENSE_LIMIT (random successor structure with merging edges, zero rows, counts up to 2^40) for a
    dyadic and a non-dyadic Laplace constant; and the threaded log table == math.log element by element."""
    import math
    from mvd import codes, engine
    rng = np.random.default_rng(3)
    S, R, T = codes.DENSE_LIMIT + 1500, 4, 3
    nxt = rng.integers(0, S, size=(S, R), dtype=np.uint32)
    nxt[::3, 1] = nxt[::3, 0]                          # edges that share a successor
    nxt[::7, 3] = nxt[::7, 2]
    tab = codes.StateTable(k=1, n=2, m=4, metrics=np.zeros((S, 16), dtype=np.uint8), nxt=nxt)
    counts = rng.integers(0, 1 << 40, size=(T, S, R), dtype=np.uint64)
    counts[0, ::5] = 0                                 # unvisited rows
    counts[1] = rng.integers(0, 3, size=(S, R))
    for laplace in (1.0, 0.3):
        got = codes.p1_tables_from_edge_counts(tab, counts, laplace)
        for t in range(T):
            assert np.array_equal(got[t], codes.p1_closed_form_numpy(tab, counts[t], laplace))
        assert np.array_equal(codes.p1_from_edge_counts(tab, counts[2], laplace), got[2])
    big = rng.random(300000) ** 8
    big[::1000] = 0.0
    logs = engine._log_table(big)
    assert logs.tolist() == [math.log(max(float(v), 1e-300)) for v in big]


def test_small_tables_closed_form_equals_the_dense_replay():
    """Small state tables (S <= DENSE_LIMIT): with integer counts and a dyadic Laplace constant the closed form in C
    gives the same bits as replaying the reference's dense ``P = counts + laplace; P /= P.sum(axis=1)``
    (Pd_plotter.py:166-167) -- real tables of the m = 2 / m = 3 codes, learned-count-sized and huge counts, zero rows;
    a non-dyadic constant still takes the dense replay."""
    from mvd import codes
    rng = np.random.default_rng(11)
    for gen, m in (([[[1, 1, 1]], [[1, 0, 1]]], 2), ([[[1, 1, 1, 1]], [[1, 0, 1, 1]]], 3)):
        tab = codes.enumerate_states(gen, m, 1, 2)
        S, R = tab.S, tab.R
        counts = rng.integers(0, 5000, size=(4, S, R), dtype=np.uint64)
        counts[1, ::4] = 0
        counts[2] = rng.integers(0, 1 << 35, size=(S, R), dtype=np.uint64)
        counts[3] = 0
        for laplace in (1.0, 0.5, 2.0, 0.0009765625, 0.3):
            assert codes._exact_row_sums(counts, laplace, S) == (laplace != 0.3)
            got = codes.p1_tables_from_edge_counts(tab, counts, laplace)
            for t in range(4):
                dense = codes.p1_dense(tab, counts[t], laplace)
                assert np.array_equal(got[t], dense[np.arange(S)[:, None], tab.nxt]), (m, laplace, t)
    assert not codes._exact_row_sums(np.zeros((1, 3, 4)), 1.0, 3)                  # float counts: not taken for exact
    assert not codes._exact_row_sums(np.full((1, 3, 4), 1 << 37, dtype=np.uint64), 1.0, 3)


def test_host_log_table_is_math_log():
    """The C helper behind large log-likelihood tables returns math.log(max(v, 1e-300)) bit for bit
    (Pd_plotter.py:114-115) -- and _log_table switches to it without changing a single value."""
    import math
    from mvd import engine
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.random(20000), rng.random(2000) * 1e-6, [0.0, 1e-300, 1e-310, 1.0, 0.25, 0.5]])
    got = engine._log_table(v)
    want = np.array([math.log(max(x, 1e-300)) for x in v.tolist()])
    assert np.array_equal(got, want)
    assert np.array_equal(engine._log_table(v[:100]), want[:100])


def test_log_prob_sequence_host(golden):
    """Host log_prob_sequence (Pd_plotter.py:106-116) reproduces the reference's per-trial sums."""
    import Pd_plotter as pdp
    import ref_port
    from mvd import codes
    g = golden["experiments"]["c75_c65_lap"]
    tab = codes.enumerate_states(codes.freeze_generator(g["gen1"]), g["m"], g["k"], g["n"])
    states = tab.state_tuples()
    index = {s: i for i, s in enumerate(states)}
    Tref = codes.dense_from_edges(tab, codes.tref_half_table(tab))
    sim = ref_port.simulate_markov_sequence(g["gen1"], g["m"], 1, g["n"], g["N_list"][0], g["p_vec"][0], True,
                                            g["seed"], decoder_matrix=g["gen1"], stream=0, trial=0)
    assert pdp.log_prob_sequence(sim["metrics"], index, Tref) == g["logps"][1]
    with pytest.raises(KeyError):
        pdp.log_prob_sequence([(9, 9, 9, 9), (0, 0, 0, 0)], index, Tref)


@pytest.mark.parametrize("exp", ["c75_c65_small", "c75_c65_lap", "m3_small", "c75_c65_n1000"])
def test_p1_from_edge_counts(golden, exp):
    """Host Laplace + row normalisation (Pd_plotter.py:166-167) from oracle edge counts == the
    reference's learned P1 on every edge; off-edge entries == the reference's minimum."""
    import c_oracle as co
    from mvd import bitsource, codes
    g = golden["experiments"][exp]
    gen = codes.freeze_generator(g["gen1"])
    tab = codes.enumerate_states(gen, g["m"], g["k"], g["n"])
    taps = codes.tap_masks(gen, g["m"], g["k"])
    otab = co.Table(tab.metrics, g["m"])
    L = max(5000, 200 * tab.S) if g["learn_len"] is None else g["learn_len"]
    for p in g["p_vec"]:
        edge, _ = co.learn_chain(taps, taps, g["n"], g["m"], L, g["learn_burn"], bitsource.bsc_threshold(p), g["seed"],
                                 bitsource.LEARN_STREAM, 0, otab)
        got = codes.p1_from_edge_counts(tab, edge, g["laplace"])
        want = np.array(g["P1_edge"][repr(p)]["edge"])
        assert np.array_equal(got, want)
        dense = codes.p1_dense(tab, edge, g["laplace"])
        assert hashlib.sha256(dense.tobytes()).hexdigest()[:16] == g["P1_edge"][repr(p)]["sha"]
        assert float(dense.min()) == g["P1_edge"][repr(p)]["off_edge_min"]


def test_p1_closed_form_matches_dense():
    """Above DENSE_LIMIT the closed form is used: check it against the dense replay (1e-12;
    bit-exact for laplace = 1)."""
    from mvd import codes
    gen = codes.freeze_generator([[[1, 1, 1, 1]], [[1, 0, 1, 1]]])
    tab = codes.enumerate_states(gen, 3, 1, 2)
    rng = np.random.default_rng(0)
    edge = rng.integers(0, 50, size=(tab.S, tab.R)).astype(np.uint64)
    saved = codes.DENSE_LIMIT
    try:
        for lam in (1.0, 0.5, 0.1):
            codes.DENSE_LIMIT = 1 << 20
            dense = codes.p1_from_edge_counts(tab, edge, lam)
            codes.DENSE_LIMIT = 0
            closed = codes.p1_from_edge_counts(tab, edge, lam)
            if lam in (1.0, 0.5):
                assert np.array_equal(dense, closed)
            else:
                np.testing.assert_allclose(closed, dense, rtol=1e-12)
    finally:
        codes.DENSE_LIMIT = saved


# ------------------------------------------------------------------ bit source / layout
def test_bitsource_philox_matches_oracle():
    import c_oracle as co
    from mvd import bitsource
    rng = np.random.default_rng(1)
    for _ in range(20):
        ctr = [int(x) for x in rng.integers(0, 1 << 32, 4)]
        key = [int(x) for x in rng.integers(0, 1 << 32, 2)]
        assert list(bitsource.philox4x32_10(tuple(ctr), tuple(key))) == co.philox(ctr, key)
    seed = (7 << 32) | 9
    v = bitsource.philox4x32_10_np(np.arange(5), 3, 4, 5, seed)
    for q in range(5):
        assert v[q].tolist() == co.philox([q, 3, 4, 5], [9, 7])


@pytest.mark.parametrize("N,p", [(1, 0.1), (32, 0.5), (33, 0.001), (129, 0.25), (500, 0.1), (257, 1.0), (64, 0.0)])
def test_bitsource_trial_words_match_oracle(N, p):
    """MVD-PHILOX-2 in the product's host module == the C oracle's independent implementation."""
    import c_oracle as co
    from mvd import bitsource
    T = bitsource.bsc_threshold(p)
    for trial in (0, 5, (1 << 35) + 3):
        U, E = bitsource.trial_words(99, 6, trial, N, 2, T)
        oU, oE = co.trial_words(99, 6, trial, N, 2, T)
        assert np.array_equal(U, oU) and np.array_equal(E, oE)
    if p == 0.0:
        assert not E.any()
    if p == 1.0:      # threshold clipped to 2^32 - 1: all valid lanes flip (up to 2^-32)
        bits = bitsource.words_to_bits(E, N)
        assert bits.sum() >= 2 * N - 1


def test_bsc_threshold():
    from mvd import bitsource
    assert bitsource.bsc_threshold(0.0) == 0
    assert bitsource.bsc_threshold(0.5) == 1 << 31
    assert bitsource.bsc_threshold(1.0) == 0xFFFFFFFF
    assert abs(bitsource.bsc_threshold(0.1) / 2 ** 32 - 0.1) <= 2 ** -33
    with pytest.raises(ValueError):
        bitsource.bsc_threshold(1.5)


def test_flip_rate_statistics():
    """Lazy Bernoulli words flip at rate p (5 sigma)."""
    from mvd import bitsource
    p, N = 0.1, 4096
    T = bitsource.bsc_threshold(p)
    tot = 0
    for trial in range(8):
        _, E = bitsource.trial_words(1, 0, trial, N, 2, T)
        tot += int(bitsource.words_to_bits(E, N).sum())
    cnt = 8 * 2 * N
    assert abs(tot - p * cnt) < 5 * math.sqrt(cnt * p * (1 - p))


def test_pack_bitstreams_layout():
    """128-bit word index (sb * (1 + n) + c) * ntrials + trial; bit b of lane w = step 128 sb + 32 w + b."""
    from mvd import bitsource
    rng = np.random.default_rng(3)
    ntr, n, N = 5, 2, 300
    u = rng.integers(0, 2, (ntr, N), dtype=np.uint8)
    e = rng.integers(0, 2, (ntr, n, N), dtype=np.uint8)
    w = bitsource.pack_bitstreams(u, e)
    assert w.shape == (3, 3, ntr, 4) and w.dtype == np.uint32
    flat = w.reshape(-1, 4)
    for (tr, t) in [(0, 0), (4, 299), (2, 127), (3, 128), (1, 31), (1, 32)]:
        sb, lane, b = t // 128, (t % 128) // 32, t % 32
        assert (int(flat[(sb * 3 + 0) * ntr + tr, lane]) >> b) & 1 == u[tr, t]
        for j in range(n):
            assert (int(flat[(sb * 3 + 1 + j) * ntr + tr, lane]) >> b) & 1 == e[tr, j, t]
    # padding bits beyond N are zero
    assert (int(flat[(2 * 3) * ntr, 1]) >> 12) == 0
    back = bitsource.words_to_bits(bitsource.bits_to_words(u), N)
    assert np.array_equal(back, u)


def test_tap_masks_and_encoder_tables(golden):
    from mvd import codes
    g = golden["code_kats"]["c65"]
    gen = codes.freeze_generator(g["gen"])
    assert codes.tap_masks(gen, 2, 1) == [0b011, 0b101]
    eo, en = codes.encoder_tables(gen, 2, 1)
    for s, u, out, ns in g["branches"]:
        assert int(eo[s, u[0]]) == 2 * out[0] + out[1] and int(en[s, u[0]]) == ns
    with pytest.raises(ValueError):
        codes.tap_masks(codes.freeze_generator([[[1, 1], [1, 0]], [[1, 0], [1, 1]]]), 1, 2)


def test_k2_quirk_is_reproduced():
    """For k > 1 every input sees the same register (reference quirk): host tables
    follow the reference's arithmetic, the port in oracle/ agrees."""
    import ref_port
    from mvd import codes
    gen = [[[1, 1, 1], [1, 0, 1]], [[1, 0, 1], [1, 1, 1]]]
    for s in range(4):
        for u in [(0, 0), (0, 1), (1, 0), (1, 1)]:
            assert codes.encoder_branch(s, u, gen, 2, 2) == ref_port.branch(s, u, gen, 2, 2)


# ------------------------------------------------------------------ drop-in module surface
def test_dropin_module_surface():
    """Names, defaults and signatures the reference's callers rely on."""
    import inspect

    import Pd_plotter as pdp
    import demo_script
    import viterbi_markov as vm
    for name in ("state_bits_from_int", "bits_to_int", "branch_output_and_next_state", "hamming_distance",
                 "build_trellis", "viterbi_metric_step", "enumerate_markov_states_allzero", "build_symbolic_T",
                 "simulate_markov_sequence"):
        assert callable(getattr(vm, name))
    assert pdp.DEFAULTS == {"num_iter": 10000, "p_vec": [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5], "seed": 12345,
                            "learn_len": None, "learn_burn": 200, "laplace": 1.0, "save_dir": "results_experiments"}
    assert pdp.N_SPECTRUM_BY_M == {1: [5, 10, 20, 50, 100, 200], 2: [500], 3: [500], 4: [50, 100, 200, 300, 500]}
    sig = list(inspect.signature(pdp.run_experiment).parameters)
    assert sig[:11] == ["k", "n", "m", "gen1", "gen2", "num_iter", "p_vec", "learn_len", "learn_burn", "laplace", "seed"]
    sig = list(inspect.signature(pdp.learn_P1_empirical.__wrapped__).parameters)
    assert sig == ["gens_tuple", "k", "n", "m", "p", "learn_len", "learn_burn", "laplace", "seed"]
    sig = list(inspect.signature(vm.simulate_markov_sequence).parameters)
    assert sig[:8] == ["generator_matrix", "m", "k", "n", "length", "p_val", "random_input", "seed"]
    assert set(demo_script.EXAMPLE_CODES) >= {"1", "2"}
    assert demo_script.EXAMPLE_CODES["1"]["gen1"] == [[[1, 1, 1]], [[1, 0, 1]]]
    assert callable(demo_script.read_generators)


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (or any CPU fallback)."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "detecting-convolutional-codes-via-markovian-statistics_b200")
    pat = re.compile(r"^\s*(from|import)\s+(c_oracle|ref_port|oracle)\b", re.M)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f"{f} imports the oracle"
                assert "mvd_oracle" not in text, f"{f} references the oracle library"


def test_hot_path_fails_loudly_without_gpu():
    """No CUDA device here: the public entry points raise instead of falling back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import Pd_plotter as pdp
    import viterbi_markov as vm
    from mvd import _capi
    gen1, gen2 = [[[1, 1, 1]], [[1, 0, 1]]], [[[1, 1, 0]], [[1, 0, 1]]]
    with pytest.raises(_capi.MvdError) as ei:
        pdp.run_experiment(1, 2, 2, gen1, gen2, 4, [0.1], None, 200, 1.0, 1)
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(_capi.MvdError):
        vm.simulate_markov_sequence(gen1, 2, 1, 2, 10, 0.1, True, 1)
    # the other entry points of the reference (alpha_exponent.py, comp_parity.py) and the GPU enumeration
    import alpha_exponent as ae
    import comp_parity as cp
    from mvd.engine import Detector
    taps = [[1, 1, 1], [1, 0, 1]]
    with pytest.raises(_capi.MvdError):
        ae.learn_transition_tensor(taps, taps, 2, 0.1, length=100, burn_in=10, seed=1)
    with pytest.raises(_capi.MvdError):
        ae.compute_error_exponent(np.full((3, 3, 4), 1 / 12), np.full((3, 3, 4), 1 / 12), 5)
    with pytest.raises(_capi.MvdError):
        cp.run_parity_experiment(gen1, gen2, 2, [50], [0.1], 0.6, 10, 1)
    with pytest.raises(_capi.MvdError):
        Detector(gen1, 1, 2, 2, enumerate_with="gpu")


def test_plots_compare_helpers_and_csv_chain(tmp_path, capsys):
    """plots_compare drop-in (reference plots_compare.py:35-148): helper semantics, and the CSVs committed under
    profiles/paper_sweeps (written by Pd_plotter.run_experiment / comp_parity.run_parity_experiment on the GPU) load."""
    import os
    import pandas as pd
    import plots_compare as pc
    assert pc.p_error([1.2, 0.75, -0.5]).tolist() == [0.0, 0.25, 1.0]
    df = pd.DataFrame({"N": [500, 100, 100, 500], "p": [0.3, 0.2, 0.1, 0.1], "Pd": [0.1, 0.5, 0.9, 1.0], "Pc": [0.5, 0.7, 0.95, 1.0]})
    x, y = pc.extract_by_N(df, 100)
    assert x.tolist() == [0.1, 0.2] and y.tolist() == [0.95, 0.7]
    x, y = pc.extract_by_p(df, 0.1 + 1e-12)
    assert x.tolist() == [100, 500] and y.tolist() == [0.95, 1.0]
    assert pc.extract_by_N(df, 7)[0].size == 0
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "paper_sweeps")
    hybrid = os.path.join(root, "Pd_hybrid_m2_7_5_vs_6_5_vs_p.csv")
    base = os.path.join(root, "results_parity", "Pd_parity_results.csv")
    only_pd = tmp_path / "only_pd.csv"
    pd.read_csv(base)[["N", "p", "Pd"]].to_csv(only_pd, index=False)          # a CSV without Pc: Pd stands in
    h, b = pc.load_results(hybrid, str(only_pd))
    assert list(h.columns) == ["N", "p", "Pd", "Pc"] and (b["Pc"] == b["Pd"]).all() and len(h) == 35 and len(b) == 35
    pc.main(hybrid, base, str(tmp_path / "plots"))                              # plots, or tables without matplotlib
    out = capsys.readouterr().out
    assert os.listdir(tmp_path / "plots") or "P_err hybrid" in out
