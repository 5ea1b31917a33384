"""Codes with k = 2 and k = 3 inputs per step: viterbi_markov.py:82-132 is generic in k (every input sees the register
[u_i, s_0, ..]).  The device takes such codes as tables (mvd_set_code_tables / mvd_set_encoders) and runs the Markov-state
walk with a table-driven encoder.  Golden vectors: tests/golden/k2_kats.json, written by `oracle/make_golden.py --k2-only`
from the reference's own functions (branch / trellis / step / BFS, and run_experiment unmodified with the injected simulator).
CPU tests pin the host tables and the oracle's *_tab functions; GPU tests compare the device path with both."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def k2():
    with open(os.path.join(GOLDEN, "k2_kats.json")) as f:
        return json.load(f)


def _tables(spec):
    from mvd import codes
    gen = codes.freeze_generator(spec["gen"])
    prev, blab = codes.trellis_arrays(gen, spec["m"], spec["k"], spec["n"])
    enc_out, enc_next = codes.encoder_tables(gen, spec["m"], spec["k"])
    return gen, prev, blab, enc_next, enc_out


# ------------------------------------------------------------------------------------------------ CPU: host tables + oracle
@pytest.mark.parametrize("name", ["k2b", "k2c", "k2d", "k3a", "k3b"])
def test_host_tables_match_reference(k2, name):
    """encoder_tables / trellis_arrays / enumerate_states for k = 2 == the reference's branch_output_and_next_state,
    build_trellis and enumerate_markov_states_allzero (state list, BFS order, NEXT)."""
    import itertools
    from mvd import codes
    g = k2["codes"][name]
    gen, prev, blab, enc_next, enc_out = _tables(g)
    inputs = list(itertools.product((0, 1), repeat=g["k"]))
    for s, u, out, nxt in g["branches"]:
        ui = inputs.index(tuple(u))
        assert enc_out[s, ui] == codes.label_of(out) and enc_next[s, ui] == nxt
    for ns, lst in g["trellis"].items():
        assert [int(v) for v in prev[int(ns)]] == [ps for ps, _, _ in lst]
        assert [int(v) for v in blab[int(ns)]] == [codes.label_of(o) for _, _, o in lst]
    tab = codes.enumerate_states(gen, g["m"], g["k"], g["n"])
    assert tab.S == g["S"]
    assert [list(map(int, row)) for row in tab.metrics] == g["states"]
    assert [list(map(int, row)) for row in tab.nxt] == g["next"]
    assert [list(map(int, row)) for row in tab.mult] == g["mult"]
    if "T_edge" in g:
        assert np.array_equal(codes.tref_half_table(tab), np.array(g["T_edge"]["0.5"]))


SIMS = [("k2c_self", "k2c", "k2c"), ("k2c_vs_d", "k2c", "k2d"), ("k2b_self", "k2b", "k2b"), ("k3a_self", "k3a", "k3a"),
        ("k3a_vs_b", "k3a", "k3b")]


@pytest.mark.parametrize("sim,dec,enc", SIMS)
def test_oracle_tab_trajectories(k2, sim, dec, enc):
    """mvdo_trial_words_k + mvdo_simulate_tab == the reference's branch + step functions under MVD-PHILOX-2 with k = 2
    (info bits of input i from slot 32 + i): bits, received words and metric vectors."""
    import c_oracle as co
    from mvd import bitsource, codes
    g, spec, espec = k2["sims"][sim], k2["codes"][dec], k2["codes"][enc]
    _, prev, blab, _, _ = _tables(spec)
    _, _, _, enc_next, enc_out = _tables(espec)
    k, n, m, N = spec["k"], spec["n"], spec["m"], g["N"]
    T = bitsource.bsc_threshold(g["p"])
    U, E = co.trial_words_k(g["seed"], g["stream"], g["trial"], N, k, n, T)
    U2, E2 = bitsource.trial_words_k(g["seed"], g["stream"], g["trial"], N, k, n, T)
    assert np.array_equal(U, U2) and np.array_equal(E, E2)
    assert bitsource.words_to_bits(U, N).T.tolist() == g["u_bits"]
    assert bitsource.words_to_bits(E, N).T.tolist() == g["e_bits"]
    tab = co.Table(np.array(spec["states"], dtype=np.uint8), m)
    idx, rseq, met = co.simulate_tab(prev, blab, enc_next, enc_out, k, n, m, N, U, E, tab, want_metrics=True)
    assert met.tolist() == g["metrics"]
    assert [codes.label_of(r) for r in g["received"]] == rseq.tolist()
    assert [spec["states"][i] for i in idx] == g["metrics"]


def test_oracle_tab_experiment(k2):
    """The k = 2 golden experiment (reference run_experiment, unmodified) through the oracle's *_tab functions: learned edge
    counts -> P1, per-trial log-likelihood pairs and tallies."""
    import c_oracle as co
    from mvd import bitsource, codes
    g = k2["experiments"]["k2c_k2d"]
    spec = dict(k=g["k"], n=g["n"], m=g["m"], gen=g["gen1"])
    gen1, prev, blab, en1, eo1 = _tables(spec)
    _, _, _, en2, eo2 = _tables(dict(spec, gen=g["gen2"]))
    k, n, m = g["k"], g["n"], g["m"]
    table = codes.enumerate_states(gen1, m, k, n)
    tab = co.Table(table.metrics, m)
    Tref = codes.tref_half_table(table)
    L = max(5000, 200 * table.S)
    logs = np.array(g["logps"]).reshape(len(g["N_list"]) * len(g["p_vec"]), g["num_iter"], 2, 2)
    P1 = {}
    for p in g["p_vec"]:
        edge = co.learn_chain_tab(prev, blab, en1, eo1, k, n, m, L, g["learn_burn"], bitsource.bsc_threshold(p), g["seed"],
                                  bitsource.LEARN_STREAM, 0, tab)
        P1[p] = codes.p1_from_edge_counts(table, edge, g["laplace"])
        assert np.array_equal(P1[p], np.array(g["P1_edge"][repr(p)]["edge"]))
    q = 0
    rows = []
    for N in g["N_list"]:
        for p in g["p_vec"]:
            wins = []
            for h, (en, eo) in enumerate(((en1, eo1), (en2, eo2))):
                w, lp = co.run_trials_tab(prev, blab, en, eo, k, n, m, N, bitsource.bsc_threshold(p), g["seed"], 2 * q + h, 0,
                                          g["num_iter"], tab, P1[p], Tref, h, want_logp=True)
                assert np.array_equal(lp, logs[q, :, h, :])
                wins.append(w)
            rows.append((N, p, wins[0] / g["num_iter"], (wins[0] + wins[1]) / (2 * g["num_iter"])))
            q += 1
    assert [(r["N"], r["p"], r["Pd"], r["Pc"]) for r in g["rows"]] == rows


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("sim,dec,enc", SIMS)
@pytest.mark.parametrize("source", ["philox", "bitstream"])
def test_gpu_trace_k2(k2, sim, dec, enc, source):
    """simulate_markov_sequence (drop-in) for k = 2 on the device: the metric trajectory of the reference's own functions,
    from the on-device bit source and from host-supplied bits."""
    import viterbi_markov as vm
    g, spec, espec = k2["sims"][sim], k2["codes"][dec], k2["codes"][enc]
    kw = dict(decoder_matrix=spec["gen"], stream=g["stream"], trial=g["trial"])
    if source == "bitstream":
        kw.update(u_bits=g["u_bits"], e_bits=g["e_bits"])
    out = vm.simulate_markov_sequence(espec["gen"], spec["m"], spec["k"], spec["n"], g["N"], g["p"], True, g["seed"], **kw)
    assert [list(d) for d in out["metrics"]] == g["metrics"]


@pytest.mark.gpu
def test_gpu_run_experiment_k2(k2):
    """Drop-in run_experiment with a k = 2 code pair == the reference's run_experiment (unmodified, injected simulator):
    rows, CSV text, P1 and per-trial log-likelihood pairs, bit for bit."""
    import Pd_plotter as pdp
    import viterbi_markov as vm
    from mvd import bitsource, codes
    from mvd.engine import Seg
    g = k2["experiments"]["k2c_k2d"]
    details = {}
    df = pdp.run_experiment(g["k"], g["n"], g["m"], g["gen1"], g["gen2"], g["num_iter"], g["p_vec"], g["learn_len"],
                            g["learn_burn"], g["laplace"], g["seed"], N_spectrum=g["N_list"], details=details)
    assert df.to_dict(orient="records") == g["rows"]
    assert df.to_csv(index=False) == g["csv"]
    for i, p in enumerate(details["distinct_p"]):
        assert np.array_equal(details["p1_tables"][i], np.array(g["P1_edge"][repr(p)]["edge"]))
    det = vm._detector(codes.freeze_generator(g["gen1"]), g["k"], g["n"], g["m"], 0)
    assert det.table_code
    logs = np.array(g["logps"]).reshape(len(g["N_list"]) * len(g["p_vec"]), g["num_iter"], 2, 2)
    tindex = {p: i for i, p in enumerate(details["distinct_p"])}
    q = 0
    for N in g["N_list"]:
        for p in g["p_vec"]:
            for h, gen in enumerate((g["gen1"], g["gen2"])):
                seg = Seg(N=N, threshold=bitsource.bsc_threshold(p), stream=2 * q + h, table=tindex[p], enc_taps=det.taps_of(gen),
                          decide=h, trial_begin=0, trial_end=g["num_iter"])
                _, lp = det.detect([seg], seed=g["seed"], engine="auto", want_logp=True)
                assert np.array_equal(lp, logs[q, :, h, :])
            q += 1


@pytest.mark.gpu
def test_gpu_k2_vs_oracle_and_envelope(k2):
    """3 000 trials per hypothesis and a 60 000-step learning chain of a k = 2 pair against the oracle's *_tab functions (tallies,
    per-trial sums, edge counts, ragged N); what the table form cannot do says so (MVD_E_UNSUPPORTED), and a state table that
    does not close under Eq. 4-5 on the given trellis is refused."""
    import c_oracle as co
    from mvd import _capi, bitsource, codes
    from mvd.engine import Detector, Seg
    spec, espec = k2["codes"]["k2c"], k2["codes"]["k2d"]
    gen1, prev, blab, en1, eo1 = _tables(spec)
    _, _, _, en2, eo2 = _tables(espec)
    k, n, m = spec["k"], spec["n"], spec["m"]
    with Detector(spec["gen"], k, n, m) as det:
        tab = co.Table(det.table.metrics, m)
        T = bitsource.bsc_threshold(0.08)
        learn = Seg(N=60000, threshold=T, stream=bitsource.LEARN_STREAM, enc_taps=det.dec_taps)
        counts = det.learn_counts([learn], burn=200, seed=31)[0]
        want = co.learn_chain_tab(prev, blab, en1, eo1, k, n, m, 60000, 200, T, 31, bitsource.LEARN_STREAM, 0, tab)
        assert np.array_equal(counts, want)
        P1 = codes.p1_from_edge_counts(det.table, counts, 1.0)
        Tref = codes.tref_half_table(det.table)
        det.set_models([P1])
        ntr = 3000
        for N in (333, 97):
            segs = [Seg(N=N, threshold=T, stream=h, enc_taps=det.taps_of((spec, espec)[h]["gen"]), decide=h, trial_begin=11,
                        trial_end=11 + ntr) for h in (0, 1)]
            tallies, lp = det.detect(segs, seed=31, engine="fsm", want_logp=True)
            assert det.last_kernel_kind() == 0                     # the generic, checked kernels
            for h, (en, eo) in enumerate(((en1, eo1), (en2, eo2))):
                w, wlp = co.run_trials_tab(prev, blab, en, eo, k, n, m, N, T, 31, h, 11, 11 + ntr, tab, P1, Tref, h, want_logp=True)
                assert int(tallies[h]) == w and np.array_equal(lp[h * ntr:(h + 1) * ntr], wlp)
        with pytest.raises(_capi.MvdError) as exc:
            det.detect(segs, seed=31, engine="acs")
        assert exc.value.code == -3
        bad = det.table.nxt.copy()
        bad[3, 1] = (bad[3, 1] + 1) % det.table.S
        met = np.ascontiguousarray(det.table.metrics, dtype=np.uint8)
        assert det.lib.mvd_set_states(det.ctx, det.table.S, met.ctypes.data, np.ascontiguousarray(bad, dtype=np.uint32).ctypes.data) == 0
        with pytest.raises(_capi.MvdError):
            det.detect(segs, seed=31, engine="fsm")


@pytest.mark.gpu
def test_integration_md_table_stub_runs_as_written(k2):
    """The k > 1 stub of INTEGRATION.md (tables from the reference-named trellis / branch functions -> mvd_set_code_tables /
    mvd_set_encoders) executed verbatim, then one detection call through the raw C ABI == the Detector path."""
    import ctypes as C
    import re
    import viterbi_markov as vm
    from mvd import _capi, bitsource, codes
    from mvd.engine import Detector, Seg
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    snippet = [b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S) if "mvd_set_code_tables" in b][0]
    spec, espec = k2["codes"]["k2c"], k2["codes"]["k2d"]
    k, n, m = spec["k"], spec["n"], spec["m"]
    lib = C.CDLL(_capi.LIB_PATH)
    lib.mvd_last_error.restype = C.c_char_p
    ctx = C.c_void_p()
    assert lib.mvd_create(C.byref(ctx), 0) == 0
    env = dict(vm=vm, np=np, lib=lib, ctx=ctx, k=k, n=n, m=m, gen1=spec["gen"], gen2=espec["gen"],
               ptr=lambda a: a.ctypes.data_as(C.c_void_p))
    exec(compile(snippet, "INTEGRATION.md#tables", "exec"), env)
    with Detector(spec["gen"], k, n, m) as det:
        T = bitsource.bsc_threshold(0.1)
        counts = det.learn_counts([Seg(N=20000, threshold=T, stream=bitsource.LEARN_STREAM, enc_taps=det.dec_taps)], burn=200, seed=3)[0]
        P1 = codes.p1_from_edge_counts(det.table, counts, 1.0)
        det.set_models([P1])
        segs = [Seg(N=120, threshold=T, stream=h, enc_taps=det.taps_of((spec, espec)[h]["gen"]), decide=h, trial_begin=0, trial_end=2000)
                for h in (0, 1)]
        want = det.detect(segs, seed=3)
        met = np.ascontiguousarray(det.table.metrics, dtype=np.uint8)
        nxt = np.ascontiguousarray(det.table.nxt, dtype=np.uint32)
        assert lib.mvd_set_states(ctx, det.table.S, env["ptr"](met), env["ptr"](nxt)) == 0, lib.mvd_last_error(ctx)
        assert lib.mvd_set_loglik(ctx, 1, env["ptr"](det.logP1), env["ptr"](det.logTref)) == 0
        raw = (_capi.Segment * 2)()
        for h in (0, 1):
            raw[h].N, raw[h].threshold, raw[h].stream, raw[h].table, raw[h].decide, raw[h].random_input = 120, T, h, 0, h, 1
            raw[h].enc_taps[0] = h                                  # encoder INDEX: 0 = gen1, 1 = gen2
            raw[h].trial_begin, raw[h].trial_end = 0, 2000
        src = _capi.Src()
        src.mode, src.seed = _capi.SRC_PHILOX, 3
        tallies = np.zeros(2, dtype=np.uint64)
        assert lib.mvd_detect(ctx, C.byref(src), raw, 2, 0, env["ptr"](tallies), None, None) == 0, lib.mvd_last_error(ctx)
        assert tallies.tolist() == want.tolist()
    lib.mvd_destroy(ctx)
