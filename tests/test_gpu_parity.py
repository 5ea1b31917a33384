"""GPU parity tests: every CUDA path against golden vectors (from the reference's own functions)
and against the CPU oracle on identical bit streams.  Integer results are compared bit-exactly;
float64 log-likelihoods bit-exactly too (tolerance stated where it is looser).  All calls go through
the C ABI (ctypes -> libmvd.so)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ENGINES = ["acs", "fsm"]


def _taps(spec):
    from mvd import codes
    return codes.tap_masks(codes.freeze_generator(spec["gen"]), spec["m"], spec["k"])


@pytest.fixture(scope="module")
def dets(codes_spec):
    from mvd.engine import Detector
    cache = {}

    def get(name):
        if name not in cache:
            s = codes_spec[name]
            cache[name] = Detector(s["gen"], s["k"], s["n"], s["m"])
        return cache[name]

    yield get
    for d in cache.values():
        d.close()


SIM_CASES = [("c75_self", "c75", "c75"), ("c75_vs65", "c75", "c65"), ("c75_p001", "c75", "c75"),
             ("c75_p05", "c75", "c75"), ("m3_self", "m3a", "m3a"), ("m3_vs", "m3a", "m3b"),
             ("r13_self", "r13", "r13"), ("m1_self", "m1", "m1")]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("case,dec,enc", SIM_CASES)
def test_trace_bitstream_matches_reference(golden, codes_spec, dets, case, dec, enc, engine):
    """Host-supplied bits -> relative metrics identical to the reference's step function."""
    from mvd import bitsource
    from mvd.engine import Seg
    g = golden["sim_kats"][case]
    det = dets(dec)
    N, n = g["N"], codes_spec[dec]["n"]
    u = np.array(g["u_bits"], dtype=np.uint8).reshape(1, N)
    e = np.array(g["e_bits"], dtype=np.uint8).reshape(N, n).T.reshape(1, n, N)
    seg = Seg(N=N, enc_taps=_taps(codes_spec[enc]), trial_begin=0, trial_end=1)
    idx, met = det.trace(seg, bits=bitsource.pack_bitstreams(u, e), engine=engine)
    want = np.array(g["metrics"], dtype=np.uint8)
    assert np.array_equal(met[0], want)
    states = {tuple(r): i for i, r in enumerate(det.table.metrics.tolist())}
    assert idx[0].tolist() == [states[tuple(r)] for r in g["metrics"]]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("case,dec,enc", SIM_CASES)
def test_trace_philox_matches_reference(golden, codes_spec, dets, case, dec, enc, engine):
    """On-device MVD-PHILOX-2 bits + encoder + BSC + recursion == golden trajectory."""
    from mvd import bitsource
    from mvd.engine import Seg
    g = golden["sim_kats"][case]
    det = dets(dec)
    seg = Seg(N=g["N"], threshold=bitsource.bsc_threshold(g["p"]), stream=g["stream"], enc_taps=_taps(codes_spec[enc]),
              trial_begin=g["trial"], trial_end=g["trial"] + 1)
    idx, met = det.trace(seg, seed=g["seed"], engine=engine)
    assert np.array_equal(met[0], np.array(g["metrics"], dtype=np.uint8))


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("N", [1, 31, 32, 33, 127, 128, 129, 300])
def test_trace_lengths_vs_oracle(codes_spec, dets, engine, N):
    """Ragged block / superblock boundaries, several trials per launch, vs the C oracle."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    det = dets("c75")
    T = bitsource.bsc_threshold(0.07)
    tab = co.Table(det.table.metrics, 2)
    seg = Seg(N=N, threshold=T, stream=9, enc_taps=_taps(codes_spec["c65"]), trial_begin=1000, trial_end=1000 + 300)
    idx, met = det.trace(seg, seed=77, engine=engine)
    for t in (0, 1, 255, 256, 299):
        U, E = co.trial_words(77, 9, 1000 + t, N, 2, T)
        oi, _, om = co.simulate(_taps(codes_spec["c75"]), _taps(codes_spec["c65"]), 2, 2, N, U, E, tab, True)
        assert np.array_equal(idx[t], oi.astype(np.uint32))
        assert np.array_equal(met[t], om)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", ["c75", "m3a", "r13"])
def test_learn_counts_vs_oracle(codes_spec, dets, engine, name):
    """Transition counts of a learning chain: bit-exact vs the oracle (Pd_plotter.py:158-163)."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    s = codes_spec[name]
    det = dets(name)
    tab = co.Table(det.table.metrics, s["m"])
    taps = _taps(s)
    L, burn = 20000, 200
    ps = [0.01, 0.2]
    segs = [Seg(N=L, threshold=bitsource.bsc_threshold(p), stream=bitsource.LEARN_STREAM, enc_taps=taps) for p in ps]
    got = det.learn_counts(segs, burn=burn, seed=123, engine=engine)
    for i, p in enumerate(ps):
        want, _ = co.learn_chain(taps, taps, s["n"], s["m"], L, burn, bitsource.bsc_threshold(p), 123,
                                 bitsource.LEARN_STREAM, 0, tab)
        assert np.array_equal(got[i], want)
        assert int(got[i].sum()) == L - burn


@pytest.mark.parametrize("name,L,burn,p", [("c75", 6200, 200, 0.1), ("c75", 6200, 200, 0.5), ("c75", 100, 0, 0.3),
                                           ("c75", 129, 200, 0.3), ("c75", 4097, 33, 0.001), ("m3a", 87000, 200, 0.05),
                                           ("r13", 5000, 100, 0.2), ("m1", 777, 7, 0.4), ("c65", 3000, 0, 0.0)])
@pytest.mark.parametrize("warm", [128, 0, 32])
def test_learn_chunk_parallel_vs_oracle(codes_spec, dets, name, L, burn, p, warm):
    """Chunk-parallel learning chain (speculate / check / fix) == the sequential chain of the oracle,
    bit-exact, with the default warm-up, with none (every chunk repaired) and with a short one."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    s = codes_spec[name]
    det = dets(name)
    tab = co.Table(det.table.metrics, s["m"])
    taps = _taps(s)
    T = bitsource.bsc_threshold(p)
    seg = Seg(N=L, threshold=T, stream=bitsource.LEARN_STREAM, enc_taps=taps)
    want, _ = co.learn_chain(taps, taps, s["n"], s["m"], L, burn, T, 321, bitsource.LEARN_STREAM, 0, tab)
    det.learn_warm(warm)
    try:
        got = det.learn_counts([seg, seg], burn=burn, seed=321, engine="fsm")
        assert det.last_kernel_kind() == 1024
        dirty = det.learn_dirty_chunks()
    finally:
        det.learn_warm(128)
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)
    assert int(got[0].sum()) == max(0, L - burn)
    if warm == 128 and name != "m3a":
        assert dirty == 0
    det.force_generic(True)
    try:
        ser = det.learn_counts([seg], burn=burn, seed=321, engine="fsm")[0]
        assert det.last_kernel_kind() == 0
    finally:
        det.force_generic(False)
    assert np.array_equal(ser, want)


@pytest.mark.parametrize("chunk", [256, 512, 1024])
@pytest.mark.parametrize("name,L,burn,p,warm", [("c75", 6200, 200, 0.1, 128), ("m3a", 87000, 200, 0.05, 128), ("m3a", 20011, 33, 0.3, 32),
                                                ("c75", 4097, 0, 0.5, 0)])
def test_learn_chunk_lengths_vs_oracle(codes_spec, dets, monkeypatch, name, L, burn, p, warm, chunk):
    """The longer chunks long chains are cut into (256 / 512 / 1 024 steps, forced here through MVD_LEARN_CHUNK; chosen per call
    otherwise) give the sequential chain's counts bit for bit -- with the default warm-up, a short one and none (every chunk
    repaired by the fix pass), ragged last chunks."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    s = codes_spec[name]
    det = dets(name)
    tab = co.Table(det.table.metrics, s["m"])
    taps = _taps(s)
    T = bitsource.bsc_threshold(p)
    seg = Seg(N=L, threshold=T, stream=bitsource.LEARN_STREAM, enc_taps=taps)
    want, _ = co.learn_chain(taps, taps, s["n"], s["m"], L, burn, T, 77, bitsource.LEARN_STREAM, 0, tab)
    monkeypatch.setenv("MVD_LEARN_CHUNK", str(chunk))
    det.learn_warm(warm)
    try:
        got = det.learn_counts([seg, seg], burn=burn, seed=77, engine="fsm")
        assert det.last_kernel_kind() == 1024
    finally:
        det.learn_warm(128)
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)


def test_learn_counts_many_chains(codes_spec, dets):
    """Several chains per segment aggregate into one histogram (shared-memory atomics)."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    det = dets("c75")
    tab = co.Table(det.table.metrics, 2)
    taps = _taps(codes_spec["c75"])
    T = bitsource.bsc_threshold(0.1)
    seg = Seg(N=700, threshold=T, stream=3, enc_taps=taps, trial_begin=5, trial_end=5 + 600)
    for engine in ENGINES:
        got = det.learn_counts([seg], burn=100, seed=9, engine=engine)[0]
        want = np.zeros_like(got)
        for t in range(5, 605):
            w, _ = co.learn_chain(taps, taps, 2, 2, 700, 100, T, 9, 3, t, tab)
            want += w
        assert np.array_equal(got, want)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("exp", ["c75_c65_small", "c75_c65_lap", "m3_small", "c75_c65_n1000"])
def test_run_experiment_matches_reference(golden, exp, engine):
    """Drop-in run_experiment == the reference's run_experiment (unmodified, injected simulator):
    identical Pd / Pc, P1 and per-trial log-likelihoods."""
    import Pd_plotter as pdp
    from mvd import codes
    g = golden["experiments"][exp]
    details = {}
    df = pdp.run_experiment(g["k"], g["n"], g["m"], g["gen1"], g["gen2"], g["num_iter"], g["p_vec"], g["learn_len"],
                            g["learn_burn"], g["laplace"], g["seed"], N_spectrum=g["N_list"], engine=engine,
                            details=details)
    assert df.to_dict(orient="records") == g["rows"]
    assert df.to_csv(index=False) == g["csv"]
    for i, p in enumerate(details["distinct_p"]):
        want = np.array(g["P1_edge"][repr(p)]["edge"])
        got = details["p1_tables"][i]
        if g["laplace"] == 1.0:
            assert np.array_equal(got, want)                       # bit-exact
        else:
            np.testing.assert_allclose(got, want, rtol=1e-12)      # north star: 1e-6
    # per-trial log-likelihood pairs, in the reference's call order (logp1, ref, logp2, ref)
    import viterbi_markov as vm
    from mvd import bitsource
    from mvd.engine import Seg
    det = vm._detector(codes.freeze_generator(g["gen1"]), g["k"], g["n"], g["m"], 0)
    logs = np.array(g["logps"]).reshape(len(g["N_list"]) * len(g["p_vec"]), g["num_iter"], 2, 2)
    q = 0
    tindex = {p: i for i, p in enumerate(details["distinct_p"])}
    for N in g["N_list"]:
        for p in g["p_vec"]:
            for h, gen in enumerate((g["gen1"], g["gen2"])):
                seg = Seg(N=N, threshold=bitsource.bsc_threshold(p), stream=2 * q + h, table=tindex[p],
                          enc_taps=det.taps_of(gen), decide=h, trial_begin=0, trial_end=g["num_iter"])
                _, lp = det.detect([seg], seed=g["seed"], engine=engine, want_logp=True)
                np.testing.assert_allclose(lp, logs[q, :, h, :], rtol=1e-6)     # north-star tolerance
                assert np.array_equal(lp, logs[q, :, h, :])                      # and in fact bit-exact
            q += 1


def _oracle_models(det, spec, p, learn_len, seed):
    import c_oracle as co
    from mvd import bitsource, codes
    taps = _taps(spec)
    tab = co.Table(det.table.metrics, spec["m"])
    edge, _ = co.learn_chain(taps, taps, spec["n"], spec["m"], learn_len, 200, bitsource.bsc_threshold(p), seed,
                             bitsource.LEARN_STREAM, 0, tab)
    P1 = codes.p1_from_edge_counts(det.table, edge, 1.0)
    return tab, P1, codes.tref_half_table(det.table)


@pytest.mark.parametrize("path", ["fast", "fastg", "fast1", "generic"])
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("dec,enc,N,p", [("c75", "c65", 500, 0.1), ("c75", "c75", 200, 0.3), ("m3a", "m3b", 333, 0.05),
                                         ("r13", "r13", 97, 0.15), ("m1", "m1", 64, 0.2), ("c75", "c65", 7, 0.4),
                                         ("c75", "c65", 129, 0.001), ("m3a", "m3a", 40, 0.5), ("c65", "c75", 300, 0.1)])
def test_detect_vs_oracle(codes_spec, dets, engine, dec, enc, N, p, path):
    """Tallies and per-trial float64 sums vs the oracle, 3000 trials, both decision rules; through
    the fast kernels (mvd_detect2.cuh: n = 2 codes with shared-memory tables) and the generic ones.
    "fastg" = the pair kernel with the general branch-metric table where the complement-label short cut
    would apply ((7,5) has both end taps set in both generators, (6,5) has not and always takes it)."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    spec = codes_spec[dec]
    det = dets(dec)
    tab, P1, Tref = _oracle_models(det, spec, p, 8000, 5)
    det.set_models([P1])
    T = bitsource.bsc_threshold(p)
    ntr = 3000
    segs = [Seg(N=N, threshold=T, stream=10 + d, enc_taps=_taps(codes_spec[enc]), decide=d, trial_begin=17,
                trial_end=17 + ntr) for d in (0, 1)]
    det.force_generic(path == "generic")
    det.no_pair(1 if path == "fast1" else 2)
    det.no_fsm1(path == "fast1")
    det.no_antipodal(path == "fastg")
    try:
        tallies, lp = det.detect(segs, seed=2024, engine=engine, want_logp=True)
        kind = det.last_kernel_kind()
    finally:
        det.force_generic(False)
        det.no_pair(False)
        det.no_fsm1(False)
        det.no_antipodal(False)
    if path == "fastg":
        path = "fast"
    if path == "generic" or (spec["n"] != 2 and engine == "acs"):
        assert kind == 0
    elif spec["n"] != 2:
        # rate 1/3 on the NEXT-walk engines (run_trial_n3): the one-load walk when log T(1/2) packs, else the two-load one
        assert kind != 0 and (kind - 1) % 16 in ((2,) if path == "fast1" else (2, 3))
    else:
        want_lookup = (2 if path == "fast1" else 3) if engine == "fsm" else (0 if spec["m"] <= 2 else 1)
        assert kind != 0 and (kind - 1) % 16 == want_lookup
        assert (kind >= 256) == (path == "fast" and engine == "acs" and spec["m"] in (2, 3))
    for d in (0, 1):
        want, wlp = co.run_trials(_taps(spec), _taps(codes_spec[enc]), spec["n"], spec["m"], N, T, 2024, 10 + d, 17,
                                  17 + ntr, tab, P1, Tref, d, want_logp=True)
        assert int(tallies[d]) == want
        assert np.array_equal(lp[d * ntr:(d + 1) * ntr], wlp)


M2_DECODERS = [(g0, g1) for g0 in range(1, 8) for g1 in range(1, 8) if ((g0 | g1) & 4) and (g0, g1) not in ((5, 5), (7, 7))]


@pytest.mark.parametrize("g0,g1", M2_DECODERS)
def test_pair_kernel_every_m2_decoder_vs_oracle(g0, g1):
    """The two-trials-per-thread kernel over EVERY rate-1/2 memory-2 decoder (S = 3 .. 65 Markov states): the
    host-searched offset-invariant state-table key, the difference-form butterfly of complement-label decoders and the
    general table of the others, the 32 KB-aligned shared-memory layout at S = 65 (two blocks per SM), a ragged N
    (4 blocks + 3 steps) -- tallies and per-trial float64 sums against the oracle, bit for bit."""
    import c_oracle as co
    from mvd import bitsource, codes
    from mvd.engine import Detector, Seg
    bits = lambda g: [(g >> 2) & 1, (g >> 1) & 1, g & 1]
    gen, enc_gen = [[bits(g0)], [bits(g1)]], [[[1, 1, 1]], [[1, 0, 1]]]
    p, N, ntr, seed = 0.08, 131, 700, 99
    T = bitsource.bsc_threshold(p)
    with Detector(gen, 1, 2, 2) as det:
        taps, etaps = det.taps_of(gen), det.taps_of(enc_gen)
        tab = co.Table(det.table.metrics, 2)
        edge, _ = co.learn_chain(taps, taps, 2, 2, 6000, 200, T, seed, bitsource.LEARN_STREAM, 0, tab)
        P1 = codes.p1_from_edge_counts(det.table, edge, 1.0)
        Tref = codes.tref_half_table(det.table)
        det.set_models([P1])
        segs = [Seg(N=N, threshold=T, stream=3 + d, enc_taps=(taps, etaps)[d], decide=d, trial_begin=5, trial_end=5 + ntr)
                for d in (0, 1)]
        det.no_pair(2)
        tallies, lp = det.detect(segs, seed=seed, engine="acs", want_logp=True)
        kind = det.last_kernel_kind()
        if int(np.asarray(det.table.metrics).max()) <= 3:
            assert kind >= 256, "the pair kernel did not run"
        for d in (0, 1):
            want, wlp = co.run_trials(taps, (taps, etaps)[d], 2, 2, N, T, seed, 3 + d, 5, 5 + ntr, tab, P1, Tref, d, want_logp=True)
            assert int(tallies[d]) == want
            assert np.array_equal(lp[d * ntr:(d + 1) * ntr], wlp)


def test_m3_pair_kernel_full_blocks_vs_oracle(codes_spec):
    """The memory-3 pair kernel at its production geometry: enough trials that the dispatcher keeps blocks of 768 threads
    (one per SM, 204 KB of shared tables) instead of shrinking them, last block ragged, N = 37 (one 32-step block + 5
    steps) -- tallies and every per-trial float64 sum against the oracle."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Detector, Seg
    spec = codes_spec["m3a"]
    with Detector(spec["gen"], 1, 2, 3) as det:
        tab, P1, Tref = _oracle_models(det, spec, 0.05, 20000, 3)
        det.set_models([P1])
        T = bitsource.bsc_threshold(0.05)
        taps, taps2 = _taps(spec), _taps(codes_spec["m3b"])
        ntr = 160 * 2 * 768 + 301
        seg = Seg(N=37, threshold=T, stream=1, enc_taps=taps2, decide=1, trial_begin=3, trial_end=3 + ntr)
        t, lp = det.detect([seg], seed=8, engine="acs", want_logp=True)
        assert det.last_kernel_kind() & 256, "detect3p_kernel did not run"
        want, wlp = co.run_trials(taps, taps2, 2, 3, 37, T, 8, 1, 3, 3 + ntr, tab, P1, Tref, 1, want_logp=True)
        assert int(t[0]) == want and np.array_equal(lp, wlp)


@pytest.mark.parametrize("g0,g1,S,lls", [(11, 15, 435, 6), (12, 14, 165, 7), (9, 13, 665, 5)])
def test_next_walk_two_big_blocks_vs_oracle(g0, g1, S, lls):
    """The one-load NEXT walk at its large-launch geometry: two blocks of 768 threads per SM around a table with twice the
    copies three blocks of 512 could hold (4, 8 and 2 copies for these three memory-3 decoders) -- enough trials that the
    dispatcher takes it, last block ragged, N = 37; tallies and every per-trial float64 sum against the oracle."""
    import c_oracle as co
    from mvd import bitsource, codes
    from mvd.engine import Detector, Seg
    bits = lambda g: [(g >> 3) & 1, (g >> 2) & 1, (g >> 1) & 1, g & 1]
    gen, enc_gen = [[bits(g0)], [bits(g1)]], [[[1, 1, 1, 1]], [[1, 0, 1, 1]]]
    T, seed = bitsource.bsc_threshold(0.07), 5
    with Detector(gen, 1, 2, 3) as det:
        assert det.S == S
        taps, etaps = det.taps_of(gen), det.taps_of(enc_gen)
        tab = co.Table(det.table.metrics, 3)
        edge, _ = co.learn_chain(taps, taps, 2, 3, 20000, 200, T, seed, bitsource.LEARN_STREAM, 0, tab)
        P1 = codes.p1_from_edge_counts(det.table, edge, 1.0)
        Tref = codes.tref_half_table(det.table)
        det.set_models([P1])
        ntr = 150 * 2 * 768 + 301
        seg = Seg(N=37, threshold=T, stream=2, enc_taps=etaps, decide=1, trial_begin=9, trial_end=9 + ntr)
        t, lp = det.detect([seg], seed=seed, engine="fsm", want_logp=True)
        kind = det.last_kernel_kind()
        assert (kind - 1) % 16 == 3 and (kind - 1) // 16 == lls, kind       # one-load walk, log2 of the row stride
        want, wlp = co.run_trials(taps, etaps, 2, 3, 37, T, seed, 2, 9, 9 + ntr, tab, P1, Tref, 1, want_logp=True)
        assert int(t[0]) == want and np.array_equal(lp, wlp)


M3_DECODERS = [(g0, g1) for g0 in range(9, 16) for g1 in range(9, 16) if g0 != g1]


@pytest.mark.parametrize("g0,g1", M3_DECODERS)
def test_pair_kernel_every_m3_decoder_vs_oracle(g0, g1):
    """The two-trials-per-thread kernel of memory-3 codes (mvd_detect3p.cuh) over every rate-1/2 decoder whose two
    generators both reach back three steps (S = 5 .. 987 Markov states; those with S <= 435 fit its shared-memory
    layout, the others take the one-trial kernel): offset-invariant perfect-hash key, Eq. 5 once per 32-step block,
    byte-row butterflies of the complement-label decoders and selector picks of the others, a full block of 768
    pair-threads plus a partly filled one, a ragged N (4 blocks + 3 steps) -- tallies and per-trial float64 sums
    against the oracle, bit for bit."""
    import c_oracle as co
    from mvd import bitsource, codes
    from mvd.engine import Detector, Seg
    bits = lambda g: [(g >> 3) & 1, (g >> 2) & 1, (g >> 1) & 1, g & 1]
    gen, enc_gen = [[bits(g0)], [bits(g1)]], [[[1, 1, 1, 1]], [[1, 0, 1, 1]]]
    p, N, ntr, seed = 0.08, 131, 1700, 99
    T = bitsource.bsc_threshold(p)
    with Detector(gen, 1, 2, 3) as det:
        taps, etaps = det.taps_of(gen), det.taps_of(enc_gen)
        tab = co.Table(det.table.metrics, 3)
        edge, _ = co.learn_chain(taps, taps, 2, 3, 6000, 200, T, seed, bitsource.LEARN_STREAM, 0, tab)
        P1 = codes.p1_from_edge_counts(det.table, edge, 1.0)
        Tref = codes.tref_half_table(det.table)
        det.set_models([P1])
        segs = [Seg(N=N, threshold=T, stream=3 + d, enc_taps=(taps, etaps)[d], decide=d, trial_begin=5, trial_end=5 + ntr)
                for d in (0, 1)]
        det.no_pair(2)
        tallies, lp = det.detect(segs, seed=seed, engine="acs", want_logp=True)
        kind = det.last_kernel_kind()
        assert kind != 0, "generic kernel ran instead of a fast one"
        if det.S <= 435:
            assert kind >= 256, "the pair kernel did not run"
        for d in (0, 1):
            want, wlp = co.run_trials(taps, (taps, etaps)[d], 2, 3, N, T, seed, 3 + d, 5, 5 + ntr, tab, P1, Tref, d, want_logp=True)
            assert int(tallies[d]) == want
            assert np.array_equal(lp[d * ntr:(d + 1) * ntr], wlp)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("dec,enc", [("c75", "c65"), ("m3a", "m3b")])
def test_detect_against_numeric_T_of_any_p(codes_spec, dets, engine, dec, enc):
    """scoring against T(p_ref) for p_ref != 1/2 (sympy-free numeric T(p); log T is no
    longer a multiple of log 1/2, so the NEXT walk takes its two-load form) -- bit-exact vs the oracle."""
    import c_oracle as co
    from mvd import bitsource, codes
    from mvd.engine import Seg
    spec = codes_spec[dec]
    det = dets(dec)
    tab, P1, _ = _oracle_models(det, spec, 0.1, 8000, 5)
    Tref = codes.t_edge_table(det.table, 0.3)
    assert not np.array_equal(Tref, codes.tref_half_table(det.table))
    det.set_models([P1], Tref)
    T = bitsource.bsc_threshold(0.1)
    ntr = 2000
    segs = [Seg(N=211, threshold=T, stream=20 + d, enc_taps=_taps(codes_spec[enc]), decide=d, trial_begin=0,
                trial_end=ntr) for d in (0, 1)]
    tallies, lp = det.detect(segs, seed=99, engine=engine, want_logp=True)
    for d in (0, 1):
        want, wlp = co.run_trials(_taps(spec), _taps(codes_spec[enc]), spec["n"], spec["m"], 211, T, 99, 20 + d, 0,
                                  ntr, tab, P1, Tref, d, want_logp=True)
        assert int(tallies[d]) == want
        assert np.array_equal(lp[d * ntr:(d + 1) * ntr], wlp)


@pytest.mark.parametrize("pair", [0, 2])
@pytest.mark.parametrize("dec,enc", [("c75", "c65"), ("c65", "c75"), ("m3a", "m3b")])
@pytest.mark.parametrize("engine", ENGINES)
def test_detect_fast_bitstream_vs_generic(codes_spec, dets, engine, dec, enc, pair):
    """Host-supplied bit streams through the fast kernels == generic kernels == oracle sums; pair = 2 forces the
    two-trials-per-thread kernels (their bit-stream instances: complement-label and general table at m = 2, m = 3)."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    det = dets(dec)
    spec = codes_spec[dec]
    tab, P1, Tref = _oracle_models(det, spec, 0.1, 6200, 123)
    det.set_models([P1])
    rng = np.random.default_rng(11)
    ntr, N = 777, 301
    u = rng.integers(0, 2, (ntr, N), dtype=np.uint8)
    e = (rng.random((ntr, 2, N)) < 0.12).astype(np.uint8)
    bits = bitsource.pack_bitstreams(u, e)
    taps2 = _taps(codes_spec[enc])
    seg = Seg(N=N, enc_taps=taps2, decide=1, trial_begin=0, trial_end=ntr)
    det.no_pair(pair)
    try:
        t_fast, lp_fast = det.detect([seg], bits=bits, engine=engine, want_logp=True)
        kind = det.last_kernel_kind()
    finally:
        det.no_pair(0)
    assert kind != 0
    assert (kind >= 256) == (pair == 2 and engine == "acs")
    det.force_generic(True)
    try:
        t_gen, lp_gen = det.detect([seg], bits=bits, engine=engine, want_logp=True)
        assert det.last_kernel_kind() == 0
    finally:
        det.force_generic(False)
    assert int(t_fast[0]) == int(t_gen[0]) and np.array_equal(lp_fast, lp_gen)
    # oracle on the same bits (first 40 trials)
    for t in range(40):
        U = bitsource.bits_to_words(u[t])[:(N + 31) // 32]
        E = bitsource.bits_to_words(e[t])[:, :(N + 31) // 32]
        idx, rseq, _ = co.simulate(_taps(spec), taps2, 2, spec["m"], N, U, E, tab)
        assert co.log_prob(idx, rseq, N, 2, P1) == lp_fast[t, 0]
        assert co.log_prob(idx, rseq, N, 2, Tref) == lp_fast[t, 1]


@pytest.mark.parametrize("no_fsm1", [False, True])
def test_rate13_fast_walk_bitstream(codes_spec, dets, no_fsm1):
    """Rate 1/3 (n = 3, R = 8 received words) through the fast NEXT-walk kernels (run_trial_n3) on host-supplied bit streams and
    on the device bit source: == the generic kernels == the oracle, ragged N, both walk variants."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    spec = codes_spec["r13"]
    det = dets("r13")
    tab, P1, Tref = _oracle_models(det, spec, 0.12, 8000, 7)
    det.set_models([P1])
    rng = np.random.default_rng(5)
    ntr, N = 600, 203
    u = rng.integers(0, 2, (ntr, N), dtype=np.uint8)
    e = (rng.random((ntr, 3, N)) < 0.12).astype(np.uint8)
    bits = bitsource.pack_bitstreams(u, e)
    taps = _taps(spec)
    seg = Seg(N=N, enc_taps=taps, decide=1, trial_begin=0, trial_end=ntr)
    det.no_fsm1(no_fsm1)
    try:
        t_fast, lp_fast = det.detect([seg], bits=bits, engine="fsm", want_logp=True)
        kind = det.last_kernel_kind()
        T = bitsource.bsc_threshold(0.12)
        pseg = Seg(N=N, threshold=T, stream=4, enc_taps=taps, decide=0, trial_begin=9, trial_end=9 + ntr)
        t_ph, lp_ph = det.detect([pseg], seed=77, engine="fsm", want_logp=True)
        assert det.last_kernel_kind() == kind
    finally:
        det.no_fsm1(False)
    assert kind != 0 and (kind - 1) % 16 in ((2,) if no_fsm1 else (2, 3))
    det.force_generic(True)
    try:
        t_gen, lp_gen = det.detect([seg], bits=bits, engine="fsm", want_logp=True)
        assert det.last_kernel_kind() == 0
    finally:
        det.force_generic(False)
    assert int(t_fast[0]) == int(t_gen[0]) and np.array_equal(lp_fast, lp_gen)
    for t in range(30):
        U = bitsource.bits_to_words(u[t])[:(N + 31) // 32]
        E = bitsource.bits_to_words(e[t])[:, :(N + 31) // 32]
        idx, rseq, _ = co.simulate(taps, taps, 3, spec["m"], N, U, E, tab)
        assert co.log_prob(idx, rseq, N, 3, P1) == lp_fast[t, 0]
        assert co.log_prob(idx, rseq, N, 3, Tref) == lp_fast[t, 1]
    want, wlp = co.run_trials(taps, taps, 3, spec["m"], N, T, 77, 4, 9, 9 + ntr, tab, P1, Tref, 0, want_logp=True)
    assert int(t_ph[0]) == want and np.array_equal(lp_ph, wlp)


@pytest.mark.parametrize("warm", [128, 0])
@pytest.mark.parametrize("dec,enc,Ns,p", [("c75", "c65", (1500, 3001), 0.1), ("c75", "c75", (513, 512), 0.3),
                                          ("m3a", "m3b", (2050, 700), 0.05), ("m1", "m1", (4096, 33), 0.2),
                                          ("r13", "r13", (1100, 600), 0.15)])
def test_split_long_trials_vs_oracle(codes_spec, dets, dec, enc, Ns, p, warm):
    """Few long trials split along the time axis (mvd_split.cuh: chunk walk / fix / in-order scoring): tallies
    and per-trial float64 sums bit-identical to the oracle and to the one-thread-per-trial kernels, ragged
    lengths, two segments of different N, with the default warm-up and with none (every chunk repaired)."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    spec = codes_spec[dec]
    det = dets(dec)
    tab, P1, Tref = _oracle_models(det, spec, p, 8000, 5)
    det.set_models([P1])
    T = bitsource.bsc_threshold(p)
    ntr = 37
    segs = [Seg(N=Ns[d], threshold=T, stream=30 + d, enc_taps=_taps(codes_spec[enc]), decide=d, trial_begin=5,
                trial_end=5 + ntr) for d in (0, 1)]
    det.split_trials(2)
    plain_t, plain_lp = det.detect(segs, seed=77, engine="fsm", want_logp=True)
    assert (det.last_kernel_kind() & 16384) == 0
    det.split_trials(1)
    det.learn_warm(warm)
    try:
        tallies, lp = det.detect(segs, seed=77, engine="fsm", want_logp=True)
        assert (det.last_kernel_kind() & 16384) != 0
        dirty = det.learn_dirty_chunks()
        subs, seq = det.split_stats()
        assert subs == ntr * sum(-(-N // 128) for N in Ns)
        if max(Ns) >= 4096:
            assert seq < subs, "no sub-chunk took the re-associated sums"
        det.split_sequential(True)                                # every term added one by one, in step order
        seq_t, seq_lp = det.detect(segs, seed=77, engine="fsm", want_logp=True)
        assert (det.last_kernel_kind() & 16384) != 0
        assert det.learn_dirty_chunks() == dirty
        assert det.split_stats() == (subs, subs)
        det.split_sequential(2)                                   # re-association without the class counting of the second sum
        gt, glp = det.detect(segs, seed=77, engine="fsm", want_logp=True)
        assert np.array_equal(gt, tallies) and np.array_equal(glp, lp)
        assert det.split_stats()[0] == subs
        det.split_sequential(False)
        for chunk in (256, 512, 1024):                            # steps per walker thread (chosen per call otherwise)
            det.split_chunk(chunk)
            ct, clp = det.detect(segs, seed=77, engine="fsm", want_logp=True)
            assert np.array_equal(ct, tallies) and np.array_equal(clp, lp), chunk
            assert det.split_stats()[0] == subs        # (the share taken term by term may differ by a sub-chunk: the estimates of a repaired chunk come from its speculated trajectory)
    finally:
        det.split_chunk(0)
        det.split_sequential(False)
        det.split_trials(0)
        det.learn_warm(128)
    assert np.array_equal(tallies, plain_t) and np.array_equal(lp, plain_lp)
    assert np.array_equal(tallies, seq_t) and np.array_equal(lp, seq_lp)
    for d in (0, 1):
        want, wlp = co.run_trials(_taps(spec), _taps(codes_spec[enc]), spec["n"], spec["m"], Ns[d], T, 77, 30 + d, 5,
                                  5 + ntr, tab, P1, Tref, d, want_logp=True)
        assert int(tallies[d]) == want
        assert np.array_equal(lp[d * ntr:(d + 1) * ntr], wlp)
    if warm == 0 and max(Ns) > 1024:
        assert dirty > 0


def test_split_is_automatic_for_few_long_trials(codes_spec, dets):
    """Automatic choice: 64 trials of N = 20 000 take the split path, 50 000 trials of N = 500 do not; m = 4
    (NEXT table in global memory) splits too and agrees with the unsplit run."""
    from mvd import bitsource
    from mvd.engine import Detector, Seg
    det = dets("c75")
    tab, P1, Tref = _oracle_models(det, codes_spec["c75"], 0.1, 8000, 5)
    det.set_models([P1])
    T = bitsource.bsc_threshold(0.1)
    long_seg = [Seg(N=20000, threshold=T, stream=d, enc_taps=_taps(codes_spec["c65"]), decide=d, trial_begin=0, trial_end=64) for d in (0, 1)]
    t1, lp1 = det.detect(long_seg, seed=3, engine="auto", want_logp=True)
    assert (det.last_kernel_kind() & 16384) != 0
    det.split_trials(2)
    try:
        t2, lp2 = det.detect(long_seg, seed=3, engine="auto", want_logp=True)
        assert (det.last_kernel_kind() & 16384) == 0
    finally:
        det.split_trials(0)
    assert np.array_equal(t1, t2) and np.array_equal(lp1, lp2)
    det.detect([Seg(N=500, threshold=T, stream=0, decide=0, trial_begin=0, trial_end=50000)], seed=3, engine="auto")
    assert (det.last_kernel_kind() & 16384) == 0
    spec = codes_spec["m4a"]
    with Detector(spec["gen"], 1, 2, 4, enumerate_with="gpu", max_states=1 << 16) as d4:
        _, P1, _ = _oracle_models(d4, spec, 0.05, 60000, 3)
        d4.set_models([P1])
        seg = [Seg(N=6000, threshold=bitsource.bsc_threshold(0.05), stream=1, enc_taps=_taps(codes_spec["m4b"]), decide=1,
                   trial_begin=0, trial_end=40)]
        a, la = d4.detect(seg, seed=8, engine="fsm", want_logp=True)
        assert (d4.last_kernel_kind() & 16384) != 0
        d4.split_trials(2)
        b, lb = d4.detect(seg, seed=8, engine="fsm", want_logp=True)
        assert np.array_equal(a, b) and np.array_equal(la, lb)


def test_detect_rules_are_complementary(codes_spec, dets):
    """On the *same* stream the H1 rule (>) and the H2 rule (<=) partition the trials."""
    from mvd import bitsource
    from mvd.engine import Seg
    det = dets("c75")
    _, P1, _ = _oracle_models(det, codes_spec["c75"], 0.1, 6200, 123)
    det.set_models([P1])
    T = bitsource.bsc_threshold(0.1)
    segs = [Seg(N=100, threshold=T, stream=4, decide=d, trial_begin=0, trial_end=5000) for d in (0, 1)]
    for engine in ENGINES:
        t = det.detect(segs, seed=1, engine=engine)
        assert int(t[0]) + int(t[1]) == 5000


def test_m4_global_tables_vs_oracle(codes_spec):
    """m = 4 (S = 25 751): state / log tables leave shared memory (L2-resident path)."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Detector, Seg
    spec = codes_spec["m4a"]
    with Detector(spec["gen"], 1, 2, 4, enumerate_with="lib") as det:
        assert det.S == 25751
        om, on = co.enumerate_states(_taps(spec), 2, 4)
        assert np.array_equal(det.table.metrics, om) and np.array_equal(det.table.nxt, on)
        tab, P1, Tref = _oracle_models(det, spec, 0.05, 60000, 3)
        det.set_models([P1])
        T = bitsource.bsc_threshold(0.05)
        taps2 = _taps(codes_spec["m4b"])
        seg = Seg(N=150, threshold=T, stream=1, enc_taps=taps2, decide=1, trial_begin=0, trial_end=700)
        want, wlp = co.run_trials(_taps(spec), taps2, 2, 4, 150, T, 8, 1, 0, 700, tab, P1, Tref, 1, want_logp=True)
        for engine in ENGINES:
            for generic in (False, True):
                det.force_generic(generic)
                try:
                    t, lp = det.detect([seg], seed=8, engine=engine, want_logp=True)
                    kind = det.last_kernel_kind()
                finally:
                    det.force_generic(False)
                assert (kind == 0) == generic and (generic or kind & 512)      # fast path with tables in global memory
                assert int(t[0]) == want
                assert np.array_equal(lp, wlp)
        # ACS engine, two trials per thread with the perfect-hash lookup (mvd_detect3p.cuh, m = 4 variant)
        # (both instances: byte-row butterflies of a complement-label decoder -- (31,33) is one -- and the selector picks
        # every other decoder takes)
        for general in (False, True):
            det.no_pair(2)
            det.no_antipodal(general)
            try:
                t, lp = det.detect([seg], seed=8, engine="acs", want_logp=True)
                kind = det.last_kernel_kind()
            finally:
                det.no_pair(False)
                det.no_antipodal(False)
            assert kind & 256 and kind & 512
            assert int(t[0]) == want and np.array_equal(lp, wlp)
        # learning chain through the global-memory histogram
        got = det.learn_counts([Seg(N=60000, threshold=T, stream=bitsource.LEARN_STREAM, enc_taps=_taps(spec))],
                               burn=200, seed=3)[0]
        want_edge, _ = co.learn_chain(_taps(spec), _taps(spec), 2, 4, 60000, 200, T, 3, bitsource.LEARN_STREAM, 0, tab)
        assert np.array_equal(got, want_edge)


@pytest.mark.parametrize("chunk", [0, 1, 7, 64])
@pytest.mark.parametrize("name", ["c75", "c65", "m3a", "m3b", "r13", "m1"])
def test_gpu_bfs_matches_reference_order(golden, codes_spec, name, chunk):
    """enumerate_markov_states_allzero on the GPU -- state set, BFS *index order* and NEXT
    identical to the reference's (golden sha256 of repr(states) / repr(NEXT) from the reference run),
    for the default chunk and for tiny chunks that split BFS levels and force cross-chunk duplicates."""
    import hashlib
    from mvd.engine import Detector
    s, g = codes_spec[name], golden["code_kats"][name]
    with Detector(s["gen"], s["k"], s["n"], s["m"], enumerate_with="lib") as det:
        st = det._enumerate_gpu(1 << 12, install=True, chunk_parents=chunk)
        tab = det._fetch_states(st["S"])
        assert st["closed"] and st["S"] == g["S"] == st["frontier"]
        assert st["max_metric"] == g["max_metric"] == tab.max_metric
        assert st["candidates"] == g["S"] * (1 << s["n"])
        assert sum(st["levels"]) == g["S"] and st["levels"][0] == 1 and st["levels"][1] <= 1 << s["n"]
        states = [tuple(int(v) for v in row) for row in tab.metrics]
        assert hashlib.sha256(repr(states).encode()).hexdigest()[:16] == g["states_sha"]
        assert hashlib.sha256(repr(tab.nxt.tolist()).encode()).hexdigest()[:16] == g["next_sha"]
        assert np.array_equal(tab.mult, np.array(g["mult"], dtype=np.uint8))
        assert det.last_kernel_kind() == 2048


@pytest.mark.parametrize("name,S", [("m4a", 25751), ("m4c", 150743)])
def test_gpu_bfs_m4_vs_host_bfs(golden, codes_spec, name, S):
    """m = 4 (S = 25 751 and 150 743): GPU enumeration == the reference's own BFS (golden hashes of its
    state list and NEXT table) and == the C oracle's sequential BFS, index for index."""
    import hashlib
    import c_oracle as co
    from mvd.engine import Detector
    spec, g = codes_spec[name], golden["m4_kats"][name]
    om, on = co.enumerate_states(_taps(spec), 2, 4)
    assert om.shape[0] == S == g["S"]
    assert hashlib.sha256(repr([tuple(r) for r in om.tolist()]).encode()).hexdigest()[:16] == g["states_sha"]
    assert hashlib.sha256(repr(on.tolist()).encode()).hexdigest()[:16] == g["next_sha"]
    for chunk in (0, 1000):
        with Detector(spec["gen"], 1, 2, 4, enumerate_with="gpu", max_states=1 << 18) as det:
            if chunk:
                st = det._enumerate_gpu(1 << 18, install=True, chunk_parents=chunk)
                det.table = det._fetch_states(st["S"])
            assert det.S == S
            assert np.array_equal(det.table.metrics, om) and np.array_equal(det.table.nxt, on)


def test_m4_large_table_vs_oracle(codes_spec):
    """m = 4 with S = 150 743, the largest checkable configuration (the reference cannot run it: 182 GB of dense
    counts): learning-chain edge counts (4 x 10^5 steps, Pd_plotter.py:158-163), detection tallies and per-trial
    float64 sums (Pd_plotter.py:210-223) of the NEXT-table walk, of the two-trials-per-thread ACS kernel (perfect hash
    over 2^19 slots in global memory) and of the generic checked kernel, all bit-identical to the C oracle."""
    import c_oracle as co
    from mvd import bitsource, codes
    from mvd.engine import Detector, Seg
    spec = codes_spec["m4c"]
    with Detector(spec["gen"], 1, 2, 4, enumerate_with="gpu", max_states=1 << 18) as det:
        assert det.S == 150743
        tab = co.Table(det.table.metrics, 4)
        T = bitsource.bsc_threshold(0.05)
        L = 400000
        counts = det.learn_counts([Seg(N=L, threshold=T, stream=bitsource.LEARN_STREAM, enc_taps=det.dec_taps)], burn=200, seed=3)[0]
        assert det.last_kernel_kind() == 1024
        want_counts, _ = co.learn_chain(det.dec_taps, det.dec_taps, 2, 4, L, 200, T, 3, bitsource.LEARN_STREAM, 0, tab)
        assert np.array_equal(counts, want_counts) and int(counts.sum()) == L - 200
        P1 = codes.p1_from_edge_counts(det.table, counts, 1.0)
        Tref = codes.tref_half_table(det.table)
        det.set_models([P1])
        taps2 = [det.dec_taps[1], det.dec_taps[0]]
        ntr = 3000
        segs = [Seg(N=300, threshold=T, stream=d, enc_taps=(det.dec_taps if d == 0 else taps2), decide=d, trial_begin=0, trial_end=ntr)
                for d in (0, 1)]
        want = [co.run_trials(det.dec_taps, det.dec_taps if d == 0 else taps2, 2, 4, 300, T, 5, d, 0, ntr, tab, P1, Tref, d,
                              want_logp=True) for d in (0, 1)]
        a, la = det.detect(segs, seed=5, engine="fsm", want_logp=True)
        det.no_pair(2)
        try:
            b, lb = det.detect(segs, seed=5, engine="acs", want_logp=True)
            kind = det.last_kernel_kind()
        finally:
            det.no_pair(False)
        assert kind & 256 and kind & 512
        det.force_generic(True)
        try:
            c, lc = det.detect(segs, seed=5, engine="acs", want_logp=True)
            assert det.last_kernel_kind() == 0
        finally:
            det.force_generic(False)
        for t, lp in ((a, la), (b, lb), (c, lc)):
            for d in (0, 1):
                assert int(t[d]) == want[d][0]
                assert np.array_equal(lp[d * ntr:(d + 1) * ntr], want[d][1])


def test_split_reassociation_with_tie_terms(codes_spec, dets):
    """The split path re-associates the float64 additions of log_prob_sequence (Pd_plotter.py:106-116) inside a binade; a
    term v with v / ulp ending in exactly one half is the one case in which the step-by-step rounding depends on the running
    sum (round-half-even).  Log tables with such terms planted in the binades the sums pass through (2^8 .. 2^15), positive
    terms (no monotone sum: every term is added in order) and an all-zero table: per-trial sums bit-identical to Python's own
    float additions over the oracle's trajectory."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    spec = codes_spec["c75"]
    det = dets("c75")
    tab, P1, Tref = _oracle_models(det, spec, 0.1, 8000, 5)
    T = bitsource.bsc_threshold(0.1)
    ntr, N, seed = 5, 30011, 99
    taps = _taps(spec)
    trajs = []
    for t in range(ntr):
        U, E = co.trial_words(seed, 7, t, N, 2, T)
        idx, rseq, _ = co.simulate(taps, taps, 2, spec["m"], N, U, E, tab)
        trajs.append((np.asarray(idx[:N], dtype=np.int64), np.asarray(rseq[:N], dtype=np.int64)))
    logP1 = np.log(np.maximum(np.asarray(P1, dtype=np.float64), 1e-300)).reshape(det.S, det.R)
    logT = np.log(np.maximum(np.asarray(Tref, dtype=np.float64), 1e-300)).reshape(det.S, det.R)
    visits = np.zeros((det.S, det.R), dtype=np.int64)
    for idx, rseq in trajs:
        np.add.at(visits, (idx, rseq), 1)
    busy = np.dstack(np.unravel_index(np.argsort(-visits, axis=None)[:12], visits.shape))[0]
    tie1, tie0 = logP1.copy(), logT.copy()
    for j, (i, r) in enumerate(busy):                              # the most visited edges become tie terms of binade 2^k
        k = 8 + j % 8
        tie1[i, r] = -(0.5 + 2.0 ** (k - 53))
        if j < 3:
            tie0[i, r] = -(0.75 + 2.0 ** (k + 3 - 53))
    vals, cnts = np.unique(logT[logT != 0.0], return_counts=True)
    cls0 = logT.copy()                                             # still three distinct values (class mode), the commonest one a tie
    cls0[logT == vals[np.argmax(cnts)]] = -(0.5 + 2.0 ** (12 - 53))    # term of binade 2^12
    pos1 = logP1.copy()
    pos1[busy[0][0], busy[0][1]] = 0.25                            # a positive term: the sums are not monotone
    seg = [Seg(N=N, threshold=T, stream=7, enc_taps=taps, decide=0, trial_begin=0, trial_end=ntr)]
    det.split_trials(1)
    try:
        for name, l1, l0 in (("plain", logP1, logT), ("ties", tie1, tie0), ("class_tie", logP1, cls0), ("positive", pos1, logT),
                             ("zero", np.zeros_like(logP1), logT)):
            det.set_loglik(l1[None], l0)
            _, lp = det.detect(seg, seed=seed, engine="fsm", want_logp=True)
            assert (det.last_kernel_kind() & 16384) != 0
            subs, seq = det.split_stats()
            for t, (idx, rseq) in enumerate(trajs):
                a1 = a0 = 0.0
                for v1, v0 in zip(l1[idx, rseq].tolist(), l0[idx, rseq].tolist()):
                    a1 += v1
                    a0 += v0
                assert (a1, a0) == (lp[t, 0], lp[t, 1]), (name, t)
            if name == "plain":
                assert seq < 0.3 * subs
                plain_seq = seq
            elif name in ("ties", "class_tie"):
                assert plain_seq < seq < subs, "the planted tie terms did not force any sub-chunk back to term-by-term additions"
            elif name == "positive":
                assert seq == subs
    finally:
        det.split_trials(0)
        det.set_models([P1])


@pytest.mark.parametrize("dec,enc,p", [("c75", "c65", 0.1), ("m3a", "m3b", 0.05)])
def test_config3_blocklengths_vs_oracle(codes_spec, dets, dec, enc, p):
    """BASELINE config 3 (Pd vs blocklength, N = 10^2 .. 10^5): N = 10^4, 10^5 and the ragged 100 003 through the
    time-split path, the one-thread-per-trial NEXT walk and the ACS kernels -- tallies and per-trial float64 sums
    bit-identical to the C oracle over the whole length."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    spec = codes_spec[dec]
    det = dets(dec)
    tab, P1, Tref = _oracle_models(det, spec, p, 8000, 5)
    det.set_models([P1])
    T = bitsource.bsc_threshold(p)
    ntr = 9
    Ns = (10000, 100000, 100003)
    segs = [Seg(N=N, threshold=T, stream=40 + 2 * i + d, enc_taps=_taps(codes_spec[enc if d else dec]), decide=d, trial_begin=3,
                trial_end=3 + ntr) for i, N in enumerate(Ns) for d in (0, 1)]
    want = [co.run_trials(_taps(spec), _taps(codes_spec[enc if d else dec]), spec["n"], spec["m"], N, T, 77, 40 + 2 * i + d, 3,
                          3 + ntr, tab, P1, Tref, d, want_logp=True) for i, N in enumerate(Ns) for d in (0, 1)]
    runs = []
    det.split_trials(1)
    try:
        runs.append(det.detect(segs, seed=77, engine="fsm", want_logp=True))
        assert (det.last_kernel_kind() & 16384) != 0
        subs, seq = det.split_stats()
        assert subs == ntr * 2 * sum(-(-N // 128) for N in Ns) and seq < 0.15 * subs    # most additions re-associated
        det.split_trials(0)
        runs.append(det.detect(segs, seed=77, engine="auto", want_logp=True))       # automatic choice: few long trials split
        assert (det.last_kernel_kind() & 16384) != 0
        det.split_trials(2)
        runs.append(det.detect(segs, seed=77, engine="fsm", want_logp=True))
        assert (det.last_kernel_kind() & 16384) == 0
        runs.append(det.detect(segs, seed=77, engine="acs", want_logp=True))
    finally:
        det.split_trials(0)
    for tallies, lp in runs:
        for j, (wt, wlp) in enumerate(want):
            assert int(tallies[j]) == wt
            assert np.array_equal(lp[j * ntr:(j + 1) * ntr], wlp)


def test_config3_full_size_paths_agree(codes_spec, dets):
    """BASELINE config 3 at full size (N = 10^5, 2 000 trials per hypothesis, 4e8 trellis steps): the split path with
    re-associated sums == the same path adding every term in order == one thread per trial -- tallies and all 4 000 pairs of
    float64 sums; 8 of the trials against the C oracle over their whole length."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import Seg
    spec = codes_spec["c75"]
    det = dets("c75")
    tab, P1, Tref = _oracle_models(det, spec, 0.1, 8000, 5)
    det.set_models([P1])
    T = bitsource.bsc_threshold(0.1)
    N, ntr = 100000, 2000
    segs = [Seg(N=N, threshold=T, stream=60 + d, enc_taps=_taps(codes_spec["c65" if d else "c75"]), decide=d, trial_begin=0,
                trial_end=ntr) for d in (0, 1)]
    runs = []
    try:
        for split, seq in ((1, False), (1, True), (2, False)):
            det.split_trials(split)
            det.split_sequential(seq)
            runs.append(det.detect(segs, seed=5, engine="fsm", want_logp=True))
            assert bool(det.last_kernel_kind() & 16384) == (split == 1)
            if split == 1 and not seq:
                subs, nseq = det.split_stats()
                assert subs == 2 * ntr * 782 and nseq < 0.05 * subs
    finally:
        det.split_trials(0)
        det.split_sequential(False)
    for t, lp in runs[1:]:
        assert np.array_equal(t, runs[0][0]) and np.array_equal(lp, runs[0][1])
    for d in (0, 1):
        want, wlp = co.run_trials(_taps(spec), _taps(codes_spec["c65" if d else "c75"]), 2, 2, N, T, 5, 60 + d, 0, 4, tab, P1, Tref, d,
                                  want_logp=True)
        assert np.array_equal(runs[0][1][d * ntr:d * ntr + 4], wlp)


_BENCH_GEOMETRY_WANT = {}


@pytest.mark.parametrize("engine", ENGINES)
def test_bench_geometry_one_segment_vs_oracle(codes_spec, dets, engine):
    """The launch geometry bench.py times (BASELINE configs[1]: 7 p x 2 hypotheses = 14 segments in one launch,
    grid.y = 14, N = 500, P1 learned per p as run_experiment does) at 2 x 10^5 trials per segment: the tallies of the
    p = 0.1 point -- both hypotheses, every one of the 2 x 10^5 trials -- equal the C oracle's, and every other
    segment equals its own single-segment launch."""
    import c_oracle as co
    import Pd_plotter as pdp
    from mvd import bitsource
    from mvd.engine import Seg
    det = dets("c75")
    p_vec = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
    _, tables = pdp._learn_edge_tables(det, p_vec, None, 200, 1.0, 12345)
    det.set_models(tables)
    t1, t2 = _taps(codes_spec["c75"]), _taps(codes_spec["c65"])
    ntr = 200_000
    segs = []
    for q, p in enumerate(p_vec):
        T = bitsource.bsc_threshold(p)
        segs.append(Seg(N=500, threshold=T, stream=2 * q, table=q, enc_taps=t1, decide=0, trial_begin=0, trial_end=ntr))
        segs.append(Seg(N=500, threshold=T, stream=2 * q + 1, table=q, enc_taps=t2, decide=1, trial_begin=0, trial_end=ntr))
    tallies = det.detect(segs, seed=12345, engine=engine)
    kind = det.last_kernel_kind()
    assert kind != 0 and ((kind & 256) != 0) == (engine == "acs")           # the headline kernels, not the generic ones
    tab = co.Table(det.table.metrics, 2)
    Tref = __import__("mvd.codes", fromlist=["codes"]).tref_half_table(det.table)
    q = 2
    for h, enc in enumerate((t1, t2)):
        if h not in _BENCH_GEOMETRY_WANT:                                    # 10^8 oracle steps: once for both engines
            _BENCH_GEOMETRY_WANT[h] = co.run_trials(t1, enc, 2, 2, 500, bitsource.bsc_threshold(0.1), 12345, 2 * q + h, 0, ntr,
                                                    tab, tables[q], Tref, h)
        assert int(tallies[2 * q + h]) == _BENCH_GEOMETRY_WANT[h]
    assert 0 < int(tallies[2 * q]) < ntr                                     # p = 0.1 is the informative point (Pd ~ 0.26)
    for j in (0, 7, 13):
        assert int(det.detect([segs[j]], seed=12345, engine=engine)[0]) == int(tallies[j])


def test_gpu_bfs_limits_and_count_only(codes_spec):
    """max_states exhausted -> MVD_E_NOMEM with a usable lower bound; count-only mode keeps no NEXT table;
    the recursion of m = 5 is enumerated up to a budget (its closure is far larger, SURVEY 8 a5 note)."""
    from mvd import _capi
    from mvd.engine import Detector, HashOnlyDetector
    s = codes_spec["m3a"]
    with Detector(s["gen"], 1, 2, 3, enumerate_with="lib") as det:
        with pytest.raises(_capi.MvdError) as ei:
            det._enumerate_gpu(100, install=False)
        assert ei.value.code == -5
        st = det._enumerate_gpu(100, install=False, allow_partial=True)
        assert not st["closed"] and 0 < st["S"] <= 100 and st["frontier"] < st["S"]
        st = det._enumerate_gpu(1 << 10, install=False, count_only=True)
        assert st["closed"] and st["S"] == 435
    s = codes_spec["m5"]
    with HashOnlyDetector(s["gen"], 1, 2, 5) as det:
        st = det._enumerate_gpu(1 << 21, install=False, count_only=True, allow_partial=True)
        assert not st["closed"] and st["S"] > 1 << 20 and st["max_metric"] <= 15


@pytest.mark.parametrize("name", ["m5", "m6", "m4a", "c75"])
def test_acs_hash_large_memory(codes_spec, name):
    """Eq. 4-5 recursion for memories without an enumerable state set: trajectory hash and
    final metric vector vs the oracle."""
    import c_oracle as co
    from mvd import bitsource
    from mvd.engine import HashOnlyDetector, Seg
    spec = codes_spec[name]
    taps = _taps(spec)
    T = bitsource.bsc_threshold(0.08)
    with HashOnlyDetector(spec["gen"], 1, spec["n"], spec["m"]) as det:
        seg = Seg(N=257, threshold=T, stream=2, enc_taps=taps, trial_begin=40, trial_end=40 + 300)
        h, fin = det.acs_hash(seg, seed=31)
        for t in (0, 1, 150, 299):
            U, E = co.trial_words(31, 2, 40 + t, 257, spec["n"], T)
            wh, wf = co.acs_hash(taps, taps, spec["n"], spec["m"], 257, U, E)
            assert int(h[t]) == wh
            assert np.array_equal(fin[t], wf)
        # throughput form (two trials per thread, every metric pair in registers): the same final vectors
        fin2 = det.acs_final(seg, seed=31)
        assert det.last_kernel_kind() == 32768
        assert np.array_equal(fin2, fin)
        odd = Seg(N=33, threshold=T, stream=5, enc_taps=taps, trial_begin=7, trial_end=7 + 513)     # ragged block, odd trial count
        assert np.array_equal(det.acs_final(odd, seed=9), det.acs_hash(odd, seed=9)[1])


def test_engines_and_sources_agree_at_scale(codes_spec, dets):
    """Size-independent properties at 2 x 10^5 trials x N = 500: ACS == FSM tallies, Philox ==
    bitstream of the same bits (subset), and shard invariance (split ranges sum to the whole)."""
    from mvd import bitsource
    from mvd.engine import Seg
    det = dets("c75")
    _, P1, _ = _oracle_models(det, codes_spec["c75"], 0.1, 6200, 123)
    det.set_models([P1])
    T = bitsource.bsc_threshold(0.1)
    taps2 = _taps(codes_spec["c65"])
    ntr = 200_000
    whole = [Seg(N=500, threshold=T, stream=0, decide=0, trial_begin=0, trial_end=ntr),
             Seg(N=500, threshold=T, stream=1, enc_taps=taps2, decide=1, trial_begin=0, trial_end=ntr)]
    ta = det.detect(whole, seed=12345, engine="acs")
    tf = det.detect(whole, seed=12345, engine="fsm")
    assert np.array_equal(ta, tf)
    cuts = [0, 1, 70_001, 133_337, ntr]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        parts += [Seg(N=500, threshold=T, stream=0, decide=0, trial_begin=a, trial_end=b),
                  Seg(N=500, threshold=T, stream=1, enc_taps=taps2, decide=1, trial_begin=a, trial_end=b)]
    tp = det.detect(parts, seed=12345, engine="fsm")
    assert int(tp[0::2].sum()) == int(tf[0]) and int(tp[1::2].sum()) == int(tf[1])
    # bitstream of the same bits for the first 64 trials
    u, e = bitsource.philox_bitstreams(12345, 1, 0, 64, 500, 2, T)
    seg = Seg(N=500, enc_taps=taps2, decide=1, trial_begin=0, trial_end=64)
    tb, lpb = det.detect([seg], bits=bitsource.pack_bitstreams(u, e), engine="fsm", want_logp=True)
    seg_p = Seg(N=500, threshold=T, stream=1, enc_taps=taps2, decide=1, trial_begin=0, trial_end=64)
    tq, lpq = det.detect([seg_p], seed=12345, engine="acs", want_logp=True)
    assert int(tb[0]) == int(tq[0]) and np.array_equal(lpb, lpq)


def test_async_detect_queue_matches_synchronous_calls(codes_spec, dets):
    """MVD_OPT_ASYNC_DETECT: five sweeps queued without a host round trip (two alternating device tally buffers, more
    launches than the in-flight ring holds on a second run) give the tallies of five synchronous calls; the kernel
    times of all of them are accounted; a call that reads results drains the queue."""
    import torch
    from mvd import bitsource
    from mvd.engine import Seg
    spec = codes_spec["c75"]
    det = dets("c75")
    _, P1, _ = _oracle_models(det, spec, 0.1, 8000, 5)
    det.set_models([P1])
    T = bitsource.bsc_threshold(0.1)
    mk = lambda i: [Seg(N=200, threshold=T, stream=7 * i + d, enc_taps=_taps(codes_spec[("c75", "c65")[d]]), decide=d,
                        trial_begin=0, trial_end=5000) for d in (0, 1)]
    want = [det.detect(mk(i), seed=77, engine="acs") for i in range(5)]
    bufs = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(5)]
    torch.cuda.synchronize()
    det.async_detect(True)
    try:
        det.async_stats()
        for i in range(5):
            assert det.detect(mk(i), seed=77, engine="acs", d_tallies_ptr=bufs[i].data_ptr(), host_tallies=False) is None
        det.synchronize()
        ms, n = det.async_stats()
        assert n == 5 and ms > 0
        for i in range(5):
            assert bufs[i].cpu().numpy().astype(np.uint64).tolist() == want[i].tolist()
        for i in range(70):                                   # more than the 64 launches kept in flight
            det.detect(mk(i % 5), seed=77, engine="acs", d_tallies_ptr=bufs[i % 5].data_ptr(), host_tallies=False)
        got = det.detect(mk(0), seed=77, engine="acs")        # a synchronous call drains the queue first
        assert got.tolist() == want[0].tolist()
        assert det.async_stats()[1] == 70
        assert bufs[4].cpu().numpy().astype(np.uint64).tolist() == want[4].tolist()
    finally:
        det.async_detect(False)


def test_edge_cases(codes_spec, dets):
    from mvd import bitsource
    from mvd.engine import Seg
    det = dets("c75")
    _, P1, _ = _oracle_models(det, codes_spec["c75"], 0.1, 6200, 123)
    det.set_models([P1])
    # empty trial range, N = 0 (logp = 0: H1 rule fails, H2 rule succeeds), p = 0, p = 1
    segs = [Seg(N=100, threshold=5, stream=0, decide=0, trial_begin=7, trial_end=7),
            Seg(N=0, threshold=5, stream=0, decide=0, trial_begin=0, trial_end=10),
            Seg(N=0, threshold=5, stream=0, decide=1, trial_begin=0, trial_end=10)]
    for engine in ENGINES:
        t = det.detect(segs, seed=1, engine=engine)
        assert t.tolist() == [0, 0, 10]
    # p = 0: noiseless all-zero-input chain stays in state 0 under its own decoder
    seg = Seg(N=200, threshold=0, stream=0, random_input=False, trial_begin=0, trial_end=3)
    for engine in ENGINES:
        idx, met = det.trace(seg, seed=1, engine=engine)
        # the all-zero codeword through a noiseless channel: metric of state 0 stays 0
        assert (met[:, :, 0] == 0).all()
    # p = 1 (threshold clipped to 2^32 - 1): every bit flips (up to the 2^-32 clip)
    T1 = bitsource.bsc_threshold(1.0)
    assert T1 == 0xFFFFFFFF
    idx1, _ = det.trace(Seg(N=64, threshold=T1, stream=0, random_input=False, trial_begin=0, trial_end=2), seed=1, engine="fsm")
    ones = np.ones((2, 2, 64), dtype=np.uint8)
    idx2, _ = det.trace(Seg(N=64, random_input=False, trial_begin=0, trial_end=2),
                        bits=bitsource.pack_bitstreams(np.zeros((2, 64), dtype=np.uint8), ones), engine="fsm")
    assert np.array_equal(idx1, idx2)


def test_unknown_state_is_an_error(codes_spec):
    """A metric vector outside the table = the reference's KeyError (Pd_plotter.py:112)."""
    from mvd import _capi, codes
    from mvd.engine import Detector, Seg
    s75, s65 = codes_spec["c75"], codes_spec["c65"]
    wrong = codes.enumerate_states(codes.freeze_generator(s65["gen"]), 2, 1, 2)      # (6,5)'s 5 states
    with Detector(s75["gen"], 1, 2, 2, table=wrong) as det:
        with pytest.raises(_capi.UnknownStateError):
            det.trace(Seg(N=200, threshold=1 << 30, trial_begin=0, trial_end=4), seed=3, engine="acs")
        assert isinstance(_capi.UnknownStateError(-6, "x"), KeyError)


def test_abi_argument_errors(codes_spec, dets):
    from mvd import _capi
    from mvd.engine import Detector, Seg
    det = dets("c75")
    with pytest.raises(_capi.MvdError):
        det.detect([Seg(N=10, table=99, trial_begin=0, trial_end=1)], seed=1)           # table out of range
    with pytest.raises(_capi.MvdError):
        det.detect([Seg(N=10, trial_begin=5, trial_end=1)], seed=1)                     # end < begin
    with pytest.raises(_capi.MvdError):
        det.detect([Seg(N=300, trial_begin=0, trial_end=4)], bits=np.zeros((1, 3, 4, 4), dtype=np.uint32))  # short
    with pytest.raises(_capi.MvdError) as exc:
        Detector([[[1, 1, 1]] * 4, [[1, 0, 1]] * 4], 4, 2, 2)                            # k = 4 > MVD_MAX_K inputs per step
    assert exc.value.code == -3
    with pytest.raises(_capi.MvdError):                                                 # tap-mask form is k = 1 (tables: test_k2_codes.py)
        det._ck(det.lib.mvd_set_code(det.ctx, 2, 2, 2, (C.c_uint32 * 2)(7, 5)))


def test_int_peak_and_info(dets):
    det = dets("c75")
    info = det.device_info()
    assert info["sm_count"] > 0
    alu, mixed = det.int_peak()
    assert alu > 1000 and mixed > 1000          # Gop/s
    assert det.launch_count() > 0


# ------------------------------------------------------------------ entry points, as a user runs them
def _run_script(args, stdin_text, cwd):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "detecting-convolutional-codes-via-markovian-statistics_b200")
    env = dict(os.environ, PYTHONPATH=pkg)
    return subprocess.run([sys.executable] + [os.path.join(pkg, args[0])] + args[1:], input=stdin_text, text=True,
                          capture_output=True, cwd=cwd, env=env, timeout=600)


def test_pd_plotter_main_writes_the_reference_csv(tmp_path):
    """`python Pd_plotter.py` (reference Pd_plotter.py:242-264): results_experiments/Pd_hybrid_results.csv
    with header N,p,Pd,Pc, 7 rows (N = 500, the reference's p grid), 10^4 iterations per point."""
    import csv
    out = _run_script(["Pd_plotter.py"], "", str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    path = tmp_path / "results_experiments" / "Pd_hybrid_results.csv"
    rows = list(csv.reader(open(path)))
    assert rows[0] == ["N", "p", "Pd", "Pc"]
    assert [r[0] for r in rows[1:]] == ["500"] * 7
    assert [float(r[1]) for r in rows[1:]] == [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
    pd_, pc_ = [float(r[2]) for r in rows[1:]], [float(r[3]) for r in rows[1:]]
    assert pd_[0] == 1.0 and pc_[0] == 1.0                     # clean channel: perfect detection
    assert all(0.0 <= v <= 1.0 for v in pd_ + pc_)
    assert abs(pc_[-1] - 0.5) < 0.02                           # p = 1/2: the channel carries nothing
    assert all(round(v * 10000) == pytest.approx(v * 10000, abs=1e-6) for v in pd_)     # multiples of 1 / num_iter


def test_demo_script_predefined_pair(tmp_path):
    """`python demo_script.py`, option 1 / example 1 (reference demo_script.py:82-131): runs, prints a
    5-row table (matplotlib is absent here), same numbers as run_experiment with the demo's arguments."""
    import Pd_plotter as pdp
    out = _run_script(["demo_script.py"], "1\n1\n", str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    df = pdp.run_experiment(k=1, n=2, m=2, gen1=[[[1, 1, 1]], [[1, 0, 1]]], gen2=[[[1, 1, 0]], [[1, 0, 1]]], num_iter=2000,
                            p_vec=[0.01, 0.05, 0.1, 0.2, 0.3], learn_len=None, learn_burn=200, laplace=1.0, seed=123)
    assert len(df) == 5 and df["N"].tolist() == [500] * 5
    if "matplotlib is not installed" in out.stdout:
        table = [ln.split() for ln in out.stdout.splitlines() if ln.strip().startswith("500")]
        assert len(table) == 5
        assert [float(r[2]) for r in table] == pytest.approx(df["Pd"].tolist(), abs=1e-6)
        assert [float(r[3]) for r in table] == pytest.approx(df["Pc"].tolist(), abs=1e-6)


def test_simulate_markov_sequence_contract(golden):
    """The function the reference calls but does not ship: positional call as at Pd_plotter.py:212,
    keyword call as at :149-155; returns length+1 hashable metric tuples starting at all-zero."""
    import viterbi_markov as vm
    gen1 = [[[1, 1, 1]], [[1, 0, 1]]]
    a = vm.simulate_markov_sequence(gen1, 2, 1, 2, 300, 0.1, True)
    assert len(a["metrics"]) == 301 and a["metrics"][0] == (0, 0, 0, 0)
    assert all(isinstance(d, tuple) and min(d) == 0 for d in a["metrics"])
    b = vm.simulate_markov_sequence(gen1, 2, 1, 2, 300, p_val=0.1, random_input=True, seed=12345)
    c = vm.simulate_markov_sequence(gen1, 2, 1, 2, 300, p_val=0.1, random_input=True, seed=12345)
    assert b["metrics"] == c["metrics"]                        # seeded: reproducible
    g = golden["sim_kats"]["c75_self"]
    d = vm.simulate_markov_sequence(gen1, 2, 1, 2, g["N"], g["p"], True, g["seed"], stream=g["stream"], trial=g["trial"])
    assert [list(x) for x in d["metrics"]] == g["metrics"]
    e = vm.simulate_markov_sequence(gen1, 2, 1, 2, g["N"], g["p"], True, None, u_bits=g["u_bits"], e_bits=g["e_bits"])
    assert [list(x) for x in e["metrics"]] == g["metrics"]


def test_learn_P1_empirical_contract(golden):
    """learn_P1_empirical(gens_tuple, k, n, m, p, learn_len, learn_burn, laplace, seed) -> (states,
    state_index, P) like the reference (Pd_plotter.py:123-169): dense row-stochastic P, cached."""
    import Pd_plotter as pdp
    g = golden["experiments"]["c75_c65_small"]
    gens = tuple(tuple(tuple(t) for t in row) for row in g["gen1"])
    states, index, P = pdp.learn_P1_empirical(gens, 1, 2, 2, 0.1, None, 200, 1.0, 123)
    assert len(states) == 31 and index[(0, 0, 0, 0)] == 0 and P.shape == (31, 31)
    np.testing.assert_allclose(P.sum(axis=1), 1.0, rtol=1e-12)
    want = np.array(g["P1_edge"]["0.1"]["edge"])
    tab = __import__("viterbi_markov").state_table(g["gen1"], 2, 1, 2)
    assert np.array_equal(P[np.arange(31)[:, None], tab.nxt], want)
    assert pdp.learn_P1_empirical(gens, 1, 2, 2, 0.1, None, 200, 1.0, 123)[2] is P      # lru_cache, Pd_plotter.py:123


def test_integration_md_stub_runs_as_written():
    """The ctypes stub INTEGRATION.md shows a maintainer of the reference (section B) is executed verbatim against
    libmvd.so, with the reference-side objects it expects in scope, and must reproduce run_experiment's tallies."""
    import os
    import re
    import Pd_plotter as pdp
    import viterbi_markov as vm
    from mvd import _capi, codes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"## B\..*?```python\n(.*?)```", text, flags=re.S).group(1)
    block = block.replace('C.CDLL("libmvd.so")', f'C.CDLL("{_capi.LIB_PATH}")')
    k, n, m = 1, 2, 2
    gen1, gen2 = [[[1, 1, 1]], [[1, 0, 1]]], [[[1, 1, 0]], [[1, 0, 1]]]
    states, transitions, all_r = vm.enumerate_markov_states_allzero(gen1, m, k, n)
    env = dict(vm=vm, k=k, n=n, m=m, gen1=gen1, gen2=gen2, states=states, all_r=all_r, index={s: i for i, s in enumerate(states)},
               trellis=vm.build_trellis(gen1, m, k), S=len(states), seed=12345, p=0.1, N=300, num_iter=4000,
               learn_len_eff=max(5000, 200 * len(states)), learn_burn=200, laplace=1.0,
               T_ref=vm.numeric_T(states, transitions, all_r, 0.5))
    exec(compile(block, "INTEGRATION.md", "exec"), env)
    d = {}
    df = pdp.run_experiment(k, n, m, gen1, gen2, 4000, [0.1], None, 200, 1.0, 12345, N_spectrum=[300], details=d)
    assert env["tallies"].tolist() == [int(v) for v in d["tallies"]]
    assert env["Pd"] == df["Pd"][0] and env["Pc"] == df["Pc"][0]
    # the second snippet: comp_parity.py's Monte-Carlo loop through mvd_parity_detect
    import ctypes as C
    import comp_parity as cp
    import parity_eqn_check as pec
    snippet = [b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S) if "mvd_parity_detect" in b][0]
    generators = [[pec.parse_poly_token("7")], [pec.parse_poly_token("5")]]
    other = [[pec.parse_poly_token("6")], [pec.parse_poly_token("5")]]
    template, _ = cp.template_from_generators(generators, 2)
    ctx = C.c_void_p()
    assert env["lib"].mvd_create(C.byref(ctx), 0) == 0
    env.update(generators=generators, other=other, template=template, gamma=0.6, trials=5000, N=200, ctx=ctx,
               tallies=np.zeros(2, dtype=np.uint64))
    exec(compile(snippet, "INTEGRATION.md#parity", "exec"), env)
    env["lib"].mvd_destroy(ctx)
    tm = cp._template_masks(template, 2)
    want = cp.ParityContext().run([dict(N=200, m=2, taps=cp._tap_masks(g), tmpl=tm, gamma=0.6, decide=h, threshold=env["T"], stream=h,
                                        trial_begin=0, trial_end=5000) for h, g in enumerate((generators, other))], seed=12345)
    assert env["tallies"].tolist() == want.tolist() and env["accuracy"] == want[0] / 5000
