"""Error-exponent path (reference alpha_exponent.py): oracle restatement and host logic on the
CPU against golden vectors produced by the reference's own functions; the product (GPU chain + GPU spectral
radius) against the same goldens and the oracle under ``-m gpu``."""
import hashlib

import numpy as np
import pytest

CASES = ["c75_vs_c65", "c75_lap", "m3"]


def _edge_and_bg(C, nxt):
    K, _, R = C.shape
    edge = C[np.arange(K)[:, None], nxt, np.arange(R)[None, :]]
    mask = np.ones_like(C, dtype=bool)
    mask[np.arange(K)[:, None], nxt, np.arange(R)[None, :]] = False
    bg = np.array([C[i][mask[i]].max() for i in range(K)])
    return edge, bg


# ------------------------------------------------------------------------------------------ CPU: oracle + host logic
@pytest.mark.parametrize("case", CASES)
def test_oracle_tensor_matches_reference(golden, case):
    """ref_port.learn_transition_tensor == the reference's learn_transition_tensor on the same bits: the whole
    K x K x R float64 tensor (sha256 of its bytes), for both hypotheses."""
    import ref_port
    g = golden["alpha_kats"][case]
    for hyp, enc in enumerate((g["enc1"], g["enc2"])):
        C, states, sidx, all_r, counts = ref_port.learn_transition_tensor(
            enc, g["dec"], g["m"], g["p"], g["length"], g["burn_in"], g["laplace"], g["seed"], trial=hyp)
        t = g["tensors"][hyp]
        assert C.shape == (t["K"], t["K"], t["R"])
        assert hashlib.sha256(np.ascontiguousarray(C).tobytes()).hexdigest()[:16] == t["sha"]
        assert int(counts.sum()) == g["length"]


@pytest.mark.parametrize("case", ["c75_vs_c65", "c75_lap"])
def test_oracle_exponent_matches_reference(golden, case):
    import ref_port
    g = golden["alpha_kats"][case]
    Cs = [ref_port.learn_transition_tensor(enc, g["dec"], g["m"], g["p"], g["length"], g["burn_in"], g["laplace"],
                                           g["seed"], trial=h)[0] for h, enc in enumerate((g["enc1"], g["enc2"]))]
    I, u, rhos = ref_port.compute_error_exponent(Cs[0], Cs[1], g["u_grid"])
    assert I == g["I_err"] and u == g["best_u"]
    for key, want in g["rho"].items():
        got = float(np.max(np.abs(np.linalg.eigvals(ref_port.chernoff_matrix(Cs[0], Cs[1], float(key))))))
        assert got == pytest.approx(want, rel=1e-13)


def test_c_oracle_chain_with_reference_bit_order(golden):
    """The C oracle's chain fed with the effective tap masks of the reference's _encoder_step == the Python
    port's counts: ties the fast checker used on the GPU box to the pinned one."""
    import c_oracle as co
    import ref_port
    import alpha_exponent as ae
    from mvd import bitsource, codes
    g = golden["alpha_kats"]["c75_vs_c65"]
    dec_masks = codes.tap_masks(codes.freeze_generator([[t] for t in g["dec"]]), g["m"], 1)
    met, nxt = co.enumerate_states(dec_masks, 2, g["m"], max_states=4096)
    tab = co.Table(met, g["m"])
    for hyp, enc in enumerate((g["enc1"], g["enc2"])):
        _, _, _, _, counts = ref_port.learn_transition_tensor(enc, g["dec"], g["m"], g["p"], g["length"], g["burn_in"],
                                                              g["laplace"], g["seed"], trial=hyp)
        edge, _ = co.learn_chain(dec_masks, ae.effective_encoder_taps(enc, g["m"]), 2, g["m"], g["burn_in"] + g["length"],
                                 g["burn_in"], bitsource.bsc_threshold(g["p"]), g["seed"], bitsource.ALPHA_STREAM, hyp, tab)
        K, R = nxt.shape
        want = counts[np.arange(K)[:, None], nxt, np.arange(R)[None, :]]
        assert np.array_equal(edge.reshape(K, R).astype(float), want)


def test_host_helpers_match_reference(golden):
    import alpha_exponent as ae
    import viterbi_markov as vm
    g = golden["alpha_kats"]
    taps = [[1, 1, 0, 1], [1, 0, 1, 1]]
    for s, u, y, ns in g["encoder_steps"]:
        assert ae._encoder_step(s, u, taps, 3) == (tuple(y), ns)
    I, A = ae.fit_error_exponent(g["fit"]["N"], g["fit"]["Pe"])
    assert I == pytest.approx(g["fit"]["result"][0], rel=1e-12) and A == pytest.approx(g["fit"]["result"][1], rel=1e-12)
    few = ae.fit_error_exponent([10, 20], [0.1, 0.05])
    assert few[0] == g["fit"]["few"][0] == 0.0 and np.isnan(few[1])
    assert ae.spectral_radius([[0.0, 2.0], [0.5, 0.0]]) == pytest.approx(1.0)
    # the chain of _encoder_step == convolution with the effective tap masks
    rng = np.random.default_rng(1)
    for tp, m in ((taps, 3), ([[1, 1, 0], [1, 0, 1]], 2), ([[1, 0, 0, 1, 1], [1, 1, 1, 0, 1]], 4)):
        masks = ae.effective_encoder_taps(tp, m)
        plain = ae.effective_encoder_taps(tp, m, reference_bit_order=False)
        assert plain == [sum(b << i for i, b in enumerate(gj)) for gj in tp]
        st, hist = 0, []
        for t in range(100):
            u = int(rng.integers(0, 2))
            hist.append(u)
            y, st = ae._encoder_step(st, u, tp, m)
            for j, mk in enumerate(masks):
                assert y[j] == sum(hist[t - d] for d in range(m + 1) if (mk >> d) & 1 and t - d >= 0) % 2
    assert vm.octal_to_taps("7") == [1, 1, 1] and vm.octal_to_taps(6, 2) == [0, 1, 1] and vm.octal_to_taps("5", 3) == [1, 0, 1, 0]
    with pytest.raises(ValueError):
        vm.octal_to_taps("17", 2)


# ------------------------------------------------------------------------------------------ GPU: the product
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_tensor_and_exponent_match_reference(golden, case):
    """learn_transition_tensor (GPU chain) reproduces the reference's tensor bit for bit; compute_error_exponent
    and error_exponent_from_edges (GPU power iteration) reproduce its I_err / best_u (LAPACK eigenvalues) to 1e-9."""
    import alpha_exponent as ae
    g = golden["alpha_kats"][case]
    Cs, edges = [], []
    for hyp, enc in enumerate((g["enc1"], g["enc2"])):
        counts, table = ae.learn_transition_edges(enc, g["dec"], g["m"], g["p"], g["length"], g["burn_in"], g["seed"], trial=hyp)
        assert int(counts.sum()) == g["length"]
        C = ae.edges_to_tensor(table, counts, g["laplace"])
        t = g["tensors"][hyp]
        assert hashlib.sha256(np.ascontiguousarray(C).tobytes()).hexdigest()[:16] == t["sha"]
        edge, bg = _edge_and_bg(C, table.nxt)
        assert np.array_equal(edge, np.array(t["edge"])) and np.array_equal(bg, np.array(t["background"]))
        Cs.append(C)
        edges.append(counts)
    d1, d2 = {}, {}
    I, u = ae.compute_error_exponent(Cs[0], Cs[1], g["u_grid"], details=d1)
    assert I == pytest.approx(g["I_err"], rel=1e-9) and u == g["best_u"]
    I2, u2 = ae.error_exponent_from_edges(table, edges[0], edges[1], g["laplace"], g["u_grid"], details=d2)
    assert I2 == pytest.approx(g["I_err"], rel=1e-9) and u2 == g["best_u"]
    np.testing.assert_allclose(d1["rho"], d2["rho"], rtol=1e-11)
    for key, want in g["rho"].items():
        q = int(np.argmin(np.abs(d1["u"] - float(key))))
        if abs(d1["u"][q] - float(key)) < 1e-12:
            assert d1["rho"][q] == pytest.approx(want, rel=1e-10) and d2["rho"][q] == pytest.approx(want, rel=1e-10)
    assert d1["iters"].max() < 100000 and d2["iters"].max() < 100000


@pytest.mark.gpu
def test_gpu_learn_transition_tensor_contract(golden):
    """Signature and return contract of alpha_exponent.learn_transition_tensor (alpha_exponent.py:83-149)."""
    import alpha_exponent as ae
    C, states, sidx, all_r = ae.learn_transition_tensor([[1, 1, 1], [1, 0, 1]], [[1, 1, 1], [1, 0, 1]], 2, 0.1,
                                                        length=5000, burn_in=100, laplace=1.0, seed=3)
    assert C.shape == (31, 31, 4) and states[0] == (0, 0, 0, 0) and sidx[states[7]] == 7
    assert all_r == [(0, 0), (0, 1), (1, 0), (1, 1)]
    np.testing.assert_allclose(C.sum(axis=(1, 2)), 1.0, rtol=1e-12)
    C2 = ae.learn_transition_tensor([[1, 1, 1], [1, 0, 1]], [[1, 1, 1], [1, 0, 1]], 2, 0.1, 5000, 100, 1.0, 3)[0]
    assert np.array_equal(C, C2)                                 # seeded: reproducible
    np.random.seed(11)
    a = ae.learn_transition_tensor([[1, 1, 1], [1, 0, 1]], [[1, 1, 1], [1, 0, 1]], 2, 0.1, 2000, 0)[0]
    b = ae.learn_transition_tensor([[1, 1, 1], [1, 0, 1]], [[1, 1, 1], [1, 0, 1]], 2, 0.1, 2000, 0)[0]
    assert not np.array_equal(a, b)                              # seed=None: no reseed (alpha_exponent.py:105-106)


@pytest.mark.gpu
def test_gpu_exponent_m4_edges_vs_dense_eigenvalues(codes_spec):
    """m = 4 (K = 25 751): the edge-form kernel runs where no dense tensor can exist; checked against scipy's
    sparse + rank-one eigenvalue of the same operator."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    import alpha_exponent as ae
    dec = [g[0] for g in codes_spec["m4a"]["gen"]]
    enc2 = [g[0] for g in codes_spec["m4b"]["gen"]]
    c1, table = ae.learn_transition_edges(dec, dec, 4, 0.05, 400000, 1000, 21, reference_bit_order=False, trial=0)
    c2, _ = ae.learn_transition_edges(enc2, dec, 4, 0.05, 400000, 1000, 21, reference_bit_order=False, trial=1)
    det = {}
    I, u = ae.error_exponent_from_edges(table, c1, c2, 1.0, 11, details=det)
    K, R = table.S, table.R
    lp1, lb1 = ae.edge_log_tensors(table, c1, 1.0)
    lp2, lb2 = ae.edge_log_tensors(table, c2, 1.0)
    for q in (0, 5, 10):
        uu = det["u"][q]
        bg = np.exp(uu * lb1 + (1 - uu) * lb2)
        w = np.exp(uu * lp1 + (1 - uu) * lp2) - bg[:, None]
        Sp = sp.csr_matrix((w.reshape(-1), (np.repeat(np.arange(K), R), table.nxt.reshape(-1))), shape=(K, K))
        op = spl.LinearOperator((K, K), matvec=lambda x: Sp @ x + bg * (R * x.sum()), dtype=float)
        val = spl.eigs(op, k=1, which="LM", tol=1e-12, return_eigenvectors=False)
        assert det["rho"][q] == pytest.approx(float(np.abs(val[0])), rel=1e-9)
    assert I == pytest.approx(float(-np.log(det["rho"].min())), rel=1e-15)
