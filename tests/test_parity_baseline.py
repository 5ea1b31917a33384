"""Parity-template baseline (reference parity_eqn_check.py + comp_parity.py): host algebra and the
oracle on the CPU against goldens produced by the reference's own functions; the GPU Monte-Carlo
(mvd_parity_detect) against the same goldens, bit-exact, under ``-m gpu``."""
import numpy as np
import pytest

TRIAL_CASES = ["main", "short", "m3", "r13"]


def _total(g):
    reach = max(s for _, s in g["template"])
    return g["N"] + g["m"] - reach


# ------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("name", ["c75", "m3", "m4", "r13"])
def test_parity_algebra_matches_reference(golden, name):
    """parse_poly_token / build_parity_system / nullspace_mod2 / parity_vector_to_equation == the reference's."""
    import parity_eqn_check as pec
    g = golden["parity_kats"]["bases"][name]
    gens = [[pec.parse_poly_token(t)] for t in g["tokens"]]
    assert gens == g["gens"]
    A = pec.build_parity_system(gens, g["deg_h"])
    assert A.tolist() == g["A"]
    basis = pec.nullspace_mod2(A)
    assert basis.tolist() == g["basis"] and basis.dtype == np.uint8
    assert not (A.astype(int) @ basis.T.astype(int) % 2).any()
    eqs = [pec.parity_vector_to_equation(pec.split_parity_vector(row, len(gens), g["deg_h"])) for row in basis]
    assert eqs == g["equations"]


def test_parse_poly_token_forms():
    import parity_eqn_check as pec
    assert pec.parse_poly_token("1,0,1") == [1, 0, 1] and pec.parse_poly_token(" 110 ") == [0, 1, 1]
    assert pec.parse_poly_token("7") == [1, 1, 1] and pec.parse_poly_token("13") == [1, 1, 0, 1]
    with pytest.raises(ValueError):
        pec.parse_poly_token("9x")
    assert pec.nullspace_mod2(np.eye(3, dtype=np.uint8)).shape == (0, 3)


@pytest.mark.parametrize("case", TRIAL_CASES)
def test_parity_oracle_and_host_scalars_match_reference(golden, case):
    """ref_port.parity_trial (the oracle) and the product's host scalars (encode_convolutional,
    parity_detector) reproduce the reference's per-trial (decision, P_hat) on the same bits."""
    import comp_parity as cp
    import ref_port
    g = golden["parity_kats"]["trials"][case]
    template = [tuple(t) for t in g["template"]]
    tmpl, _ = cp.template_from_generators(g["gens1"], g["m"], g["deg"])
    assert tmpl == template
    total = _total(g)
    for h, gens in enumerate((g["gens1"], g["gens2"])):
        hyp = g["hyp"][h]
        for trial, (want_dec, want_phat) in enumerate(hyp["rows"]):
            dec, sat, tot = ref_port.parity_trial(gens, g["m"], template, g["gamma"], g["N"], g["p"], g["seed"], hyp["stream"], trial)
            assert tot == total and dec == want_dec and sat / tot == want_phat
            if trial < 5:
                u, e = ref_port.philox_bits(g["seed"], hyp["stream"], trial, g["N"] + g["m"], len(gens), ref_port.threshold_of(g["p"]))
                v = cp.encode_convolutional(u[:g["N"]], gens, g["m"])
                assert v == ref_port.encode_convolutional(u[:g["N"]], gens, g["m"])
                y = [[v[j][t] ^ e[t][j] for t in range(g["N"] + g["m"])] for j in range(len(gens))]
                assert cp.parity_detector(y, template, g["gamma"]) == (want_dec, want_phat)


# ------------------------------------------------------------------------------------------ GPU
def _segments(cp, g, ntr, h_list=(0, 1)):
    from mvd import bitsource
    tmpl = cp._template_masks([tuple(t) for t in g["template"]], len(g["gens1"]))
    segs = []
    for h in h_list:
        gens = g["gens1"] if h == 0 else g["gens2"]
        segs.append(dict(N=g["N"], m=g["m"], taps=cp._tap_masks(gens), tmpl=tmpl, gamma=g["gamma"], decide=h,
                         threshold=bitsource.bsc_threshold(g["p"]), stream=g["hyp"][h]["stream"], trial_begin=0, trial_end=ntr))
    return segs


@pytest.mark.gpu
@pytest.mark.parametrize("case", TRIAL_CASES)
def test_gpu_parity_trials_match_reference(golden, case):
    """mvd_parity_detect on the on-device bit source: satisfied counts per trial and tallies == the
    reference's parity_detector on the same bits (golden), both hypotheses in one launch."""
    import comp_parity as cp
    g = golden["parity_kats"]["trials"][case]
    ctx = cp.ParityContext()
    total = _total(g)
    tallies, sat = ctx.run(_segments(cp, g, g["ntr"]), seed=g["seed"], want_satisfied=True)
    for h in (0, 1):
        rows = g["hyp"][h]["rows"]
        got = sat[h * g["ntr"]:(h + 1) * g["ntr"]]
        assert [int(s) / total for s in got] == [r[1] for r in rows]
        want_h1 = sum(1 for r in rows if r[0])
        assert int(tallies[h]) == (want_h1 if h == 0 else g["ntr"] - want_h1)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["main", "r13"])
def test_gpu_parity_bitstream_matches_philox_and_oracle(golden, case):
    """Host-supplied bit streams (verification mode) give the same counts as the on-device source, and a
    larger batch agrees with the oracle trial by trial."""
    import comp_parity as cp
    import ref_port
    from mvd import bitsource
    g = golden["parity_kats"]["trials"][case]
    ctx = cp.ParityContext()
    ntr, n, T = 300, len(g["gens1"]), g["N"] + g["m"]
    thr = bitsource.bsc_threshold(g["p"])
    seg = _segments(cp, g, ntr, h_list=(1,))[0]
    t_ph, s_ph = ctx.run([seg], seed=g["seed"], want_satisfied=True)
    u, e = bitsource.philox_bitstreams(g["seed"], seg["stream"], 0, ntr, T, n, thr)
    u[:, g["N"]:] = 1                                    # info bits beyond N must be ignored by the kernel
    t_bs, s_bs = ctx.run([seg], bits=bitsource.pack_bitstreams(u, e), want_satisfied=True)
    assert np.array_equal(s_ph, s_bs) and int(t_ph[0]) == int(t_bs[0])
    template = [tuple(t) for t in g["template"]]
    wins = 0
    for trial in range(ntr):
        dec, sat, _ = ref_port.parity_trial(g["gens2"], g["m"], template, g["gamma"], g["N"], g["p"], g["seed"], seg["stream"], trial)
        assert sat == int(s_ph[trial])
        wins += not dec
    assert wins == int(t_ph[0])
    ctx.close()


@pytest.mark.gpu
def test_gpu_parity_experiment_properties(tmp_path, monkeypatch):
    """run_parity_experiment: noiseless H1 always satisfies its own template; results do not depend on how the
    trial range is split; __main__ prints the reference's lines and writes the CSV plots_compare.py expects."""
    import runpy
    import pandas as pd
    import comp_parity as cp
    import parity_eqn_check as pec
    g1 = [[pec.parse_poly_token("7")], [pec.parse_poly_token("5")]]
    g2 = [[pec.parse_poly_token("6")], [pec.parse_poly_token("5")]]
    df = cp.run_parity_experiment(g1, g2, 2, [50, 200], [0.0, 0.1, 0.5], 0.6, 20000, 5)
    assert list(df.columns) == ["N", "p", "Pd", "Pc"] and df["N"].tolist() == [50, 50, 50, 200, 200, 200]
    assert df["Pd"][0] == 1.0 and df["Pd"][3] == 1.0                      # p = 0: every check holds under H1
    assert abs(df["Pd"][2] - df["Pd"][5]) < 0.2 and df["Pd"][5] < 0.05   # p = 1/2: fraction ~ 1/2 < gamma
    d1, d2, d3 = {}, {}, {}
    cp.run_parity_experiment(g1, g2, 2, [200], [0.1], 0.6, 5000, 5, details=d1)
    cp.run_parity_experiment(g1, g2, 2, [200], [0.1], 0.6, 2000, 5, details=d2)
    cp.run_parity_experiment(g1, g2, 2, [200], [0.1], 0.6, 3000, 5, trial_offset=2000, details=d3)
    assert np.array_equal(d1["tallies"], d2["tallies"] + d3["tallies"])
    monkeypatch.chdir(tmp_path)
    runpy.run_module("comp_parity", run_name="__main__")
    out = pd.read_csv(tmp_path / "results_parity" / "Pd_parity_results.csv")
    assert list(out.columns) == ["N", "p", "Pd", "Pc"] and len(out) == 14
    assert ((out["Pc"] >= 0) & (out["Pc"] <= 1)).all()
