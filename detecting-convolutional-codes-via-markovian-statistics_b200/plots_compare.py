"""plots_compare -- drop-in for the reference's comparison plots (plots_compare.py:35-148): P_err = 1 - Pc of the
hybrid detector (CSV of ``Pd_plotter.py``) against the parity-template baseline (CSV of ``comp_parity.py``) versus p
at fixed N and versus N at fixed p.  Presentation only -- no GPU work; matplotlib is imported when ``main`` runs and a
table is printed instead when it is not installed.  Same helper names, CSV columns (``N,p,Pd,Pc``; ``Pd`` stands in
for a missing ``Pc`` as in the reference, :78-84) and output file names (``Perr_vs_p_N{N}.png``, ``Perr_vs_N_p{p}.png``).

    python plots_compare.py --hybrid results_experiments/Pd_hybrid_results.csv --baseline results_parity/Pd_parity_results.csv
"""
from __future__ import annotations

import argparse
import os

import numpy as np


def p_error(Pc):
    """P_err = 1 - Pc, clipped to [0, 1] (reference :35-42)."""
    return np.clip(1.0 - np.asarray(Pc, dtype=float), 0.0, 1.0)


def _curve(rows, x_col, y_col):
    if rows.empty:
        return np.array([]), np.array([])
    rows = rows.sort_values(by=x_col)
    return rows[x_col].to_numpy(), rows[y_col].to_numpy()


def extract_by_N(df, N, x_col="p", y_col="Pc"):
    """(x, y) at fixed blocklength N, sorted by x (reference :49-56)."""
    return _curve(df[df["N"] == int(N)], x_col, y_col)


def extract_by_p(df, p, x_col="N", y_col="Pc"):
    """(x, y) at fixed crossover probability p (``np.isclose``), sorted by x (reference :59-66)."""
    return _curve(df[np.isclose(df["p"], p)], x_col, y_col)


def load_results(hybrid_csv, baseline_csv):
    import pandas as pd

    frames = []
    for path in (hybrid_csv, baseline_csv):
        df = pd.read_csv(path)
        if "Pc" not in df.columns and "Pd" in df.columns:        # reference :78-84
            df["Pc"] = df["Pd"]
        frames.append(df)
    return frames


def main(hybrid_csv, baseline_csv, outdir):
    os.makedirs(outdir, exist_ok=True)
    df_h, df_b = load_results(hybrid_csv, baseline_csv)
    Ns = sorted(set(df_h["N"]).union(df_b["N"]))
    ps = sorted(set(df_h["p"]).union(df_b["p"]))
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        plt = None
        print("matplotlib is not installed -- P_err tables instead of plots:")
    sweeps = ([("p", N, extract_by_N, f"N={N}", "BSC crossover probability p", f"Perr_vs_p_N{N}.png") for N in Ns] +
              [("N", p, extract_by_p, f"p={p}", "Blocklength N", f"Perr_vs_N_p{p}.png") for p in ps])
    for xname, fixed, pick, tag, xlabel, fname in sweeps:
        xh, yh = pick(df_h, fixed)
        xb, yb = pick(df_b, fixed)
        if plt is None:
            print(f"\n{tag}: {xname} -> P_err hybrid {dict(zip(xh.tolist(), p_error(yh).tolist()))}  "
                  f"parity baseline {dict(zip(xb.tolist(), p_error(yb).tolist()))}")
            continue
        plt.figure(figsize=(6, 5))
        if len(xh):
            plt.plot(xh, p_error(yh), marker="o", label=f"Hybrid ({tag})")
        if len(xb):
            plt.plot(xb, p_error(yb), marker="s", linestyle="--", label=f"Parity baseline ({tag})")
        plt.xlabel(xlabel)
        plt.ylabel("Probability of error $P_{\\mathrm{err}}$")
        plt.title(f"$P_{{\\mathrm{{err}}}}$ vs ${xname}$ ({tag})")
        plt.grid(True)
        plt.legend()
        plt.savefig(os.path.join(outdir, fname), dpi=200, bbox_inches="tight")
        plt.close()


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Compare hybrid and parity-template detectors")
    parser.add_argument("--hybrid", required=True, help="CSV from Pd_plotter.py")
    parser.add_argument("--baseline", required=True, help="CSV from comp_parity.py")
    parser.add_argument("--outdir", default="plots", help="Output directory for plots")
    args = parser.parse_args()
    main(args.hybrid, args.baseline, args.outdir)
