"""demo_script -- drop-in for the reference's interactive demo (reference demo_script.py:35-163).

Same prompts, same predefined code pairs (``EXAMPLE_CODES``), same experiment parameters
(``num_iter=2000``, ``p_vec=[0.01, 0.05, 0.1, 0.2, 0.3]``, ``seed=123``); the detector runs on the
GPU through :func:`Pd_plotter.run_experiment`.  Plotting is optional: matplotlib is presentation
only and is skipped (with the table printed instead) when it is not installed.
"""
from __future__ import annotations

from Pd_plotter import run_experiment

# Reference demo_script.py:35-52 (the "(15,13)" label is the reference's; key on the tap lists)
EXAMPLE_CODES = {
    "1": {
        "name": "Rate-1/2, m=2 (7,5) vs (6,5)",
        "k": 1, "n": 2, "m": 2,
        "gen1": [[[1, 1, 1]], [[1, 0, 1]]],
        "gen2": [[[1, 1, 0]], [[1, 0, 1]]],
    },
    "2": {
        "name": "Rate-1/2, m=3 (15,13) vs (13,15)",
        "k": 1, "n": 2, "m": 3,
        "gen1": [[[1, 1, 1, 1]], [[1, 0, 1, 1]]],
        "gen2": [[[1, 0, 1, 1]], [[1, 1, 1, 1]]],
    },
}

DEMO_P_VEC = [0.01, 0.05, 0.1, 0.2, 0.3]
DEMO_NUM_ITER = 2000


def read_generators(k, n, m, label):
    """Prompt for n tap vectors of length m+1 (reference demo_script.py:58-75)."""
    print(f"\nEnter generator polynomials for {label}")
    print(f"Format: {k} tap vectors per output, each of length {m+1}")
    print("Example (rate 1/2): 1,1,1")
    gens = []
    for j in range(n):
        taps = None
        while taps is None:
            text = input(f"  Output v{j}: ").strip()
            try:
                cand = [int(tok) for tok in text.split(",")]
                if len(cand) != m + 1:
                    raise ValueError("wrong length")
                taps = cand
            except Exception:
                print("  Invalid format. Try again.")
        gens.append([taps])
    return gens


def choose_codes():
    print("\n=== Convolutional Code Detector Demo ===\n")
    print("Choose an option:")
    print("  [1] Use predefined example codes")
    print("  [2] Enter custom codes manually")
    mode = input("Selection [1/2]: ").strip()
    if mode == "1":
        print("\nAvailable examples:")
        for key, info in EXAMPLE_CODES.items():
            print(f"  [{key}] {info['name']}")
        cfg = EXAMPLE_CODES[input("Select example: ").strip()]
        return cfg["k"], cfg["n"], cfg["m"], cfg["gen1"], cfg["gen2"]
    k = int(input("Enter k (inputs per time step): "))
    n = int(input("Enter n (outputs per time step): "))
    m = int(input("Enter m (memory): "))
    return k, n, m, read_generators(k, n, m, "Code-1 (H1)"), read_generators(k, n, m, "Code-2 (H2)")


def run_demo(k, n, m, gen1, gen2, num_iter=DEMO_NUM_ITER, p_vec=None, **kw):
    """The demo's experiment call (reference demo_script.py:119-131)."""
    return run_experiment(k=k, n=n, m=m, gen1=gen1, gen2=gen2, num_iter=num_iter,
                          p_vec=list(DEMO_P_VEC if p_vec is None else p_vec),
                          learn_len=None, learn_burn=200, laplace=1.0, seed=123, **kw)


def plot_results(df):
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        print("\nmatplotlib is not installed -- results table instead of plots:\n")
        print(df.to_string(index=False))
        return
    for xcol, curve, xlabel, title in (("p", "N", "BSC crossover probability p", "Hybrid detector: $P_d$ vs $p$"),
                                       ("N", "p", "Blocklength N", "Hybrid detector: $P_d$ vs $N$")):
        plt.figure(figsize=(6, 5))
        for val in sorted(df[curve].unique()):
            part = df[df[curve] == val]
            plt.plot(part[xcol], part["Pd"], marker="o", label=f"{curve}={val}")
        plt.xlabel(xlabel)
        plt.ylabel("Probability of detection $P_d$")
        plt.title(title)
        plt.grid(True)
        plt.legend()
        plt.show()


if __name__ == "__main__":
    k, n, m, gen1, gen2 = choose_codes()
    print("\nRunning hybrid detector on the GPU...\n")
    plot_results(run_demo(k, n, m, gen1, gen2))
