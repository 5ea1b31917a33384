"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line on stdout with the
keys the driver reads; without a CUDA device the product arm refuses to run (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-iters", "4"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "trellis-steps/s" and d["value"] > 0 and d["vs_baseline"] is None
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness
    want_kind = "reference" if ref_harness.ref_dir() else "port"      # the reference's own files when oracle/_ref holds them
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["config"]["trials_per_point_per_gpu"] == 1_000_000      # same config dict as the GPU arm (driver: same_config)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], capture_output=True,
                         text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith("{")]


def test_reference_harness_reproduces_the_golden_run():
    """oracle/ref_harness.py (what the CPU arm times) drives the reference's unmodified run_experiment exactly like
    oracle/make_golden.py did when the fixtures were frozen: same rows on the smallest golden experiment."""
    import pytest
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness
    if ref_harness.ref_dir() is None:
        pytest.skip("reference sources not available (oracle/_ref is created by oracle/fetch_ref.py in the build container)")
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "experiments.json")))["c75_c65_lap"]
    s = ref_harness.Session()
    r = s.run(g["k"], g["n"], g["m"], g["gen1"], g["gen2"], g["num_iter"], g["p_vec"], g["learn_len"], g["learn_burn"],
              g["laplace"], g["seed"], N_spectrum=g["N_list"])
    assert r["df"].to_dict(orient="records") == g["rows"]
    assert r["trial_steps"] == 2 * g["num_iter"] * len(g["p_vec"]) * sum(g["N_list"])
    assert r["learn_steps"] == len(g["p_vec"]) * g["learn_len"]
    again = s.run(g["k"], g["n"], g["m"], g["gen1"], g["gen2"], g["num_iter"], g["p_vec"], g["learn_len"], g["learn_burn"],
                  g["laplace"], g["seed"], N_spectrum=g["N_list"])
    assert again["learn_steps"] == 0 and again["symbolic_s"] == 0.0      # lru_cache / memo: a repeated call is the trial loop only
    assert again["df"].to_dict(orient="records") == g["rows"]
