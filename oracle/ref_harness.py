"""ref_harness.py -- run the REFERENCE'S OWN ``run_experiment`` (Pd_plotter.py:176-235), unmodified, as the
timed CPU baseline.

TEST INFRASTRUCTURE ONLY: used by bench.py's ``--impl reference`` arm and ``cpu_baseline`` leg and by tests/.

The reference's two files are imported from ``oracle/_ref/`` (placed there by ``oracle/fetch_ref.py``; they
are not part of this repository) or, in the build container, straight from ``/root/reference``.  What is
supplied from outside, because the reference does not ship it or because it is not installed:

* ``matplotlib`` / ``matplotlib.pyplot``: empty stub modules (imported at Pd_plotter.py:58, never used on this path);
* ``viterbi_markov.simulate_markov_sequence`` (called at Pd_plotter.py:149,212,219, defined nowhere -- SURVEY F2):
  the driver of ``oracle/ref_port.py`` running on the reference's OWN ``branch_output_and_next_state``,
  ``viterbi_metric_step`` and ``build_trellis`` (decoder fixed to gen1), fed by the MVD-PHILOX-2 bits;
* ``tqdm``: replaced by the identity (no progress bar on stderr).

``build_symbolic_T`` + ``evaluate_symbolic_T`` (3.3 s of sympy at S = 31, the same result every call) are
memoised across calls and their time is reported separately, as BASELINE.md section 3 step 2 prescribes; the
reference's own ``lru_cache`` keeps the learned P1 (Pd_plotter.py:123), so after :func:`Session.warm` a call of
``run`` is exactly the trial loop Pd_plotter.py:198-226.
"""
from __future__ import annotations

import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LOCAL = os.path.join(HERE, "_ref")
REF_CONTAINER = "/root/reference"


def ref_dir():
    """Directory holding the reference's viterbi_markov.py / Pd_plotter.py, or None."""
    for d in (REF_LOCAL, REF_CONTAINER):
        if all(os.path.exists(os.path.join(d, f)) for f in ("viterbi_markov.py", "Pd_plotter.py")):
            return d
    return None


def _import_reference(path):
    """Import the reference's two modules under private names (the product ships drop-ins with the same names)."""
    import importlib.util
    import warnings
    stubbed = [name for name in ("matplotlib", "matplotlib.pyplot") if name not in sys.modules]
    for name in stubbed:
        sys.modules[name] = types.ModuleType(name)
    warnings.simplefilter("ignore")                       # invalid escape sequences in the reference's docstrings
    saved = {k: sys.modules.get(k) for k in ("viterbi_markov", "Pd_plotter")}
    try:
        mods = {}
        for name in ("viterbi_markov", "Pd_plotter"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(path, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod                      # Pd_plotter does `import viterbi_markov as vm`
            spec.loader.exec_module(mod)
            mods[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for name in stubbed:                             # the stubs must not outlive the import (other code probes matplotlib)
            sys.modules.pop(name, None)
    return mods["viterbi_markov"], mods["Pd_plotter"]


class Session:
    """One process's view of the reference: modules loaded, simulator injected, symbolic T memoised."""

    def __init__(self, path=None):
        import ref_port
        self.path = path or ref_dir()
        if self.path is None:
            raise FileNotFoundError("reference sources not found (run oracle/fetch_ref.py in the build container)")
        self.vm, self.pdp = _import_reference(self.path)
        assert not hasattr(self.vm, "simulate_markov_sequence"), "reference now ships a simulator: re-survey"
        self.ref_port = ref_port
        self.symbolic_s = 0.0
        self.num_iter = 1
        self.trial_offset = 0
        self.calls = 0
        self.seed = 0
        self.gen1 = None
        self._memo = {}
        vm, pdp = self.vm, self.pdp
        pdp.tqdm = lambda it, **kw: it
        orig_build, orig_eval = vm.build_symbolic_T, pdp.evaluate_symbolic_T

        def build_memo(states, transitions, all_r, *a, **kw):
            key = ("T", tuple(states))
            if key not in self._memo:
                t0 = time.perf_counter()
                self._memo[key] = orig_build(states, transitions, all_r, *a, **kw)
                self.symbolic_s += time.perf_counter() - t0
            return self._memo[key]

        def eval_memo(T_sym, p_sym, p_val):
            key = ("E", id(T_sym), float(p_val))
            if key not in self._memo:
                t0 = time.perf_counter()
                self._memo[key] = orig_eval(T_sym, p_sym, p_val)
                self.symbolic_s += time.perf_counter() - t0
            return self._memo[key]

        vm.build_symbolic_T = build_memo
        pdp.evaluate_symbolic_T = eval_memo
        vm.simulate_markov_sequence = self._simulate

    # the injected simulator: same keying as oracle/make_golden.py::Injected
    def _simulate(self, generator_matrix, m, k, n, length, p_val, random_input=True, seed=None):
        vm = self.vm
        kw = dict(step=vm.viterbi_metric_step, branch_fn=vm.branch_output_and_next_state, trellis_fn=vm.build_trellis,
                  decoder_matrix=self.gen1)
        if seed is not None:                                    # learning chain, Pd_plotter.py:149-155
            self.seed = seed
            self.learn_steps += int(length)
            return self.ref_port.simulate_markov_sequence(generator_matrix, m, k, n, length, p_val, random_input, seed,
                                                          stream=self.ref_port.LEARN_STREAM, trial=0, **kw)
        c = self.calls                                          # trial loop, Pd_plotter.py:212,219
        self.calls += 1
        point, within = divmod(c, 2 * self.num_iter)
        trial, hyp = divmod(within, 2)
        self.trial_steps += int(length)
        return self.ref_port.simulate_markov_sequence(generator_matrix, m, k, n, length, p_val, random_input, self.seed,
                                                      stream=2 * point + hyp, trial=self.trial_offset + trial, **kw)

    def run(self, k, n, m, gen1, gen2, num_iter, p_vec, learn_len, learn_burn, laplace, seed, N_spectrum=None,
            trial_offset=0):
        """The reference's ``run_experiment(...)`` with its positional signature; ``N_spectrum`` sets the
        configuration dict ``N_SPECTRUM_BY_M[m]`` (Pd_plotter.py:78-83) for the call.  Returns a dict with the
        DataFrame, wall seconds, trial-loop / learning steps executed and the sympy seconds spent inside."""
        pdp = self.pdp
        self.gen1, self.num_iter, self.trial_offset, self.calls = gen1, int(num_iter), int(trial_offset), 0
        self.seed = seed
        self.trial_steps = self.learn_steps = 0
        sym0 = self.symbolic_s
        saved = dict(pdp.N_SPECTRUM_BY_M)
        if N_spectrum is not None:
            pdp.N_SPECTRUM_BY_M[m] = list(N_spectrum)
        try:
            t0 = time.perf_counter()
            df = pdp.run_experiment(k, n, m, gen1, gen2, num_iter, p_vec, learn_len, learn_burn, laplace, seed)
            wall = time.perf_counter() - t0
        finally:
            pdp.N_SPECTRUM_BY_M.clear()
            pdp.N_SPECTRUM_BY_M.update(saved)
        return dict(df=df, wall_s=wall, trial_steps=self.trial_steps, learn_steps=self.learn_steps,
                    symbolic_s=self.symbolic_s - sym0)

    def warm(self, k, n, m, gen1, gen2, p_vec, learn_len, learn_burn, laplace, seed, N_spectrum=None):
        """One single-iteration call: pays the sympy build and fills the reference's lru_cache of learned P1, so that
        later calls with the same arguments are the trial loop only."""
        return self.run(k, n, m, gen1, gen2, 1, p_vec, learn_len, learn_burn, laplace, seed, N_spectrum=N_spectrum,
                        trial_offset=1 << 40)


# ----------------------------------------------------------------------------- multiprocessing front end (bench.py)
_SESSION = None


def _worker_init():
    global _SESSION
    _SESSION = Session()


def _worker_warm(args):
    r = _SESSION.warm(*args)
    return dict(wall_s=r["wall_s"], symbolic_s=r["symbolic_s"], learn_steps=r["learn_steps"], trial_steps=r["trial_steps"])


def _worker_run(args):
    cfg, num_iter, offset = args
    k, n, m, gen1, gen2, p_vec, learn_len, learn_burn, laplace, seed, N_spectrum = cfg
    r = _SESSION.run(k, n, m, gen1, gen2, num_iter, p_vec, learn_len, learn_burn, laplace, seed, N_spectrum=N_spectrum,
                     trial_offset=offset)
    return dict(wall_s=r["wall_s"], trial_steps=r["trial_steps"], learn_steps=r["learn_steps"], symbolic_s=r["symbolic_s"],
                rows=r["df"].to_dict(orient="records"))


class Pool:
    """``cores`` worker processes, each holding the reference warmed for one configuration; :meth:`step` runs
    ``iters_per_core`` Monte-Carlo iterations of the whole sweep on every core (distinct global trial ids) and
    returns the executed trellis steps and the wall time of the slowest worker (= the step's wall)."""

    def __init__(self, cfg, cores):
        import multiprocessing as mp
        self.cfg, self.cores = cfg, cores
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_worker_init)
        k, n, m, gen1, gen2, p_vec, learn_len, learn_burn, laplace, seed, N_spectrum = cfg
        t0 = time.perf_counter()
        warm = self.pool.map(_worker_warm, [(k, n, m, gen1, gen2, p_vec, learn_len, learn_burn, laplace, seed, N_spectrum)] * cores,
                             chunksize=1)
        self.warm_wall_s = time.perf_counter() - t0
        self.symbolic_s = max(w["symbolic_s"] for w in warm)
        self.learn_steps = warm[0]["learn_steps"]
        self.next_offset = 0

    def step(self, iters_per_core):
        jobs = [(self.cfg, iters_per_core, self.next_offset + i * iters_per_core) for i in range(self.cores)]
        self.next_offset += self.cores * iters_per_core
        t0 = time.perf_counter()
        res = self.pool.map(_worker_run, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        assert all(r["learn_steps"] == 0 and r["symbolic_s"] == 0.0 for r in res), "a timed step relearned / rebuilt T"
        return dict(steps=sum(r["trial_steps"] for r in res), wall_s=wall, rows=res[0]["rows"])

    def close(self):
        self.pool.close()
        self.pool.join()
