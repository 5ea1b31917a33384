#!/usr/bin/env python3
"""bench.py -- trellis-steps/sec of the hybrid-detector Monte-Carlo sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--engine acs|fsm] [--trials T]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own run_experiment on host cores

One "step" = one pass of the paper sweep of Pd_plotter.py (BASELINE configs[1]): codes (7,5) vs
(6,5), N = 500, p in {0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5}, both hypotheses, ``--trials`` Monte-Carlo
trials per point *per GPU* (default 10^6, the north-star target; the reference ships 10^4).
That is 2 * 500 * 7 * trials trellis steps per GPU per step.

value : device-resident throughput -- tables already on the GPU, on-device Philox bit source, one
        kernel launch per step; wall time between synchronised brackets, max over ranks (weak scaling).
e2e   : the same sweep through the public API ``Pd_plotter.run_experiment`` from host data, nothing cached:
        learning chains (GPU), Laplace/normalise (host), table upload H2D, detection launch, allreduce,
        tallies D2H, DataFrame.  Bytes are counted at libmvd's copy call sites (mvd_copy_stats).
roofline : the binding unit is the SM integer pipe, peak measured on this GPU in this run
        with libmvd's LOP3/IMAD micro-kernel; the HBM view of the bit-stream kernel is reported too.
strong : a fixed total of --strong-trials per point sharded over the N ranks through run_experiment (models
        cached, tallies reduced on the device); on N > 1 rank 0 also runs the whole range alone: the tallies
        must be identical (tallies_equal_1gpu) and efficiency_vs_n1 = t_1 / (N t_N).
paper_sweep : wall time of the north-star sweeps (35-point Pd-vs-p, 16-point Pd-vs-N) through run_experiment.
sustained : the resident loop repeated for >= --sustain-s seconds with clocks and power sampled.
cpu_baseline : the reference's own ``run_experiment`` (oracle/_ref, injected simulator) on this box's host
        cores, bounded sample; falls back to the Python port (oracle/ref_port.py) when oracle/_ref is absent.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200")
for _p in (PKG, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GEN1 = [[[1, 1, 1]], [[1, 0, 1]]]          # (7,5)  Pd_plotter.py:247
GEN2 = [[[1, 1, 0]], [[1, 0, 1]]]          # (6,5)  Pd_plotter.py:248
P_VEC = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]   # Pd_plotter.py:69
N_BLOCK = 500                               # Pd_plotter.py:80
SEED = 12345                                # Pd_plotter.py:70
K, NOUT, M = 1, 2, 2
OPS_CORE = 5 * (1 << M) + 11                # 31 int-ops / step at m = 2
METRIC = "trellis-steps/sec (whole box) for Pd-vs-p Monte-Carlo sweep"
SWEEP_P_NS = [50, 100, 200, 500, 1000]      # scripts/paper_sweeps.py: Pd vs p (x 7 p = 35 points)
SWEEP_N_NS = [100, 200, 500, 1000, 2000, 5000, 10000, 100000]   # Pd vs N (x 2 p = 16 points), BASELINE config 3
SWEEP_N_PS = [0.05, 0.1]


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock, power and throttle reasons with NVML while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80, "sync_boost": 0x10}

    def __init__(self, index: int, period_s: float = 0.02):
        self.samples, self.power, self.reasons, self.max_mhz, self.ok = [], [], 0, None, False
        self.period = period_s
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                try:
                    self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.ok:
            self.thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.ok:
            self.thread.join(timeout=1)

    def summary(self) -> dict:
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        names = [n for n, bit in {**self.BAD, **self.NOTE}.items() if self.reasons & bit]
        out = {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": names, "samples": len(s)}
        if self.power:
            out["power_w_max"] = max(self.power)
        return out


# --------------------------------------------------------------------------------------------- CPU arm
def _port_worker(args):
    import ref_port
    gen1, gen2, N, p_vec, iters, seed, offset = args
    t0 = time.perf_counter()
    rows = ref_port.run_experiment(K, NOUT, M, gen1, gen2, iters, p_vec, None, 200, 1.0, seed, N_spectrum=[N], trial_offset=offset)
    return 2 * N * iters * len(p_vec), time.perf_counter() - t0, len(rows)


class PortPool:
    """Fallback when oracle/_ref is absent: the pure-Python port of the reference (oracle/ref_port.py).  Every step
    re-learns the 7 chains of 6 200 steps inside the timed sample; those steps are counted."""
    kind = "port"

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores)
        self.next_offset = 0
        self.symbolic_s = 0.0
        self.warm_wall_s = 0.0

    def step(self, iters):
        jobs = [(GEN1, GEN2, N_BLOCK, P_VEC, iters, SEED, self.next_offset + i * iters) for i in range(self.cores)]
        self.next_offset += self.cores * iters
        t0 = time.perf_counter()
        res = self.pool.map(_port_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        learn = len(P_VEC) * 6200 * self.cores
        return dict(steps=sum(r[0] for r in res) + learn, wall_s=wall)

    def close(self):
        self.pool.close()
        self.pool.join()


def make_cpu_pool(cores):
    """The reference's own run_experiment when its files are available (oracle/_ref, else /root/reference)."""
    import ref_harness
    if ref_harness.ref_dir() is not None:
        cfg = (K, NOUT, M, GEN1, GEN2, P_VEC, None, 200, 1.0, SEED, [N_BLOCK])
        pool = ref_harness.Pool(cfg, cores)
        pool.kind = "reference"
        return pool
    return PortPool(cores)


def cpu_sample_text(pool, iters, steps, wall):
    what = ("the reference's unmodified Pd_plotter.run_experiment from oracle/_ref (simulate_markov_sequence injected on its "
            "own branch/step/trellis functions, matplotlib stubbed, sympy T(p) memoised: %.1f s, and P1 learning, lru_cached by "
            "the reference itself, both outside the timed sample)" % pool.symbolic_s) if pool.kind == "reference" else \
           "oracle/ref_port.py (pure-Python port; learning chains inside the sample and counted)"
    return (f"(7,5)/(6,5) N={N_BLOCK} p_vec={P_VEC}, {iters} Monte-Carlo iterations x 2 hypotheses x 7 p per core on {pool.cores} "
            f"cores = {steps} trellis steps in {wall:.1f} s; {what}")


def cpu_throughput(cores, iters, reps=1):
    pool = make_cpu_pool(cores)
    try:
        pool.step(max(1, iters // 8))                            # untimed warm-up pass
        steps = wall = 0
        for _ in range(reps):
            r = pool.step(iters)
            steps += r["steps"]
            wall += r["wall_s"]
    finally:
        pool.close()
    return {"value": steps / wall, "unit": "trellis-steps/s", "cores": cores, "kind": pool.kind,
            "sample": cpu_sample_text(pool, iters, steps, wall), "sympy_T_build_s": pool.symbolic_s}


def c_oracle_throughput(ntrials: int = 2000) -> dict:
    import c_oracle as co
    from mvd import bitsource, codes
    tab_np = codes.enumerate_states(codes.freeze_generator(GEN1), M, K, NOUT)
    tab = co.Table(tab_np.metrics, M)
    T = bitsource.bsc_threshold(0.1)
    edge, _ = co.learn_chain([7, 5], [7, 5], NOUT, M, 6200, 200, T, SEED, bitsource.LEARN_STREAM, 0, tab)
    P1 = codes.p1_from_edge_counts(tab_np, edge, 1.0)
    Tref = codes.tref_half_table(tab_np)
    t0 = time.perf_counter()
    co.run_trials([7, 5], [7, 5], NOUT, M, N_BLOCK, T, SEED, 0, 0, ntrials, tab, P1, Tref, 0)
    co.run_trials([7, 5], [3, 5], NOUT, M, N_BLOCK, T, SEED, 1, 0, ntrials, tab, P1, Tref, 1)
    wall = time.perf_counter() - t0
    return {"value": 2 * ntrials * N_BLOCK / wall, "unit": "trellis-steps/s", "cores": 1, "kind": "port-c",
            "sample": f"oracle/mvd_oracle.c, {ntrials} trials x 2 hypotheses, N={N_BLOCK}, p=0.1"}


def run_reference_arm(args):
    """``--impl reference``: the reference's CPU implementation of the path on all host cores, same config / metric /
    unit as the GPU arm; every step is a bounded sample (``--ref-iters`` iterations of the whole sweep per core)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    iters = max(1, args.ref_iters)
    pool = make_cpu_pool(cores)
    try:
        for _ in range(args.warmup):
            pool.step(max(1, iters // 8))
        res = []
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res.append(pool.step(iters))
        wall = time.perf_counter() - t0
    finally:
        pool.close()
    steps = sum(r["steps"] for r in res)
    value = steps / sum(r["wall_s"] for r in res)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "trellis-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "trellis-steps/s", "cores": cores, "kind": pool.kind,
                             "sample": cpu_sample_text(pool, iters, res[-1]["steps"], res[-1]["wall_s"]) + " per step",
                             "sympy_T_build_s": pool.symbolic_s, "warm_wall_s": pool.warm_wall_s},
            "e2e": {"value": value, "unit": "trellis-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def workload_config(args):
    """Identical for both arms (the driver compares the dicts): the workload, not how an arm samples it."""
    return {"workload": "Pd_plotter.py paper sweep (BASELINE configs[1]): (7,5) vs (6,5), k=1 n=2 m=2, S=31 Markov states, "
                        f"N={N_BLOCK}, p_vec={P_VEC}, both hypotheses, learn_len=6200 burn=200 laplace=1",
            "trials_per_point_per_gpu": int(args.trials), "engine": args.engine,
            "bit_source": "on-device Philox4x32-10, position-addressed (MVD-PHILOX-2)",
            "l2_policy": "no input stream to cache: bits are generated in registers, tables (<4 KB) live in shared "
                         "memory; the bit-stream variant reads > 2 GB per step (>> 126 MB L2)",
            "parallelism": f"trial-sharded x{args.gpus}"}


# --------------------------------------------------------------------------------------------- GPU arm
_REAL_STDOUT = None


def _claim_stdout():
    """Everything libraries print (NCCL's version banner, warnings) goes to stderr; the one JSON line is written to
    the process's original stdout by :func:`_emit`."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _sha(t) -> str:
    import numpy as np
    return hashlib.sha256(np.ascontiguousarray(np.asarray(t, dtype=np.int64)).tobytes()).hexdigest()[:16]


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default="acs", choices=["acs", "fsm"])
    ap.add_argument("--trials", type=int, default=1_000_000, help="Monte-Carlo trials per (N,p) point per GPU")
    ap.add_argument("--strong-trials", type=int, default=1_000_000, help="strong-scaling / paper-sweep legs: trials per point in total")
    ap.add_argument("--sustain-s", type=float, default=10.0, help="length of the sustained leg (0 = skip)")
    ap.add_argument("--ref-iters", type=int, default=30, help="CPU arm: iterations of the whole sweep per core per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the alternate-engine / bitstream / sweep / sustained legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=300))

    import Pd_plotter as pdp
    import viterbi_markov as vm
    from mvd import bitsource, codes
    from mvd.engine import Seg

    det = vm._detector(codes.freeze_generator(GEN1), K, NOUT, M, local_rank)
    t1, t2 = det.taps_of(GEN1), det.taps_of(GEN2)
    trials = int(args.trials)
    begin, end = rank * trials, (rank + 1) * trials            # weak scaling: fixed work per GPU
    steps_per_pass_rank = 2 * N_BLOCK * len(P_VEC) * trials
    steps_per_pass = steps_per_pass_rank * world

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident setup (outside the timed region of `value`): learn P1, upload tables
    counts, tables = pdp._learn_edge_tables(det, P_VEC, None, 200, 1.0, SEED)

    def install_bench_models():
        """The 7 log-likelihood tables of the bench sweep (run_experiment legs in between install their own)."""
        det.set_models(tables)
        det._models_key = None

    install_bench_models()
    segs = []
    for q, p in enumerate(P_VEC):
        T = bitsource.bsc_threshold(p)
        segs.append(Seg(N=N_BLOCK, threshold=T, stream=2 * q, table=q, enc_taps=t1, decide=0, trial_begin=begin, trial_end=end))
        segs.append(Seg(N=N_BLOCK, threshold=T, stream=2 * q + 1, table=q, enc_taps=t2, decide=1, trial_begin=begin, trial_end=end))
    # The resident loop queues its launches: mvd_detect(device tallies only) returns once the sweep's work is on the
    # context's stream (MVD_OPT_ASYNC_DETECT), and for N > 1 the all_reduce of pass i is queued behind it on the same
    # stream as an asynchronous collective, so pass i + 1 starts while the 112-byte reduction is in flight.  Two tally
    # buffers alternate; a pass first makes the stream wait for the collective that last read its buffer.
    mvd_stream = torch.cuda.Stream()
    det.set_stream(mvd_stream.cuda_stream)
    main._keep = mvd_stream                                      # the context runs on it until the process ends
    d_bufs = [torch.zeros(len(segs), dtype=torch.int64, device="cuda") for _ in range(2)]
    works = [None, None]
    passes = [0]
    torch.cuda.synchronize()

    def device_pass(engine):
        b = passes[0] & 1
        passes[0] += 1
        with torch.cuda.stream(mvd_stream):
            if works[b] is not None:
                works[b].wait()
                works[b] = None
            det.detect(segs, seed=SEED, engine=engine, d_tallies_ptr=d_bufs[b].data_ptr(), host_tallies=False)
            if world > 1:
                works[b] = dist.all_reduce(d_bufs[b], op=dist.ReduceOp.SUM, async_op=True)     # the one data-path collective
        return None

    def last_tallies():
        return d_bufs[(passes[0] - 1) & 1]

    def drain():
        det.synchronize()
        with torch.cuda.stream(mvd_stream):
            for b in (0, 1):
                if works[b] is not None:
                    works[b].wait()
                    works[b] = None
        mvd_stream.synchronize()

    def timed(fn, steps, warmup, queued=False):
        """`steps` calls of fn between two barriers; queued=True: fn only queues device work (drained before the
        closing barrier) and the kernel times come from the events libmvd recorded around every launch."""
        if queued:
            det.async_detect(True)
        for _ in range(warmup):
            fn()
        if queued:
            drain()
            det.async_stats()
        sync_all()
        l0 = det.launch_count()
        kms = []
        t0 = time.perf_counter()
        for _ in range(steps):
            kms.append(fn())
        if queued:
            drain()
        sync_all()
        dt = max_over_ranks(time.perf_counter() - t0)
        if queued:
            ms_sum, nl = det.async_stats()
            kms = [ms_sum / max(nl, 1)] * steps
            det.async_detect(False)
        return dt, kms, det.launch_count() - l0

    with ClockSampler(local_rank) as clk:
        dt, kms, launches = timed(lambda: device_pass(args.engine), args.steps, args.warmup, queued=True)
    clocks = clk.summary()
    value = steps_per_pass * args.steps / dt
    kernel_ms = float(np.mean(kms))
    final_tallies = last_tallies().cpu().numpy().copy()

    # ---- e2e through the public API (host data in, DataFrame out), every step, nothing cached
    def api_pass():
        det_details = {}
        df = pdp.run_experiment(K, NOUT, M, GEN1, GEN2, trials * world, P_VEC, None, 200, 1.0, SEED,
                                engine=args.engine, device=local_rank, details=det_details, cache_models=False)
        api_pass.df, api_pass.details = df, det_details
        return det_details["detect_kernel_ms"]

    for _ in range(args.warmup):
        api_pass()
    h2d0, d2h0 = det.copy_stats()
    dt_e2e, _, launches_e2e = timed(api_pass, args.steps, 0)
    h2d1, d2h1 = det.copy_stats()
    e2e_value = steps_per_pass * args.steps / dt_e2e
    h2d = (h2d1 - h2d0) / args.steps
    d2h = (d2h1 - d2h0) / args.steps + (8 * len(segs) if world > 1 else 0)     # + the reduced tallies read by torch (N > 1)
    # the API path must give the same tallies as the resident path (same seeds, same trial ids)
    api_t = api_pass.details["tallies"]
    same = bool(np.array_equal(np.asarray(api_t, dtype=np.int64), final_tallies.astype(np.int64)))

    line = {"metric": METRIC, "value": value, "unit": "trellis-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16x2 metrics + f64 log-likelihood", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": "trellis-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "bytes_counted": "libmvd cudaMemcpyAsync call sites (mvd_copy_stats), per rank",
                    "ms_per_step": 1e3 * dt_e2e / args.steps, "api": "Pd_plotter.run_experiment(cache_models=False)",
                    "tallies_equal_resident_path": same},
            "gpu_launches": launches, "gpu_launches_e2e": launches_e2e, "clocks": clocks,
            "kernel_ms_per_step": kernel_ms, "tallies_sha": _sha(final_tallies),
            "pd_pc": [{"p": p, "Pd": int(final_tallies[2 * q]) / (trials * world),
                       "Pc": (int(final_tallies[2 * q]) + int(final_tallies[2 * q + 1])) / (2 * trials * world)}
                      for q, p in enumerate(P_VEC)]}

    # ---- strong scaling + the north-star sweeps: a FIXED total of trials per point, sharded over the ranks
    def sweep(num_iter, p_vec, Ns, engine, shard=True, reps=3):
        d = {}
        run = lambda: pdp.run_experiment(K, NOUT, M, GEN1, GEN2, num_iter, p_vec, None, 200, 1.0, SEED, N_spectrum=Ns,
                                         engine=engine, device=local_rank, details=d, shard=shard)
        run()                                                    # learn / upload once: later calls hit the model cache
        best = None
        for _ in range(reps):
            if shard:
                sync_all()
            else:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            run()
            w = time.perf_counter() - t0
            w = max_over_ranks(w) if shard else w
            best = w if best is None else min(best, w)
        return best, d

    if not args.no_extras:
        T_total = int(args.strong_trials)
        t_n, d_n = sweep(T_total, P_VEC, [N_BLOCK], args.engine)
        strong = {"trials_total_per_point": T_total, "workload": "the bench sweep (7 p x N=500 x 2 hypotheses) through run_experiment, "
                  "models cached, tallies all-reduced on the device", "ms_per_sweep": 1e3 * t_n,
                  "detect_kernel_ms": d_n["detect_kernel_ms"], "tallies_sha": _sha(d_n["tallies"]),
                  "steps_per_s": d_n["steps"] / t_n}
        if world > 1:
            t_1, d_1 = (None, None)
            if rank == 0:                                        # the same range on ONE GPU (the others wait at the barrier)
                t_1, d_1 = sweep(T_total, P_VEC, [N_BLOCK], args.engine, shard=False)
            sync_all()
            if rank == 0:
                strong.update(ms_per_sweep_1gpu=1e3 * t_1, efficiency_vs_n1=t_1 / (world * t_n),
                              tallies_equal_1gpu=bool(np.array_equal(np.asarray(d_1["tallies"], dtype=np.int64),
                                                                     np.asarray(d_n["tallies"], dtype=np.int64))))
        else:
            strong.update(ms_per_sweep_1gpu=1e3 * t_n, efficiency_vs_n1=1.0, tallies_equal_1gpu=True)
        line["strong"] = strong
        ps = {"trials_per_point": T_total, "api": "Pd_plotter.run_experiment (models cached), best of 3",
              "pd_vs_p_points": len(SWEEP_P_NS) * len(P_VEC), "pd_vs_n_points": len(SWEEP_N_NS) * len(SWEEP_N_PS)}
        for eng in ("acs", "auto"):
            tp, dp = sweep(T_total, P_VEC, SWEEP_P_NS, eng)
            tn, dn = sweep(T_total, SWEEP_N_PS, SWEEP_N_NS, eng, reps=2)
            ps[eng] = {"pd_vs_p_ms": 1e3 * tp, "pd_vs_p_kernel_ms": dp["detect_kernel_ms"], "pd_vs_p_steps": dp["steps"],
                       "pd_vs_n_ms": 1e3 * tn, "pd_vs_n_kernel_ms": dn["detect_kernel_ms"], "pd_vs_n_steps": dn["steps"],
                       "pd_vs_p_sha": _sha(dp["tallies"]), "pd_vs_n_sha": _sha(dn["tallies"])}
        # BASELINE config 3 as the reference runs it: 10^4 iterations per point in total (few long trials per GPU: the time-split
        # path with exactly re-associated float64 sums, csrc/mvd_split.cuh)
        t3, d3 = sweep(10000, SWEEP_N_PS, SWEEP_N_NS, "auto")
        ps["config3_ref_iters"] = {"trials_per_point": 10000, "pd_vs_n_ms": 1e3 * t3, "pd_vs_n_kernel_ms": d3["detect_kernel_ms"],
                                   "pd_vs_n_steps": d3["steps"], "pd_vs_n_sha": _sha(d3["tallies"])}
        line["paper_sweep"] = ps

        # ---- sustained: the resident loop for >= sustain_s seconds, clocks and power sampled throughout
        install_bench_models()
        if args.sustain_s > 0:
            reps = max(args.steps, int(args.sustain_s / max(dt / args.steps, 1e-6)) + 1)
            with ClockSampler(local_rank, period_s=0.1) as sclk:
                dts, kms_s, _ = timed(lambda: device_pass(args.engine), reps, 1, queued=True)
            sc = sclk.summary()
            line["sustained"] = {"seconds": dts, "passes": reps, "value": steps_per_pass * reps / dts,
                                 "kernel_ms_per_step": float(np.mean(kms_s)), "sm_mhz_median": sc.get("sm_mhz"),
                                 "power_w_max": sc.get("power_w_max"), "reasons": sc.get("reasons"), "samples": sc.get("samples")}

    if rank == 0:
        # ---- roofline: integer pipe (binding) measured on this device, this run
        alu_gops, mixed_gops = det.int_peak()
        rate_kernel = steps_per_pass_rank / (kernel_ms * 1e-3)             # this GPU's kernel-only rate
        peaks, traffic, mix = {}, None, {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic = tj.get(f"{args.engine}_detect")
            mix = tj.get(f"{args.engine}_detect_instr_mix", {})
        except Exception:
            pass
        info = det.device_info()
        nominal = info["sm_count"] * 4 * 32 * (clocks.get("sm_max_mhz") or 1965) * 1e6 * 1e-9      # one warp instruction / cycle / SMSP
        ach = OPS_CORE * rate_kernel * 1e-9
        rng_ops = mix.get("bit_source_instr_per_step")
        line["roofline"] = {
            "bound": "int_alu", "unit": "Gop/s",
            "achieved": ach, "peak": mixed_gops, "frac": ach / mixed_gops,
            "peak_alu_pipe_only": alu_gops, "frac_alu_pipe_only": ach / alu_gops,
            "peak_nominal": nominal, "frac_of_nominal": ach / nominal,
            "frac_core_plus_rng": (OPS_CORE + rng_ops) * rate_kernel * 1e-9 / mixed_gops if rng_ops else None,
            # issued warp-instructions (ncu source page) x measured rate / issue-rate peak: what the schedulers actually did;
            # below frac because the 16x2 instructions do two of the survey's scalar ops each
            "issue_utilisation": mix.get("instr_per_step") * rate_kernel * 1e-9 / mixed_gops if mix.get("instr_per_step") else None,
            "ops_per_step": {"core": OPS_CORE, "f64_adds": 2, "rng_executed": rng_ops,
                             "instr_issued_per_step": mix.get("instr_per_step"), "instr_mix_source": mix.get("source"),
                             "note": "core = SURVEY 8(d): 2^(m+k+1) + 2^m + 3n + 5 at m=2, k=1, n=2; rng_executed = thread "
                                     "instructions per trellis step the kernel spends on Philox + lazy Bernoulli + encode "
                                     "(ncu source page, profiles/); frac counts core only, frac_core_plus_rng adds them"},
            "peak_source": "libmvd mvd_int_peak(), measured in this run on this GPU: `peak` = alternating LOP3/IMAD chains "
                           "(ALU + FMA pipes = warp-instruction issue rate, the most any integer code can retire); "
                           "`peak_alu_pipe_only` = LOP3-only chains (min/shift/logic/permute can only issue there); "
                           "`peak_nominal` = SMs x 4 schedulers x 32 lanes x max SM clock. "
                           "MEASURED_PEAKS.json has no integer figure; the path moves ~0 HBM bytes (bits are generated in "
                           "registers), so the HBM roofline does not bind it -- see roofline_hbm_bitstream for the HBM view",
            "kernel": ("detect2p_kernel (two trials/thread ACS)" if args.engine == "acs" else "detect2_kernel<FSM1> (one-load NEXT walk)"),
            "kernel_ms": kernel_ms, "steps_per_launch": steps_per_pass_rank,
            "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/)",
            "hbm_bytes_per_step_algorithmic": 0.0,
        }
        if not args.no_extras and world == 1:       # single-GPU legs only: nothing below may enter a collective
            other = "fsm" if args.engine == "acs" else "acs"
            install_bench_models()
            dto, kmo, _ = timed(lambda: device_pass(other), max(3, args.steps // 2), 3, queued=True)
            line["alt_engine"] = {"engine": other, "value": steps_per_pass * max(3, args.steps // 2) / dto,
                                  "kernel_ms_per_step": float(np.mean(kmo)),
                                  "tallies_equal": bool(np.array_equal(last_tallies().cpu().numpy(), final_tallies))}
            # ---- the other BASELINE code pairs through the same public API (kernel time of the detection launch, median of 3):
            # the demo's m = 3 pair (S = 435, tables in shared memory) and an m = 4 full detector (S = 25 751, tables in L2)
            oc = {}
            for name, g1, g2, mm, its, pv, ll in (
                    ("m3_demo_pair_17_13_vs_13_17", [[[1, 1, 1, 1]], [[1, 0, 1, 1]]], [[[1, 0, 1, 1]], [[1, 1, 1, 1]]], 3, 200000,
                     [0.01, 0.05, 0.1, 0.2, 0.3], None),
                    ("m4_31_33_vs_33_31", [[[1, 1, 0, 0, 1]], [[1, 1, 0, 1, 1]]], [[[1, 1, 0, 1, 1]], [[1, 1, 0, 0, 1]]], 4, 100000,
                     [0.05, 0.1], 200000)):
                try:
                    ent = {"trials_per_point": its, "N": N_BLOCK, "p_vec": pv}
                    for eng in ("acs", "auto"):
                        dd, ms = {}, []
                        for it in range(4):
                            pdp.run_experiment(1, 2, mm, g1, g2, its, pv, ll, 200, 1.0, SEED, N_spectrum=[N_BLOCK], engine=eng,
                                               device=local_rank, details=dd, shard=False)
                            if it:
                                ms.append(dd["detect_kernel_ms"])
                        km = float(np.median(ms))
                        ent[eng] = {"kernel_ms": km, "steps_per_s": dd["steps"] / (km * 1e-3), "kernel_kind": dd["kernel_kind"],
                                    "tallies_sha": _sha(dd["tallies"])}
                    ops = 5 * 2 ** mm + 11                                   # SURVEY 8(d) at k = 1, n = 2
                    ent.update(S=dd["S"], ops_per_step_core=ops, tallies_equal=ent["acs"]["tallies_sha"] == ent["auto"]["tallies_sha"],
                               frac_issue_peak_acs=ops * ent["acs"]["steps_per_s"] * 1e-9 / mixed_gops)
                    oc[name] = ent
                except Exception as exc:                                     # an extra leg must not take the bench line down
                    oc[name] = {"error": str(exc)[:200]}
            oc["note"] = ("acs = Eq. 4-5 in registers, two trials per thread (detect3p_kernel): m = 3 bound by shared-memory wavefronts / issue / "
                          "ALU pipe alike, m = 4 by the L1TEX gather rate (one 32-byte sector per trial-step); auto = NEXT-table walk; profiles/r03*")
            line["other_configs"] = oc
            install_bench_models()
            # bit-stream (verification-mode) kernel: HBM view.  10^5 trials/point -> 3 bits/step from HBM.
            bt = min(trials, 100_000)
            nsb = (N_BLOCK + 127) // 128
            words_per_seg = nsb * 3 * bt
            g = torch.Generator(device="cuda").manual_seed(1)
            bits = torch.randint(0, 2 ** 31 - 1, (len(segs) * words_per_seg, 4), dtype=torch.int32, device="cuda", generator=g)
            bsegs = [Seg(N=N_BLOCK, table=s.table, enc_taps=s.enc_taps, decide=s.decide, trial_begin=0, trial_end=bt,
                         bits_offset=i * words_per_seg) for i, s in enumerate(segs)]
            torch.cuda.synchronize()
            for _ in range(3):
                det.detect(bsegs, engine="fsm", bits_device_ptr=bits.data_ptr(), bits_words=bits.shape[0])
            ms = []
            for _ in range(5):
                det.detect(bsegs, engine="fsm", bits_device_ptr=bits.data_ptr(), bits_words=bits.shape[0])
                ms.append(det.last_kernel_ms())
            bsteps = len(segs) * bt * N_BLOCK
            bbytes = bits.numel() * 4
            bms = float(np.mean(ms))
            line["roofline_hbm_bitstream"] = {
                "bound": "hbm", "unit": "GB/s", "achieved": bbytes / (bms * 1e-3) * 1e-9,
                "peak": peaks.get("hbm_gbs"), "frac": (bbytes / (bms * 1e-3) * 1e-9) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                "bytes_per_step": bbytes / bsteps, "steps_per_s": bsteps / (bms * 1e-3), "kernel": "detect2_kernel<FSM1> bitstream (host-supplied bits resident in HBM, no RNG)",
                "note": "inputs 3 bits/step >> L2; this path is integer/LSU-bound, not HBM-bound"}
            del bits
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            base = cpu_throughput(cores, max(4, 2 * args.ref_iters))
            line["cpu_baseline"] = base
            line["cpu_baseline_1core"] = cpu_throughput(1, max(4, args.ref_iters // 2))       # as shipped: single-process
            try:
                line["cpu_baseline_c_oracle"] = c_oracle_throughput()
            except Exception as exc:     # the C oracle is optional for the bench
                line["cpu_baseline_c_oracle"] = {"error": str(exc)}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
