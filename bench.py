#!/usr/bin/env python3
"""bench.py -- trellis-steps/sec of the hybrid-detector Monte-Carlo sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--engine acs|fsm] [--trials T]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

One "step" = one pass of the paper sweep of Pd_plotter.py (BASELINE configs[1]): codes (7,5) vs
(6,5), N = 500, p in {0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5}, both hypotheses, ``--trials`` Monte-Carlo
trials per point *per GPU* (default 10^6, the north-star target; the reference ships 10^4).
That is 2 * 500 * 7 * trials trellis steps per GPU per step.

value : device-resident throughput -- tables already on the GPU, on-device Philox bit source, one
        kernel launch per step; wall time between synchronised brackets, max over ranks.
e2e   : the same sweep through the public API ``Pd_plotter.run_experiment`` from host data: learning
        chains (GPU), Laplace/normalise (host), table upload H2D, detection launch, tallies D2H,
        allreduce, DataFrame.
roofline : the binding unit is the SM integer pipe, peak measured on this GPU in this run
        with libmvd's IADD/LOP3 micro-kernel; the HBM view of the bit-stream kernel is reported too.
cpu_baseline : the reference's Python path (oracle/ref_port.py, reference data structures and
        math.log per step) on this box's host cores, bounded sample.
"""
from __future__ import annotations

import argparse
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


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80, "sync_boost": 0x10}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], 0, None, False
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
                    self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            self._stop.wait(0.05)

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
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": names, "samples": len(s)}


# --------------------------------------------------------------------------------------------- CPU arm
def _port_worker(args):
    import ref_port
    gen1, gen2, N, p, iters, seed, offset = args
    t0 = time.perf_counter()
    steps, tallies = ref_port.timed_steps(gen1, gen2, M, K, NOUT, N, p, iters, seed, trial_offset=offset)
    return steps, time.perf_counter() - t0, tallies


def cpu_port_throughput(iters_per_core: int, cores: int) -> dict:
    """Reference trial loop (Pd_plotter.py:210-223) via the Python port on ``cores`` processes."""
    import multiprocessing as mp
    jobs = [(GEN1, GEN2, N_BLOCK, 0.1, iters_per_core, SEED, i * iters_per_core) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_port_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_port_worker, jobs)
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    return {"value": steps / wall, "unit": "trellis-steps/s", "cores": cores, "kind": "port",
            "sample": f"(7,5)/(6,5) N={N_BLOCK} p=0.1, {iters_per_core} iterations x 2 hypotheses per core "
                      f"(+ one 6200-step learning chain per core), oracle/ref_port.py (pure Python, reference data "
                      f"structures), {wall:.1f} s wall",
            "steps": steps, "wall_s": wall}


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
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: every step ~ 8 s of wall on all cores
    per_core = max(4, min(args.ref_iters, (args.ref_iters * 10) // max(1, args.steps)))   # whole run stays ~ 2 min
    vals = []
    for _ in range(args.warmup and 1):
        cpu_port_throughput(max(2, per_core // 8), cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_port_throughput(per_core, cores))
    wall = time.perf_counter() - t0
    steps = sum(v["steps"] for v in vals)
    value = steps / sum(v["wall_s"] for v in vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "trellis-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": workload_config(args, per_gpu_trials=None),
            "cpu_baseline": {"value": value, "unit": "trellis-steps/s", "cores": cores, "kind": "port",
                             "sample": vals[-1]["sample"]},
            "e2e": {"value": value, "unit": "trellis-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def workload_config(args, per_gpu_trials):
    return {"workload": "Pd_plotter.py paper sweep (BASELINE configs[1]): (7,5) vs (6,5), k=1 n=2 m=2, S=31 Markov states, "
                        f"N={N_BLOCK}, p_vec={P_VEC}, both hypotheses, learn_len=6200 burn=200 laplace=1",
            "trials_per_point_per_gpu": per_gpu_trials, "engine": args.engine,
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


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default="acs", choices=["acs", "fsm"])
    ap.add_argument("--trials", type=int, default=1_000_000, help="Monte-Carlo trials per (N,p) point per GPU")
    ap.add_argument("--ref-iters", type=int, default=1000, help="reference arm: iterations per core per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the alternate-engine / bitstream legs")
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
                                timeout=datetime.timedelta(seconds=120))

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
    det.set_models(tables)
    segs = []
    for q, p in enumerate(P_VEC):
        T = bitsource.bsc_threshold(p)
        segs.append(Seg(N=N_BLOCK, threshold=T, stream=2 * q, table=q, enc_taps=t1, decide=0, trial_begin=begin, trial_end=end))
        segs.append(Seg(N=N_BLOCK, threshold=T, stream=2 * q + 1, table=q, enc_taps=t2, decide=1, trial_begin=begin, trial_end=end))
    d_tallies = torch.zeros(len(segs), dtype=torch.int64, device="cuda")

    def device_pass(engine):
        d_tallies.zero_()
        torch.cuda.synchronize()                 # d_tallies is zeroed on torch's stream, used on libmvd's
        det.detect(segs, seed=SEED, engine=engine, d_tallies_ptr=d_tallies.data_ptr())
        if world > 1:
            dist.all_reduce(d_tallies, op=dist.ReduceOp.SUM)     # the one data-path collective
        return det.last_kernel_ms()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        l0 = det.launch_count()
        kms = []
        t0 = time.perf_counter()
        for _ in range(steps):
            kms.append(fn())
        sync_all()
        dt = max_over_ranks(time.perf_counter() - t0)
        return dt, kms, det.launch_count() - l0

    with ClockSampler(local_rank) as clk:
        dt, kms, launches = timed(lambda: device_pass(args.engine), args.steps, args.warmup)
    clocks = clk.summary()
    value = steps_per_pass * args.steps / dt
    kernel_ms = float(np.mean(kms))
    final_tallies = d_tallies.cpu().numpy().copy()

    # ---- e2e through the public API (host data in, DataFrame out), every step
    def api_pass():
        det_details = {}
        df = pdp.run_experiment(K, NOUT, M, GEN1, GEN2, trials * world, P_VEC, None, 200, 1.0, SEED,
                                engine=args.engine, device=local_rank, details=det_details)
        api_pass.df, api_pass.details = df, det_details
        return det_details["detect_kernel_ms"]

    dt_e2e, _, launches_e2e = timed(api_pass, args.steps, args.warmup)
    e2e_value = steps_per_pass * args.steps / dt_e2e
    S, R = det.S, det.R
    h2d = len(P_VEC) * S * R * 16 + 2 * len(segs) * 96 + len(P_VEC) * 96      # log tables + segment descriptors
    d2h = len(segs) * 8 + len(P_VEC) * S * R * 8 + 8                          # tallies + edge counts + flags
    # the API path must give the same tallies as the resident path (same seeds, same trial ids)
    api_t = api_pass.details["tallies"]
    same = bool(np.array_equal(np.asarray(api_t, dtype=np.int64), final_tallies.astype(np.int64)))

    line = {"metric": METRIC, "value": value, "unit": "trellis-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16x2 metrics + f64 log-likelihood", "data": "synthetic",
            "config": workload_config(args, trials),
            "e2e": {"value": e2e_value, "unit": "trellis-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * dt_e2e / args.steps, "api": "Pd_plotter.run_experiment",
                    "tallies_equal_resident_path": same},
            "gpu_launches": launches, "gpu_launches_e2e": launches_e2e, "clocks": clocks,
            "kernel_ms_per_step": kernel_ms,
            "pd_pc": [{"p": p, "Pd": int(final_tallies[2 * q]) / (trials * world),
                       "Pc": (int(final_tallies[2 * q]) + int(final_tallies[2 * q + 1])) / (2 * trials * world)}
                      for q, p in enumerate(P_VEC)]}

    if rank == 0:
        # ---- roofline: integer pipe (binding) measured on this device, this run
        alu_gops, mixed_gops = det.int_peak()
        rate_kernel = steps_per_pass_rank / (kernel_ms * 1e-3)             # this GPU's kernel-only rate
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(f"{args.engine}_detect")
        except Exception:
            pass
        ach = OPS_CORE * rate_kernel * 1e-9
        line["roofline"] = {
            "bound": "int_alu", "unit": "Gop/s",
            "achieved": ach, "peak": mixed_gops, "frac": ach / mixed_gops,
            "peak_alu_pipe_only": alu_gops, "frac_alu_pipe_only": ach / alu_gops,
            "ops_per_step": {"core": OPS_CORE, "f64_adds": 2,
                             "note": "SURVEY 8(d): 2^(m+k+1) + 2^m + 3n + 5 at m=2, k=1, n=2; bit generation is not "
                                     "counted (ncu: 33.0 warp-instructions issued per trellis step, ~15 of them Philox + "
                                     "lazy Bernoulli + encode, profiles/r01n_*)"},
            "peak_source": "libmvd mvd_int_peak(), measured in this run on this GPU: `peak` = alternating LOP3/IMAD chains "
                           "(ALU + FMA pipes = warp-instruction issue rate, the most any integer code can retire); "
                           "`peak_alu_pipe_only` = LOP3-only chains (min/shift/logic/permute can only issue there). "
                           "MEASURED_PEAKS.json has no integer figure; the path moves ~0 HBM bytes (bits are generated in "
                           "registers), so the HBM roofline does not bind it -- see roofline_hbm_bitstream for the HBM view",
            "kernel": ("detect2p_kernel (two trials/thread ACS)" if args.engine == "acs" else "detect2_kernel<FSM1> (one-load NEXT walk)"),
            "kernel_ms": kernel_ms, "steps_per_launch": steps_per_pass_rank,
            "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/)",
            "hbm_bytes_per_step_algorithmic": 0.0,
        }
        if not args.no_extras and world == 1:       # single-GPU legs only: nothing below may enter a collective
            other = "fsm" if args.engine == "acs" else "acs"
            dto, kmo, _ = timed(lambda: device_pass(other), max(3, args.steps // 2), 3)
            line["alt_engine"] = {"engine": other, "value": steps_per_pass * max(3, args.steps // 2) / dto,
                                  "kernel_ms_per_step": float(np.mean(kmo)),
                                  "tallies_equal": bool(np.array_equal(d_tallies.cpu().numpy(), final_tallies))}
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
            base = cpu_port_throughput(args.ref_iters, cores)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            one = cpu_port_throughput(max(20, args.ref_iters // 3), 1)      # as shipped: the reference is single-process
            line["cpu_baseline_1core"] = {k: one[k] for k in ("value", "unit", "cores", "kind", "sample")}
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
