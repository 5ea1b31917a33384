#!/usr/bin/env python3
"""Where the per-sweep fixed cost of run_experiment goes (host timers around each stage, models cached).
usage: [torchrun ...] scripts/micro/profile_sweep_overhead.py [trials]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
import numpy as np
import Pd_plotter as pdp
from mvd import dist
G1, G2 = [[[1, 1, 1]], [[1, 0, 1]]], [[[1, 1, 0]], [[1, 0, 1]]]
P7 = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rank, ws = dist.world()
if ws > 1:
    import torch
    import torch.distributed as td
for Ns in ([500], [50, 100, 200, 500, 1000]):
    walls, stages = [], []
    for it in range(30):
        if ws > 1:
            torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
        d = {}
        t0 = time.perf_counter()
        pdp.run_experiment(1, 2, 2, G1, G2, trials, P7, None, 200, 1.0, 12345, N_spectrum=Ns, details=d, engine="acs")
        walls.append(time.perf_counter() - t0)
        stages.append(dict(d["wall_s"], kernel=d["detect_kernel_ms"] * 1e-3, **d.get("stage_s", {})))
    if rank == 0:
        med = lambda xs: float(np.median(xs[5:]))
        print(json.dumps(dict(n_gpus=ws, points=len(Ns) * 7, trials=trials, wall_us=round(1e6 * med(walls), 1),
                              **{k + "_us": round(1e6 * med([s[k] for s in stages]), 1) for k in stages[0]})), flush=True)
if ws > 1:
    td.destroy_process_group()
