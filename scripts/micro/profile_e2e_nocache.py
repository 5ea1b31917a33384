import sys, time, cProfile, pstats, io
sys.path.insert(0, "detecting-convolutional-codes-via-markovian-statistics_b200")
import Pd_plotter as pdp
P7 = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
args = (1, 2, 2, [[[1,1,1]],[[1,0,1]]], [[[1,1,0]],[[1,0,1]]], 1000000, P7, None, 200, 1.0, 12345)
KW = dict(engine="acs", cache_models=False)
for _ in range(3): pdp.run_experiment(*args, **KW)
for _ in range(3):
    d = {}
    t0 = time.perf_counter(); pdp.run_experiment(*args, details=d, **KW); w = time.perf_counter() - t0
    print("wall ms", round(1e3 * w, 3), {k: round(1e3 * v, 3) for k, v in d["wall_s"].items()}, {k: round(1e3 * v, 3) for k, v in d["stage_s"].items()}, "detect kernel ms", round(d["detect_kernel_ms"], 3), "learn kernel ms", round(d["learn_kernel_ms"], 3))
pr = cProfile.Profile(); pr.enable()
for _ in range(20): pdp.run_experiment(*args, **KW)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30); print(s.getvalue()[:6000])
