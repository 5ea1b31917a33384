import sys, time, cProfile, pstats, io
sys.path.insert(0, "detecting-convolutional-codes-via-markovian-statistics_b200")
import Pd_plotter as pdp
P7 = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
NUM_ITER = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
args = (1, 2, 2, [[[1,1,1]],[[1,0,1]]], [[[1,1,0]],[[1,0,1]]], NUM_ITER, P7, None, 200, 1.0, 12345, )
KW = dict(engine=sys.argv[2]) if len(sys.argv) > 2 else {}
for _ in range(3): pdp.run_experiment(*args, **KW)
d = {}
t0 = time.perf_counter(); pdp.run_experiment(*args, details=d, **KW); print("wall ms", 1e3 * (time.perf_counter() - t0), {k: round(1e3 * v, 3) for k, v in d["wall_s"].items()}, "detect kernel ms", d["detect_kernel_ms"], "learn kernel ms", d["learn_kernel_ms"])
pr = cProfile.Profile(); pr.enable()
for _ in range(20): pdp.run_experiment(*args, **KW)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
