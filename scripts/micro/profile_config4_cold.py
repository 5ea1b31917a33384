"""BASELINE config 4 at the reference's defaults through run_experiment with nothing cached (cache_models=False):
where the wall time of a first call goes (host tables vs kernels).  usage: python scripts/micro/profile_config4_cold.py [S25751|S150743]"""
import cProfile, io, json, pstats, sys, time
sys.path.insert(0, "detecting-convolutional-codes-via-markovian-statistics_b200")
import Pd_plotter as pdp
import viterbi_markov as vm
from mvd import codes
P7 = [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5]
which = sys.argv[1] if len(sys.argv) > 1 else "S150743"
g1, g2 = {"S25751": ([[[1,1,0,0,1]],[[1,1,0,1,1]]], [[[1,1,0,1,1]],[[1,1,0,0,1]]]),
          "S150743": ([[[1,0,0,1,1]],[[1,1,1,0,1]]], [[[1,1,1,0,1]],[[1,0,0,1,1]]])}[which]
t0 = time.perf_counter()
det = vm._detector(codes.freeze_generator(g1), 1, 2, 4)
t_det = time.perf_counter() - t0
args = (1, 2, 4, g1, g2, 100000, P7, None, 200, 1.0, 12345)
pdp.run_experiment(*args, cache_models=False)
walls = []
for _ in range(3):
    d = {}
    t0 = time.perf_counter()
    pdp.run_experiment(*args, details=d, cache_models=False)
    walls.append(time.perf_counter() - t0)
print(json.dumps(dict(case=f"config 4 {which} reference defaults, nothing cached", S=d["S"], detector_s=round(t_det, 4), wall_s=[round(w, 4) for w in walls],
                      breakdown_s={k: round(v, 4) for k, v in d["wall_s"].items()}, detect_kernel_ms=round(d["detect_kernel_ms"], 3),
                      learn_kernel_ms=round(d["learn_kernel_ms"], 3), learn_len=d["learn_len"])), flush=True)
pr = cProfile.Profile(); pr.enable()
pdp.run_experiment(*args, cache_models=False)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:4500])
