#!/bin/bash
# quick GPU check: parity tests + short bench for both engines
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-q}
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu_$TAG.log
python bench.py --engine acs --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -3 $OUT/bench_$TAG.err
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value %.4g e2e %.4g kernel_ms %.3f frac %.3f alt(%s) %.4g" % (d["value"], d["e2e"]["value"], d["kernel_ms_per_step"], d["roofline"]["frac"], d["alt_engine"]["engine"], d["alt_engine"]["value"]))
print("tallies equal:", d["e2e"]["tallies_equal_resident_path"], d["alt_engine"]["tallies_equal"])
PY
