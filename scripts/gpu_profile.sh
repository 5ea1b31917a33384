#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, bench, ncu launch list and one full capture.
# usage: scripts/gpu_profile.sh <tag> [engine] [kernel regex] [launches to skip]
set -u
TAG=${1:-r01}
ENGINE=${2:-acs}
KREGEX=${3:-detect2_kernel}
SKIP=${4:-4}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/pytest_gpu_$TAG.log
python bench.py --engine $ENGINE > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
CMD="python bench.py --engine $ENGINE --steps 3 --warmup 3 --trials 200000 --no-extras --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${SKIP} -c 1 -f -o $OUT/prof_${ENGINE}_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
cat $OUT/bench_$TAG.json
