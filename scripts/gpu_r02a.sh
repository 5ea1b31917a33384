#!/bin/bash
# Round-2 first GPU pass: parity tests, smoke, bench (all legs), ncu of the m = 4 pair kernel, launch list of the bench.
set -u
TAG=${1:-r02a}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
( time python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -5 $OUT/bench_$TAG.err
python scripts/gpu_configs.py m3 m4 > $OUT/configs_$TAG.jsonl 2> $OUT/configs_$TAG.err; echo "configs rc=$?"; cat $OUT/configs_$TAG.jsonl
ncu --set full --clock-control none --import-source on -k regex:detect3p_kernel -s 2 -c 1 -f -o $OUT/prof_m4_$TAG python scripts/gpu_configs.py m4 > $OUT/ncu_m4_$TAG.log 2>&1
echo "ncu m4 rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --trials 200000 --no-extras --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
head -c 3000 $OUT/bench_$TAG.json
