#!/bin/bash
# Round-2 GPU pass on one GPU: parity tests, smoke, bench (all legs), reference arm, config-5 ladder,
# ncu of the m = 4 pair kernel and of the headline kernel, launch list of the bench.
set -u
TAG=${1:-r02g}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
( time python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -5 $OUT/bench_$TAG.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; tail -3 $OUT/bench_ref_$TAG.err
: > $OUT/config5_$TAG.jsonl
for T in 1000000 10000000 100000000; do
  python bench.py --trials $T --steps 3 --warmup 3 --no-extras --no-cpu-baseline >> $OUT/config5_$TAG.jsonl 2>> $OUT/config5_$TAG.err; echo "ladder $T rc=$?"
done
python scripts/gpu_configs.py m3 m4 > $OUT/configs_$TAG.jsonl 2> $OUT/configs_$TAG.err; echo "configs rc=$?"
ncu --set full --clock-control none --import-source on -k regex:detect3p_kernel -s 2 -c 1 -f -o $OUT/prof_m4_$TAG python scripts/gpu_configs.py m4 > $OUT/ncu_m4_$TAG.log 2>&1
echo "ncu m4 rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --trials 200000 --no-extras --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
head -c 1500 $OUT/bench_$TAG.json
