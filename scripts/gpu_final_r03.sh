#!/bin/bash
# Final one-GPU pass of round 2: parity tests, smoke, bench (all legs), reference arm, other configs, paper sweeps,
# launch list of the bench command, full ncu captures of the headline kernel and of the split kernels.
set -u
TAG=${1:-r03z}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
( time python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -3 $OUT/bench_$TAG.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
python scripts/gpu_configs.py m2 m3 m4 m56 m4full > $OUT/configs_$TAG.jsonl 2> $OUT/configs_$TAG.err; echo "configs rc=$?"
python scripts/micro/split_config3.py > $OUT/split_config3_$TAG.jsonl 2> $OUT/split_config3_$TAG.err; echo "split rc=$?"
python scripts/micro/rate13.py > $OUT/rate13_$TAG.jsonl 2>&1; echo "rate13 rc=$?"
python scripts/paper_sweeps.py $OUT/ps_$TAG 1000000 > $OUT/ps_$TAG.log 2>&1; echo "sweeps rc=$?"; grep "^{" $OUT/ps_$TAG.log | cut -c1-200
CMD="python bench.py --steps 3 --warmup 3 --trials 200000 --no-extras --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:detect2p_kernel -s 4 -c 1 -f -o $OUT/prof_acs_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/split_launches_$TAG.csv python scripts/micro/split_one.py > $OUT/split_one_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"split_(score|isum|walk2|plan)_kernel" -s 4 -c 4 -f -o $OUT/prof_split_$TAG python scripts/micro/split_one.py > $OUT/ncu_split_$TAG.log 2>&1
echo "ncu split rc=$?"
head -c 600 $OUT/bench_$TAG.json
