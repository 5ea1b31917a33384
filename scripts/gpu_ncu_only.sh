#!/bin/bash
# ncu --set full of one kernel of the bench command. usage: scripts/gpu_ncu_only.sh <tag> <engine> <kernel regex> [skip]
set -u
TAG=$1; ENGINE=$2; KREGEX=$3; SKIP=${4:-4}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --engine $ENGINE --steps 3 --warmup 3 --trials 200000 --no-extras --no-cpu-baseline"
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${SKIP} -c 1 -f -o $OUT/prof_${ENGINE}_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
