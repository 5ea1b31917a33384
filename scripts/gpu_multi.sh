#!/bin/bash
# multi-GPU bench: usage scripts/gpu_multi.sh <ngpus> <tag>
set -u
N=$1; TAG=$2; OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "rc=$?"; tail -3 $OUT/bench_${TAG}_n$N.err; cut -c1-600 $OUT/bench_${TAG}_n$N.json
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > $OUT/bench_ref_${TAG}_n$N.json 2> $OUT/bench_ref_${TAG}_n$N.err
echo "ref rc=$?"; cut -c1-400 $OUT/bench_ref_${TAG}_n$N.json
