#!/bin/bash
# multi-GPU evidence of the round: usage scripts/gpu_multi_r02.sh <ngpus> <tag>
# bench line with the strong-scaling / paper-sweep / sustained legs, one config-5 rung (1e7 trials per point per GPU), the
# paper sweeps at 1e6 trials per point and at the reference's own 1e4 (BASELINE config 3: few long trials per GPU)
set -u
N=$1; TAG=$2; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "bench rc=$?"; tail -2 $OUT/bench_${TAG}_n$N.err; cut -c1-300 $OUT/bench_${TAG}_n$N.json
timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --trials 10000000 --no-extras --no-cpu-baseline > $OUT/config5_${TAG}_n$N.jsonl 2> $OUT/config5_${TAG}_n$N.err
echo "config5 rc=$?"; cut -c1-300 $OUT/config5_${TAG}_n$N.jsonl
timeout 200 $TR --master-port 29513 scripts/paper_sweeps.py $OUT/ps_${TAG}_n${N}_1e6 1000000 > $OUT/ps_${TAG}_n${N}_1e6.jsonl 2> $OUT/ps_${TAG}_n${N}_1e6.err
echo "sweeps 1e6 rc=$?"; cat $OUT/ps_${TAG}_n${N}_1e6.jsonl
timeout 200 $TR --master-port 29514 scripts/paper_sweeps.py $OUT/ps_${TAG}_n${N}_1e4 10000 > $OUT/ps_${TAG}_n${N}_1e4.jsonl 2> $OUT/ps_${TAG}_n${N}_1e4.err
echo "sweeps 1e4 rc=$?"; cat $OUT/ps_${TAG}_n${N}_1e4.jsonl
timeout 200 python scripts/paper_sweeps.py $OUT/ps_${TAG}_n1_1e4 10000 > $OUT/ps_${TAG}_n1_1e4.jsonl 2> $OUT/ps_${TAG}_n1_1e4.err
echo "sweeps 1e4 one GPU rc=$?"; cat $OUT/ps_${TAG}_n1_1e4.jsonl
for s in 1e6 1e4; do :; done
cmp $OUT/ps_${TAG}_n${N}_1e4/Pd_hybrid_m2_7_5_vs_6_5_vs_N.csv $OUT/ps_${TAG}_n1_1e4/Pd_hybrid_m2_7_5_vs_6_5_vs_N.csv && echo "Pd vs N CSV: $N GPUs == 1 GPU"
cmp $OUT/ps_${TAG}_n${N}_1e4/Pd_hybrid_m3_demo_vs_N.csv $OUT/ps_${TAG}_n1_1e4/Pd_hybrid_m3_demo_vs_N.csv && echo "m3 Pd vs N CSV: $N GPUs == 1 GPU"
