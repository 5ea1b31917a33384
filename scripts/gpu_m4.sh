#!/bin/bash
# m = 4 pair kernel: parity tests of the m = 4 paths, throughput (three repeats), ncu summary
TAG=${1:-r03g}
OUT=gpurun_out; mkdir -p $OUT
python scripts/micro/dbg_m3.py m4 2>&1 | grep -v "bad=0" | tail -20
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "m4 or large or hash or global" > $OUT/pytest_$TAG.log 2>&1; tail -3 $OUT/pytest_$TAG.log
for i in 1 2 3; do python scripts/gpu_configs.py m4 2>> $OUT/configs_$TAG.err | tee -a $OUT/configs_$TAG.jsonl | cut -c1-420; done
ncu --set full --clock-control none --import-source on -k regex:detect3p_kernel -s 2 -c 1 -f -o $OUT/prof_m4_$TAG python scripts/gpu_configs.py m4 > $OUT/ncu_m4_$TAG.log 2>&1
echo "ncu rc=$?"
