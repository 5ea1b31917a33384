#!/bin/bash
# m = 3 / m = 4 pair kernel: parity tests, ladder against the NEXT walk, throughput, ncu summary of the m = 3 kernel
TAG=${1:-r03d}
OUT=gpurun_out; mkdir -p $OUT
python scripts/micro/dbg_m3.py m4 2>&1 | grep -v "bad=0" | tail -20
python -m pytest tests/test_gpu_parity.py tests/test_k2_codes.py -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; tail -3 $OUT/pytest_$TAG.log
python scripts/gpu_configs.py m3 m4 > $OUT/configs_$TAG.jsonl 2> $OUT/configs_$TAG.err; cut -c1-400 $OUT/configs_$TAG.jsonl; tail -3 $OUT/configs_$TAG.err
ncu --set full --clock-control none --import-source on -k regex:detect3p_kernel -s 2 -c 1 -f -o $OUT/prof_m3_$TAG python scripts/gpu_configs.py m3 > $OUT/ncu_m3_$TAG.log 2>&1
echo "ncu rc=$?"
