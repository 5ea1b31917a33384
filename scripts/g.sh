#!/bin/bash
# Local helper: rebuild libmvd.so (a stale .so travels to the GPU box as it is), then gpurun the given command.
# usage: scripts/g.sh <timeout_s> '<command>'   (add GPUS=N in the environment for --gpus N)
set -e
cd "$(dirname "$0")/.."
python detecting-convolutional-codes-via-markovian-statistics_b200/build.py > /dev/null
python -c "import sys; sys.path.insert(0,'oracle'); import c_oracle, fetch_ref; c_oracle.build(); fetch_ref.fetch()"
T=$1; shift
if [ -n "${GPUS:-}" ]; then exec gpurun --gpus $GPUS --timeout $T -- "$@"; else exec gpurun --timeout $T -- "$@"; fi
