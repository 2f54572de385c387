#!/bin/bash
# one compute-sanitizer tool per gpurun call:  bash tools/gpu_san.sh racecheck|synccheck|memcheck [S H T]
tool=$1; S=${2:-256}; H=${3:-300}; T=${4:-20}
mkdir -p gpurun_out
ICL_NO_TORCH_STREAM=1 timeout 1500 python tools/sanitize_run.py $S $H $T > gpurun_out/r2_san_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_san_plain.log; exit 1; }
ICL_NO_TORCH_STREAM=1 timeout 1700 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py $S $H $T > gpurun_out/r2_san_$tool.log 2>&1
echo "sanitizer rc $?"
tail -15 gpurun_out/r2_san_$tool.log
