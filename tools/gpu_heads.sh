#!/bin/bash
# multi-head concurrency A/B: the multitask512 workload with the heads one after another / side by side
mkdir -p gpurun_out
for hs in 0 1; do
  echo "ICL_HEAD_STREAMS=$hs"; ICL_HEAD_STREAMS=$hs timeout 300 python tools/bench_configs.py multitask512 2>&1 | tail -2
done
