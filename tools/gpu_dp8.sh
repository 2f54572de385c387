#!/bin/bash
# 8-GPU end-to-end check of the packing-thread tuner: default (measured per rank) vs fixed 4 threads per rank
mkdir -p gpurun_out
nproc
for t in auto 4; do
  if [ $t = auto ]; then unset ICL_HOST_THREADS; else export ICL_HOST_THREADS=$t; fi
  ICL_HOST_THREADS_VERBOSE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 15 --warmup 3 --no-by-config --no-cpu-baseline > gpurun_out/r2v_n8_$t.json 2> gpurun_out/r2v_n8_$t.err
  echo "threads=$t rc $?"; grep "host packing threads" gpurun_out/r2v_n8_$t.err | head -3
  python -c "
import json; d=json.load(open('gpurun_out/r2v_n8_$t.json')); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e captions/s', round(d['e2e']['value']), 'corpus cache ms', round(d['e2e_variants']['corpus_cache']['ms_per_step'],3), d['e2e_host_limit']['host_copy_bandwidth'])"
done
