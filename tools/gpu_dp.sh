#!/bin/bash
# bash tools/gpu_dp.sh N  -- data-parallel bench on N GPUs: one all-reduce after the backward vs the three overlapped buckets
N=${1:-2}
mkdir -p gpurun_out
for ov in 0 1; do
  echo "== ICL_AR_OVERLAP=$ov N=$N"
  ICL_AR_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --no-by-config > gpurun_out/r2_dp_n${N}_ov$ov.json 2> gpurun_out/r2_dp_n${N}_ov$ov.err
  echo "rc $?"; tail -2 gpurun_out/r2_dp_n${N}_ov$ov.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_dp_n${N}_ov$ov.json')); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],3), {k:round(x,3) for k,x in d['phases_ms'].items()})"
done
