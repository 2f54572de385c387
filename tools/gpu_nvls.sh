#!/bin/bash
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/check_nvls.py 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL version" | tail -6
for nv in 0 1; do
  echo "== ICL_AR_NVLS=$nv N=$N"
  ICL_AR_NVLS=$nv timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --no-by-config > gpurun_out/r2_nvls_n${N}_$nv.json 2> gpurun_out/r2_nvls_n${N}_$nv.err
  echo "rc $?"; tail -2 gpurun_out/r2_nvls_n${N}_$nv.err | cut -c1-300
  python -c "
import json; d=json.load(open('gpurun_out/r2_nvls_n${N}_$nv.json')); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), 'b2b', round(d['config']['back_to_back_ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), 'phases sum', round(sum(d['phases_ms'].values()),3))"
done
