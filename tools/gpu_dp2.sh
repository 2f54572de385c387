#!/bin/bash
# bash tools/gpu_dp2.sh N tag -- the driver's multi-GPU bench line on N GPUs (both arms), default settings
N=${1:-2}; tag=${2:-r2w}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${tag}_n${N}.json 2> gpurun_out/${tag}_n${N}.err
echo "rc $?"; tail -3 gpurun_out/${tag}_n${N}.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_n${N}.json')); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],3), {k:round(x,3) for k,x in d['phases_ms'].items()})
for k,v in d.get('by_config',{}).items(): print(k, round(v['ms_per_step'],4), 'e2e', round(v['e2e']['ms_per_step'],3))"
