#!/bin/bash
# bash tools/gpu_scale.sh N  -- the driver's own invocation of both arms on N GPUs
N=${1:-8}
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2z_ref_n$N.json 2> gpurun_out/r2z_ref_n$N.err; echo "reference arm rc $? ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2z_bench_n$N.json 2> gpurun_out/r2z_bench_n$N.err; echo "b200 arm rc $? ($(( $(date +%s) - t0 )) s)"
tail -2 gpurun_out/r2z_bench_n$N.err
python -c "
import json
r=json.load(open('gpurun_out/r2z_ref_n$N.json')); print('reference', round(r['value']), r['cpu_baseline']['cores'], r['config']['batch_per_step'])
d=json.load(open('gpurun_out/r2z_bench_n$N.json')); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],3), round(d['e2e']['value']), 'b2b', round(d['config']['back_to_back_ms_per_step'],4))
print(d.get('e2e_host_limit'))
for k,v in d.get('by_config',{}).items(): print(k, round(v['value']), round(v['ms_per_step'],3), 'e2e', round(v['e2e']['ms_per_step'],3))
for k,v in (d.get('e2e_variants') or {}).items(): print(k, round(v['ms_per_step'],3))"
