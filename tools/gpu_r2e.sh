#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_full_size.py -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/r2e_tests$i.log
tail -2 gpurun_out/r2e_tests$i.log
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r2e_bench.json')); print(d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'])
for k,v in d['by_config'].items(): print(k, round(v['ms_per_step'],3), {a:round(b,3) for a,b in v['phases_ms'].items()})"
python tools/trace_bptt.py 0 > gpurun_out/r2e_trace_bptt.txt 2>&1
