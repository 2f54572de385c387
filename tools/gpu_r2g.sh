#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_full_size.py -m gpu -q -x --timeout 600 2>&1 | tail -40 > gpurun_out/r2g_tests.log
tail -6 gpurun_out/r2g_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r2g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json')); print(d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'])
for k,v in d['by_config'].items(): print(k, round(v['ms_per_step'],3), {a:round(b,3) for a,b in v['phases_ms'].items()})"
NEV=70 python tools/trace_bptt.py 0 > gpurun_out/r2g_trace_bptt.txt 2>&1
