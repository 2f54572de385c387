#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_full_size.py -m gpu -q -x --timeout 600 2>&1 | tail -30 > gpurun_out/r2d_tests.log
tail -4 gpurun_out/r2d_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-by-config --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json')); print(d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'])"
KSTEP=5,12,18,19 timeout 300 python tools/trace_fwd16.py 0 > gpurun_out/r2d_trace_cta0.txt 2>&1
