#!/bin/bash
# round-2 GPU call C: K2 with store warps -- parity tests, short bench, trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -30 > gpurun_out/r2c_tests.log
tail -4 gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-by-config --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench.json')); print(d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'])"
KSTEP=5,12,18,19 timeout 300 python tools/trace_fwd16.py 0 > gpurun_out/r2c_trace_cta0.txt 2>&1
