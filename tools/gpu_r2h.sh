#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2h_tests.log
tail -6 gpurun_out/r2h_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-by-config > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r2h_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print(d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'])"
KSTEP=5,12,18,19 timeout 300 python tools/trace_fwd16.py 0 > gpurun_out/r2h_trace_cta0.txt 2>&1
