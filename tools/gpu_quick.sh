#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-by-config > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench.json')); print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), d['gpu_launches'], {k:round(x,4) for k,x in d['phases_ms'].items()})"
