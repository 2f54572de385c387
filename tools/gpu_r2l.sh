#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r2l_bench.json')); print(round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['phases_ms'].items()})
for k,v in d['by_config'].items(): print(k, round(v['ms_per_step'],3), {a:round(b,3) for a,b in v['phases_ms'].items()})"
