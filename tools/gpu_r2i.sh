#!/bin/bash
mkdir -p gpurun_out
for v in 2 3; do
  echo "== ICL_RF_VARIANT=$v"
  ICL_RF_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-by-config > gpurun_out/r2i_bench_v$v.json 2> gpurun_out/r2i_bench_v$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_v$v.json')); print(round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['phases_ms'].items()})"
done
ICL_TRACE_CS=0 ICL_RF_VARIANT=3 KSTEP=5,12,18 timeout 300 python tools/trace_fwd16.py 0 > gpurun_out/r2i_trace_v3.txt 2>&1
ICL_RF_VARIANT=3 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -3
