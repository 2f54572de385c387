#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -30 > gpurun_out/r2j_tests.log
tail -4 gpurun_out/r2j_tests.log
for u in 20 16; do
  echo "== ICL_RF_U=$u"
  ICL_RF_U=$u timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-by-config > gpurun_out/r2j_bench_u$u.json 2> gpurun_out/r2j_bench_u$u.err
  python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_u$u.json')); print(round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['phases_ms'].items()})"
done
ICL_TRACE_CS=0 KSTEP=5,12,18 timeout 300 python tools/trace_fwd16.py 0 > gpurun_out/r2j_trace_u16.txt 2>&1
