#!/bin/bash
# round-2 GPU call B: the three tests that failed in call A (full output) + clock64 trace of k_rec_fwd16 at early / late steps
mkdir -p gpurun_out
python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q --timeout 900 -k "nonvis512 or resident_token or resident_box" 2>&1 | tail -80 > gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
KSTEP=5,12,13,18,19 python tools/trace_fwd16.py 0 > gpurun_out/r2b_trace_cta0.txt 2>&1
KSTEP=5,18 python tools/trace_fwd16.py 47 > gpurun_out/r2b_trace_cta47.txt 2>&1
tail -3 gpurun_out/r2b_trace_cta0.txt
