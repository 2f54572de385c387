#!/bin/bash
# round-2 GPU call A: all GPU tests, the bench line, sanitizer runs on a reduced workload
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc $?"
tail -c 600 gpurun_out/r2a_bench.err
