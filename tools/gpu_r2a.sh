#!/bin/bash
# round-2 GPU call A: all GPU tests, the bench line, ncu launch list + full capture of the two recurrences
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc $?"
tail -c 600 gpurun_out/r2a_bench.err
Q="python bench.py --steps 2 --warmup 3 --no-by-config --no-cpu-baseline"
$Q > gpurun_out/r2a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv $Q > gpurun_out/r2a_ncu1.log 2>&1
echo "ncu launches rc $?"
$Q > gpurun_out/r2a_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_rec_fwd16|k_bptt_cluster' -s 8 -c 4 -f -o gpurun_out/r2a_rec $Q > gpurun_out/r2a_ncu2.log 2>&1
echo "ncu full rc $?"
ls -la gpurun_out
