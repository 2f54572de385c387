#!/bin/bash
# round-2 GPU call: all GPU tests, the bench line, ncu launch list + full capture (with source) of the two recurrences
tag=${1:-r2z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc $?"
tail -c 300 gpurun_out/${tag}_bench.err
Q="python bench.py --steps 2 --warmup 3 --no-by-config --no-cpu-baseline"
$Q > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv $Q > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu launches rc $?"
$Q > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_rec_fwd16|k_bptt_cluster' -s 8 -c 4 -f -o gpurun_out/${tag}_rec $Q > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu full rc $?"
ls -la gpurun_out | grep ${tag}
