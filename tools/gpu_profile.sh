#!/bin/bash
# final-code profiles: ncu launch list of two bench steps (card2048), then ncu --set full of the recurrences and of the factorised
# affinity layer's kernels (affinity512).  Each ncu pass only after the same command has exited 0 without ncu.
tag=${1:-r2f}
mkdir -p gpurun_out
Q="python bench.py --steps 2 --warmup 3 --no-by-config --no-cpu-baseline"
$Q > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv $Q > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu launches rc $?"
ncu --set full --clock-control none --import-source on -k 'regex:k_rec_fwd16|k_bptt_cluster' -s 8 -c 4 -f -o gpurun_out/${tag}_rec $Q > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu full rc $?"
A="python tools/bench_configs.py affinity512 --steps 2"
$A > gpurun_out/${tag}_aff_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_aff_launches.csv $A > gpurun_out/${tag}_ncu3.log 2>&1
echo "ncu affinity launches rc $?"
ls -la gpurun_out | grep ${tag}
