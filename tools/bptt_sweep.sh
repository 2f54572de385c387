#!/bin/bash
# bring-up: rec_bwd phase time of the card2048 bench under the k_bptt_step timing experiments
for cs in ${CSS:-4 2 1}; do for dbg in ${DBGS:-0 1 2 3 4 6 7}; do
  echo -n "cs=$cs dbg=$dbg: "
  ICL_BPTT_CS=$cs ICL_BPTT_DBG=$dbg timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['phases_ms']['rec_bwd'], d['ms_per_step'])"
done; done
