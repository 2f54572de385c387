#!/bin/bash
# full GPU regression: every GPU test, then the default bench line (with by_config and the CPU baseline)
tag=${1:-r2y}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
tail -4 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc $?"
tail -c 400 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), d['gpu_launches'], {k:round(x,4) for k,x in d['phases_ms'].items()})
for k,v in d.get('by_config',{}).items(): print(k, round(v['ms_per_step'],4), 'e2e', round(v['e2e']['ms_per_step'],3), {a:round(x,3) for a,x in v['phases_ms'].items()})
print(d.get('mention_box_pairs_per_sec'))
PY
