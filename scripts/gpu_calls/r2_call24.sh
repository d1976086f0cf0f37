#!/bin/bash
# round 2, call 24 (1 GPU): the default bench line at N = 1 with the final tree (extra.blip now in train mode with dropout)
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/r2c24_bench_1gpu.json 2> gpurun_out/r2c24_bench_1gpu.err; echo "bench1 exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c24_bench_1gpu.json').read().strip().splitlines()[-1])
def show(tag,o):
    r=o.get('roofline',{})
    print(tag, round(o['value'],1), o['unit'], 'e2e', round(o['e2e']['value'],1), 'frac', round(r.get('frac',0),3), 'alg', r.get('frac_step_algorithmic'), o.get('clocks'))
show('train',d)
for k,v in d.get('extra',{}).items(): show(k,v)
PY
