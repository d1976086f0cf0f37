#!/bin/bash
# round 2, call 25 (1 GPU): GPU suite + the default bench line at N = 1 after restoring the GEMM parameter struct size
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -40 > gpurun_out/r2c25_tests.log
tail -3 gpurun_out/r2c25_tests.log
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/r2c25_bench_1gpu.json 2> gpurun_out/r2c25_bench_1gpu.err; echo "bench1 exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c25_bench_1gpu.json').read().strip().splitlines()[-1])
def show(tag,o):
    r=o.get('roofline',{})
    print(tag, round(o['value'],1), o['unit'], 'e2e', round(o['e2e']['value'],1), 'frac', round(r.get('frac',0),3), 'alg', r.get('frac_step_algorithmic'), o.get('clocks'))
show('train',d)
for k,v in d.get('extra',{}).items(): show(k,v)
PY
