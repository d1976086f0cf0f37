#!/bin/bash
# round 2, call 10: chunk-size sweep (small chunks) on one box for the full pool and the 8-GPU shard; bench.py default run
mkdir -p gpurun_out
timeout 900 python scripts/bench_search_modes.py 190000 5600000 3 "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=24" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=32" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=48" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=64" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=96" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=128" 2>&1 | tee gpurun_out/r2c10_modes_full.txt
timeout 600 python scripts/bench_search_modes.py 190000 700000 7 "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=32" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=64" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=96" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=128" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=256" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=512" 2>&1 | tee gpurun_out/r2c10_modes_shard.txt
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2c10_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c10_bench.json'))
def show(tag,o):
    r=o.get('roofline',{})
    print(tag, round(o['value'],1), o['unit'], 'e2e', round(o['e2e']['value'],1), 'frac', round(r.get('frac',0),3), 'iso', round(r.get('isolated_frac',0) or 0,3), 'step_alg', round(r.get('frac_step_algorithmic',0) or 0,3), 'cpu', o.get('cpu_baseline',{}).get('value'), o.get('clocks'))
show('train',d)
for k,v in d.get('extra',{}).items(): show(k,v)
PY
