#!/bin/bash
# round 2, call 19 (4 GPUs): the default bench line at N = 4 (the one N not yet exercised) + GPU suite on the final tree
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c19_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2c19_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/r2c19_bench_4gpu.json 2> gpurun_out/r2c19_bench_4gpu.err; echo "bench exit $?"; tail -2 gpurun_out/r2c19_bench_4gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c19_bench_4gpu.json').read().strip().splitlines()[-1])
def show(tag,o):
    r=o.get('roofline',{})
    print(tag, round(o['value'],1), o['unit'], 'e2e', round(o['e2e']['value'],1), 'frac', round(r.get('frac',0),3), 'iso', r.get('isolated_frac'), 'traffic', r.get('traffic'), o.get('clocks'))
show('train',d)
for k,v in d.get('extra',{}).items(): show(k,v)
PY
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 | cut -c1-400
