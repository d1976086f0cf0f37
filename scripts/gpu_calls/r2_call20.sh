#!/bin/bash
# round 2, call 20 (2 GPUs): default bench line at N = 2 and N = 1 with the final tree; micro-batch size experiment
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2c20_bench_2gpu.json 2> gpurun_out/r2c20_bench_2gpu.err; echo "bench2 exit $?"
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2c20_bench_1gpu.json 2> gpurun_out/r2c20_bench_1gpu.err; echo "bench1 exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/r2c20_bench_2gpu.json','gpurun_out/r2c20_bench_1gpu.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    def show(tag,o):
        r=o.get('roofline',{})
        print(tag, round(o['value'],1), o['unit'], 'e2e', round(o['e2e']['value'],1), 'frac', round(r.get('frac',0),3), 'iso', r.get('isolated_frac'), 'alg', r.get('frac_step_algorithmic'), o.get('clocks'))
    show(f+' train',d)
    for k,v in d.get('extra',{}).items(): show(k,v)
PY
for mb in 64 256; do
  CUDA_VISIBLE_DEVICES=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --micro-batch $mb 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('micro-batch', $mb, round(d['value'],1), d['config']['grad_cache'])"
done
