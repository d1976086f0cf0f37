#!/bin/bash
# round 2, call 23 (1 GPU): GPU suite with the CLIP_FF dropout test; BLIP_FF configs[2] step with / without the reference's dropout
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -40 > gpurun_out/r2c23_tests.log
tail -5 gpurun_out/r2c23_tests.log
timeout 600 python bench.py --workload blip --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c23_blip_dropout.json 2> gpurun_out/r2c23_blip_dropout.err; echo "blip(dropout) exit $?"
timeout 600 python bench.py --workload blip --steps 2 --warmup 3 --no-cpu-baseline --blip-no-dropout > gpurun_out/r2c23_blip_nodropout.json 2> gpurun_out/r2c23_blip_nodropout.err; echo "blip(no dropout) exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/r2c23_blip_dropout.json','gpurun_out/r2c23_blip_nodropout.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],1), 'frac', round(r['frac'],3), 'alg', round(r['frac_step_algorithmic'],3), 'gemm share', round(r['gemm_share_of_step'],3), 'launches', d['gpu_launches'], 'loss', d['loss'], d['clocks'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
