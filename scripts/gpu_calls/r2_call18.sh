#!/bin/bash
# round 2, call 18 (8 GPUs): BASELINE configs[3] for real -- 5.6 M synthetic items row-sharded over 8 GPUs, per-rank shard
# files, sharded top-50 search of 190 k queries over those shards -- then the CLI under torch.distributed.run, then the
# default bench line at N = 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 scripts/pipeline_embed_search.py --items 5600000 --queries 190000 > gpurun_out/r2c18_pipeline_8gpu.json 2> gpurun_out/r2c18_pipeline_8gpu.err; echo "pipeline exit $?"; cat gpurun_out/r2c18_pipeline_8gpu.json; tail -3 gpurun_out/r2c18_pipeline_8gpu.err
timeout 600 python scripts/cli_e2e_multigpu.py 8 > gpurun_out/r2c18_cli_e2e_8gpu.json 2> gpurun_out/r2c18_cli_e2e_8gpu.err; echo "cli e2e 8gpu exit $?"; cat gpurun_out/r2c18_cli_e2e_8gpu.json; tail -3 gpurun_out/r2c18_cli_e2e_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2c18_bench_8gpu.json 2> gpurun_out/r2c18_bench_8gpu.err; echo "bench exit $?"; tail -2 gpurun_out/r2c18_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c18_bench_8gpu.json').read().strip().splitlines()[-1])
def show(tag,o):
    r=o.get('roofline',{})
    print(tag, round(o['value'],1), o['unit'], 'e2e', round(o['e2e']['value'],1), 'frac', round(r.get('frac',0),3), 'traffic', r.get('traffic'), o.get('clocks'))
show('train',d)
for k,v in d.get('extra',{}).items(): show(k,v)
PY
