#!/bin/bash
# round 2, call 12 (2 GPUs): NCCL correctness test, the CLI under torch.distributed.run against the single-process CLI,
# the embed -> shard files -> sharded search pipeline at reduced size, the default bench line at N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_nccl_2gpu.py -x -q -m gpu > gpurun_out/r2c12_nccl_test.log 2>&1; echo "nccl test exit $?"; tail -4 gpurun_out/r2c12_nccl_test.log
timeout 900 python scripts/cli_e2e_multigpu.py 2 > gpurun_out/r2c12_cli_e2e_2gpu.json 2> gpurun_out/r2c12_cli_e2e_2gpu.err; echo "cli e2e 2gpu exit $?"; cat gpurun_out/r2c12_cli_e2e_2gpu.json; tail -5 gpurun_out/r2c12_cli_e2e_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 scripts/pipeline_embed_search.py --items 200000 --queries 50000 > gpurun_out/r2c12_pipeline_2gpu.json 2> gpurun_out/r2c12_pipeline_2gpu.err; echo "pipeline exit $?"; cat gpurun_out/r2c12_pipeline_2gpu.json; tail -3 gpurun_out/r2c12_pipeline_2gpu.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2c12_bench_2gpu.json 2> gpurun_out/r2c12_bench_2gpu.err; echo "bench exit $?"; tail -3 gpurun_out/r2c12_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c12_bench_2gpu.json').read().strip().splitlines()[-1])
def show(tag,o):
    r=o.get('roofline',{})
    print(tag, round(o['value'],1), o['unit'], 'e2e', round(o['e2e']['value'],1), 'frac', round(r.get('frac',0),3), o.get('clocks'))
show('train',d)
for k,v in d.get('extra',{}).items(): show(k,v)
PY
