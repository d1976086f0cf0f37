#!/bin/bash
# call 13: retained micro-batches (activations kept in HBM between the embedding pass and the backward)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_clip_gpu.py -x -q -m gpu > gpurun_out/c13_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c13_tests.log
tail -6 gpurun_out/c13_tests.log
# one rank's share of the 8-GPU job (512 pairs = 8 micro-batches of 128 items), with and without retention
timeout 300 python bench.py --steps 3 --warmup 3 --pairs-per-gpu 512 --no-extra --no-cpu-baseline --retain 0 > gpurun_out/c13_bench_512_retain0.json 2> gpurun_out/c13_a.err
timeout 300 python bench.py --steps 3 --warmup 3 --pairs-per-gpu 512 --no-extra --no-cpu-baseline > gpurun_out/c13_bench_512_auto.json 2> gpurun_out/c13_b.err
for f in gpurun_out/c13_bench_512_retain0.json gpurun_out/c13_bench_512_auto.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["config"]["grad_cache"], d["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/c13_a.err gpurun_out/c13_b.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
