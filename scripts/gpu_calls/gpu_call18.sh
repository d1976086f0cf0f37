#!/bin/bash
# call 18: chunked pinned-host upload (e2e), full GPU suite, smoke, embed workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c18_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c18_tests.log
tail -4 gpurun_out/c18_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/c18_smoke.log 2>&1; tail -n 2 gpurun_out/c18_smoke.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c18_bench_train.json 2> gpurun_out/c18_bench_train.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c18_bench_train.json").read().strip().splitlines()[-1])
    print("train", d["value"], "e2e", d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/c18_bench_train.err").read()[-1500:])
PY
timeout 300 python bench.py --workload embed --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c18_bench_embed.json 2> gpurun_out/c18_bench_embed.err
cut -c1-260 gpurun_out/c18_bench_embed.json; tail -n 3 gpurun_out/c18_bench_embed.err
