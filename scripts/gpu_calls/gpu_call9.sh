#!/bin/bash
# call 9: tcgen05 attention backward v2 (epilogue warps, 4 stages, thread-per-row prep) + full GPU suite + train bench
mkdir -p gpurun_out
timeout 600 python scripts/check_attn_bwd.py > gpurun_out/c9_check.log 2>&1
cat gpurun_out/c9_check.log
timeout 200 python scripts/bench_attn.py --json gpurun_out/c9_attn.json > gpurun_out/c9_attn.log 2>&1
cat gpurun_out/c9_attn.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attn_bwd|attn_delta' --csv --log-file gpurun_out/c9_bwd_launches.csv python scripts/bench_attn.py --once > /dev/null 2>&1
grep -v "^==" gpurun_out/c9_bwd_launches.csv | cut -d, -f5,15- | head -12
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c9_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c9_tests.log
tail -5 gpurun_out/c9_tests.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/c9_bench_train.json 2> gpurun_out/c9_bench_train.err
cat gpurun_out/c9_bench_train.json | cut -c1-400
