#!/bin/bash
# call 3: full GPU suite (incl. skip-masked-towers + tcgen05 attention forward tests), attention micro-bench,
# training step with the tcgen05 attention forward enabled
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/c3_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c3_tests.log
timeout 200 python scripts/bench_attn.py --json gpurun_out/c3_attn.json > gpurun_out/c3_attn.log 2>&1
UNIIR_ATTN_TC=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c3_bench_train_tc.json 2> gpurun_out/c3_bench_train_tc.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd_tc' -c 1 -o gpurun_out/c3_attn_tc_prof python scripts/bench_attn.py --once > gpurun_out/c3_ncu.log 2>&1
tail -25 gpurun_out/c3_tests.log; cat gpurun_out/c3_attn.log; cut -c1-300 gpurun_out/c3_bench_train_tc.json; tail -3 gpurun_out/c3_bench_train_tc.err
