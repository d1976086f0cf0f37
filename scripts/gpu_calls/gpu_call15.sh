#!/bin/bash
# call 15: full GPU suite, default bench line, BLIP bench, ncu captures of the tcgen05 attention kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c15_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c15_tests.log
tail -3 gpurun_out/c15_tests.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/c15_bench_default.json 2> gpurun_out/c15_bench_default.err
cut -c1-300 gpurun_out/c15_bench_default.json; tail -n 3 gpurun_out/c15_bench_default.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd_tc|attn_bwd_tc|attn_bwd_prep' -c 3 -o gpurun_out/c15_attn_tc_prof -f python scripts/bench_attn.py --once > gpurun_out/c15_ncu.log 2>&1
tail -n 2 gpurun_out/c15_ncu.log
timeout 600 python bench.py --workload blip --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/c15_bench_blip.json 2> gpurun_out/c15_bench_blip.err
cut -c1-300 gpurun_out/c15_bench_blip.json; tail -n 3 gpurun_out/c15_bench_blip.err
