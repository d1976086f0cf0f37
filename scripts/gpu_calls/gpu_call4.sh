#!/bin/bash
# call 4: tcgen05 attention forward v2 (prefetched TMEM loads, boundary-only masks, tail warp, 3-deep P ring)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" > gpurun_out/c4_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c4_tests.log
timeout 200 python scripts/bench_attn.py --json gpurun_out/c4_attn.json > gpurun_out/c4_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd_tc' -c 1 -o gpurun_out/c4_attn_tc_prof python scripts/bench_attn.py --once > gpurun_out/c4_ncu.log 2>&1
tail -8 gpurun_out/c4_tests.log; cat gpurun_out/c4_attn.log
