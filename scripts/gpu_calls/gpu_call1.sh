#!/bin/bash
# round-1 session-2 call 1: verify the warp-per-row LayerNorm / vectorised colsum kernels, measure them
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/c1_gpu.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_clip_gpu.py tests/test_blip_gpu.py -x -q -m gpu > gpurun_out/c1_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c1_tests.log
timeout 300 python scripts/bench_rowops.py --json gpurun_out/c1_rowops.json > gpurun_out/c1_rowops.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/c1_bench_train.json 2> gpurun_out/c1_bench_train.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'layernorm_.*warp|colsum_vec' -c 14 -o gpurun_out/c1_rowops_prof python scripts/bench_rowops.py --once > gpurun_out/c1_ncu.log 2>&1
tail -5 gpurun_out/c1_tests.log; cat gpurun_out/c1_rowops.log; cat gpurun_out/c1_bench_train.json | cut -c1-600
