#!/bin/bash
# call 17: BLIP_SF parity tests + BLIP regression + a short BLIP_SF training bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_blip_sf_gpu.py tests/test_blip_gpu.py -x -q -m gpu > gpurun_out/c17_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c17_tests.log
tail -25 gpurun_out/c17_tests.log
timeout 400 python bench.py --workload blip --blip-model sf --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/c17_bench_blip_sf.json 2> gpurun_out/c17_bench_blip_sf.err
cut -c1-300 gpurun_out/c17_bench_blip_sf.json; tail -n 4 gpurun_out/c17_bench_blip_sf.err
