#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/trace_attn.py > gpurun_out/c6_trace.log 2>&1
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" > gpurun_out/c6_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c6_tests.log
timeout 200 python scripts/bench_attn.py --json gpurun_out/c6_attn.json > gpurun_out/c6_attn.log 2>&1
cat gpurun_out/c6_trace.log; tail -4 gpurun_out/c6_tests.log; cat gpurun_out/c6_attn.log
