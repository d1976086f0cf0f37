#!/bin/bash
# call 7: first run of the tcgen05 attention backward
mkdir -p gpurun_out
timeout 900 python scripts/check_attn_bwd.py > gpurun_out/c7_check.log 2>&1
cat gpurun_out/c7_check.log
timeout 200 python scripts/bench_attn.py --json gpurun_out/c7_attn.json > gpurun_out/c7_attn.log 2>&1
cat gpurun_out/c7_attn.log
