#!/bin/bash
# run every diagnostic case in its own process with a timeout; log to gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/check.log
for c in "$@"; do
  echo "=== $c" | tee -a gpurun_out/check.log
  timeout 180 python scripts/gpu_check.py $c 2>&1 | tail -40 | tee -a gpurun_out/check.log
done
