#!/bin/bash
# call 19: fused InfoNCE kernels at the configs[1] shapes: CUDA-event timing, then one ncu capture (numbers printed
# under ncu are not timings)
mkdir -p gpurun_out
timeout 100 python scripts/prof_infonce.py --json gpurun_out/c19_infonce.json > gpurun_out/c19_infonce.log 2>&1
cat gpurun_out/c19_infonce.log | tail -3
timeout 120 ncu --set full --clock-control none --import-source on -k regex:'nce_' -s 24 -c 8 -o gpurun_out/c19_infonce_prof -f python scripts/prof_infonce.py > gpurun_out/c19_ncu.log 2>&1
grep -c "Profiling" gpurun_out/c19_ncu.log
