#!/bin/bash
# call 8: per-kernel durations of the backward kernels + one full ncu capture of the tcgen05 backward (ViT-L/14 shape)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attn_bwd|attn_delta' --csv --log-file gpurun_out/c8_bwd_launches.csv python scripts/bench_attn.py --once > gpurun_out/c8_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_tc_kernel' -c 1 -o gpurun_out/c8_attn_bwd_tc_prof -f python scripts/bench_attn.py --once > gpurun_out/c8_ncu2.log 2>&1
grep -v "^==" gpurun_out/c8_bwd_launches.csv | cut -d, -f5,15- | head -40
tail -3 gpurun_out/c8_ncu2.log
