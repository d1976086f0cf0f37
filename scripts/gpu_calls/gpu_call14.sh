#!/bin/bash
# call 14: 16 elementwise warps in the tcgen05 attention backward (A/B against 8) + launch list of one training micro-batch
mkdir -p gpurun_out
timeout 600 python scripts/check_attn_bwd.py > gpurun_out/c14_check.log 2>&1
grep -c OK gpurun_out/c14_check.log; grep -v OK gpurun_out/c14_check.log
timeout 200 python scripts/bench_attn.py --json gpurun_out/c14_attn_ew16.json > gpurun_out/c14_attn_ew16.log 2>&1
grep bwd gpurun_out/c14_attn_ew16.log
UNIIR_ATTN_BWD_EW=8 timeout 200 python scripts/bench_attn.py --json gpurun_out/c14_attn_ew8.json > gpurun_out/c14_attn_ew8.log 2>&1
grep bwd gpurun_out/c14_attn_ew8.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c14_step_launches.csv python scripts/prof_step.py 64 128 > gpurun_out/c14_ncu_step.log 2>&1
python scripts/ncu_summarize.py gpurun_out/c14_step_launches.csv > gpurun_out/c14_step_summary.txt 2>&1
head -30 gpurun_out/c14_step_summary.txt
