#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/check_attn_bwd.py > gpurun_out/c11_check.log 2>&1
grep -c OK gpurun_out/c11_check.log; grep -v OK gpurun_out/c11_check.log
timeout 200 python scripts/bench_attn.py --json gpurun_out/c11_attn.json > gpurun_out/c11_attn.log 2>&1
grep bwd gpurun_out/c11_attn.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attn_bwd_prep|attn_bwd_tc' --csv --log-file gpurun_out/c11_bwd_launches.csv python scripts/bench_attn.py --once > /dev/null 2>&1
grep -v "^==" gpurun_out/c11_bwd_launches.csv | cut -d, -f5,15- | head -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_tc_kernel' -c 1 -o gpurun_out/c11_attn_bwd_tc_prof -f python scripts/bench_attn.py --once > gpurun_out/c11_ncu2.log 2>&1
timeout 300 python scripts/check_determinism.py 0 > gpurun_out/c11_determinism.log 2>&1
cat gpurun_out/c11_determinism.log
