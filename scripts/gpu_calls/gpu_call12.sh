#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/check_attn_bwd.py > gpurun_out/c12_check.log 2>&1
grep -c OK gpurun_out/c12_check.log; grep -v OK gpurun_out/c12_check.log
timeout 200 python scripts/bench_attn.py --json gpurun_out/c12_attn.json > gpurun_out/c12_attn.log 2>&1
grep bwd gpurun_out/c12_attn.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attn_bwd_prep|attn_bwd_tc' --csv --log-file gpurun_out/c12_bwd_launches.csv python scripts/bench_attn.py --once > /dev/null 2>&1
grep -v "^==" gpurun_out/c12_bwd_launches.csv | cut -d, -f5,15- | head -12
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c12_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c12_tests.log
tail -4 gpurun_out/c12_tests.log
