#!/bin/bash
# round 2, call 6: hybrid queries-on-chip retrieval kernel (TMEM + smem resident queries, 128-candidate tiles);
# the new CLI end-to-end test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py -x -q -m gpu > gpurun_out/r2c6_search_tests.log 2>&1; echo "search tests (TS) exit $?"; tail -5 gpurun_out/r2c6_search_tests.log
for cfg in "2 128" "2 512" "2 2048" "1 64"; do set -- $cfg
  UNIIR_SEARCH_MODE=$1 UNIIR_SEARCH_TPC=$2 timeout 200 python scripts/prof_search_shape.py 190000 5600000 2>&1 | tail -1 | sed "s/^/mode $1 tpc $2: /"
done
UNIIR_SEARCH_MODE=2 timeout 200 python scripts/prof_search_shape.py 190000 700000 2>&1 | tail -1 | sed "s/^/mode 2 shard: /"
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active
UNIIR_SEARCH_MODE=2 timeout 300 ncu --metrics $M --clock-control none -k regex:ip_topk --csv --log-file gpurun_out/r2c6_ts_full.csv python scripts/prof_search_shape.py 190000 5600000 > gpurun_out/r2c6_ts_full.log 2>&1
grep -v "^==" gpurun_out/r2c6_ts_full.csv | cut -d, -f13- | tr -d '"' | head -12
UNIIR_SEARCH_MODE=2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:ip_topk -o gpurun_out/r2c6_ts_small python scripts/prof_search_shape.py 37888 1400000 > gpurun_out/r2c6_ts_small.log 2>&1
timeout 900 python -m pytest tests/test_cli_e2e_gpu.py -x -q -m gpu > gpurun_out/r2c6_cli.log 2>&1; echo "cli e2e exit $?"; tail -25 gpurun_out/r2c6_cli.log
