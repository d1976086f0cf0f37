#!/bin/bash
# round 2, call 4: queries-in-TMEM retrieval kernel -- correctness (search tests), throughput vs the streaming kernel,
# chunk-size sweep, counters; cuBLAS fp16 vs bf16 peak on this box; ViT-L/14-dims test.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py -x -q -m gpu > gpurun_out/r2c4_search_tests.log 2>&1; echo "search tests (TS) exit $?"; tail -15 gpurun_out/r2c4_search_tests.log
timeout 120 python scripts/peak_fp16.py | tee gpurun_out/r2c4_peak_fp16.json
for cfg in "1 64" "2 128" "2 256" "2 512" "2 1024"; do set -- $cfg
  UNIIR_SEARCH_MODE=$1 UNIIR_SEARCH_TPC=$2 timeout 200 python scripts/prof_search_shape.py 190000 5600000 2>&1 | tail -1 | sed "s/^/mode $1 tpc $2: /"
done
UNIIR_SEARCH_MODE=2 timeout 200 python scripts/prof_search_shape.py 190000 700000 2>&1 | tail -1 | sed "s/^/mode 2 shard: /"
UNIIR_SEARCH_MODE=1 timeout 200 python scripts/prof_search_shape.py 190000 700000 2>&1 | tail -1 | sed "s/^/mode 1 shard: /"
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,lts__cycles_elapsed.max
UNIIR_SEARCH_MODE=2 timeout 300 ncu --metrics $M --clock-control none -k regex:ip_topk --csv --log-file gpurun_out/r2c4_ts_full.csv python scripts/prof_search_shape.py 190000 5600000 > gpurun_out/r2c4_ts_full.log 2>&1
grep -v "^==" gpurun_out/r2c4_ts_full.csv | cut -d, -f13- | tr -d '"' | head -12
timeout 900 python -m pytest tests/test_clip_gpu.py -x -q -m gpu -k "vit_l14 or hard_negative" > gpurun_out/r2c4_l14.log 2>&1; echo "l14 exit $?"; tail -5 gpurun_out/r2c4_l14.log
