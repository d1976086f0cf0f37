#!/bin/bash
# round 2, call 3: staggered chunk walk in the retrieval kernel -- correctness (search tests + full GPU suite incl. the
# new ViT-L/14-dims / hard-negative-golden tests), throughput sweep over the chunk size, DRAM/L2 counters before/after.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_search_gpu.py -x -q -m gpu > gpurun_out/r2c3_search_tests.log 2>&1; echo "search tests exit $?"; tail -3 gpurun_out/r2c3_search_tests.log
for cfg in "0 64" "1 32" "1 64" "1 128" "1 256"; do set -- $cfg
  UNIIR_SEARCH_ROT=$1 UNIIR_SEARCH_TPC=$2 timeout 200 python scripts/prof_search_shape.py 190000 5600000 | tail -1 | sed "s/^/rot $1 tpc $2: /"
done
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__d_sectors_fill_device.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max
run() {  # tag env Q N
  env $2 timeout 300 ncu --metrics $M --clock-control none -k regex:ip_topk --csv --log-file gpurun_out/r2c3_$1.csv python scripts/prof_search_shape.py $3 $4 > gpurun_out/r2c3_$1.log 2>&1
  echo "== $1 ($2 Q=$3 N=$4)"; grep -v "^==" gpurun_out/r2c3_$1.csv | cut -d, -f13- | tr -d '"' | head -12
}
run rot_q1round  UNIIR_SEARCH_ROT=1 18944 5600000
run rot_full64   UNIIR_SEARCH_TPC=64 190000 5600000
run rot_full128  UNIIR_SEARCH_TPC=128 190000 5600000
run rot_shard700k UNIIR_SEARCH_ROT=1 190000 700000
run norot_shard700k UNIIR_SEARCH_ROT=0 190000 700000
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c3_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c3_tests.log
tail -5 gpurun_out/r2c3_tests.log
