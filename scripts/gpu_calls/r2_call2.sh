#!/bin/bash
# round 2, call 2: where do the retrieval kernel's 3.2 TB of DRAM reads come from?  L2 counters by source unit
# (SM/TEX vs the other L2 partition over the fabric) for 1 / 2 / 10 rounds of query tiles per chunk, the single-CTA
# kernel, and TMA eviction-priority hints.
mkdir -p gpurun_out
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_ltcfabric_lookup_hit.sum,lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum,lts__d_sectors_fill_device.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max
run() {  # tag env Q N
  env $2 timeout 300 ncu --metrics $M --clock-control none -k regex:ip_topk --csv --log-file gpurun_out/r2c2_$1.csv python scripts/prof_search_shape.py $3 $4 > gpurun_out/r2c2_$1.log 2>&1
  echo "== $1 ($2 Q=$3 N=$4)"; grep -v "^==" gpurun_out/r2c2_$1.csv | cut -d, -f13- | tr -d '"' | column -s, -t | head -14
}
run q1round   X=1 18944 5600000
run q2rounds  X=1 37888 5600000
run q10rounds X=1 189440 5600000
run q10_1cta  UNIIR_SEARCH_2CTA=0 189440 5600000
run q10_h30   UNIIR_SEARCH_HINT=30 189440 5600000
run q10_h32   UNIIR_SEARCH_HINT=32 189440 5600000
run q10_h23   UNIIR_SEARCH_HINT=23 189440 5600000
run q10_h33   UNIIR_SEARCH_HINT=33 189440 5600000
# un-profiled timings of the promising variants
for h in 0 30 32 23 33; do UNIIR_SEARCH_HINT=$h timeout 120 python scripts/prof_search_shape.py 190000 5600000 | tail -1 | sed "s/^/hint $h: /"; done
