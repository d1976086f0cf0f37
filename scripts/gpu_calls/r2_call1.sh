#!/bin/bash
# round 2, call 1: cluster-occupancy probe, retrieval chunk-size sweep (throughput + clocks), DRAM/L2 counters of
# the full-size launch at two chunk sizes, GPU suite sanity.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r2c1_gpu.txt
scripts/probes/probe_clusters >> gpurun_out/r2c1_gpu.txt 2>&1
cat gpurun_out/r2c1_gpu.txt
for t in 8 16 32 64; do
  UNIIR_SEARCH_TPC=$t timeout 300 python bench.py --workload search --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2c1_search_tpc$t.json
  python -c "import json; d=json.load(open('gpurun_out/r2c1_search_tpc$t.json')); print('tpc', $t, round(d['value']), round(d['roofline']['achieved'],1), d['clocks'])"
done
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_elapsed.max,lts__cycles_elapsed.max
for t in 16 64; do
  UNIIR_SEARCH_TPC=$t timeout 400 ncu --metrics $M --clock-control none -k regex:ip_topk --csv --log-file gpurun_out/r2c1_dram_tpc$t.csv python scripts/prof_search_full.py > /dev/null 2>&1
  grep -v "^==" gpurun_out/r2c1_dram_tpc$t.csv | cut -d, -f5,13- | head -12
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c1_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c1_tests.log
tail -5 gpurun_out/r2c1_tests.log
