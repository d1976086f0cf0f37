#!/bin/bash
# round 2, call 11: evidence for profiles/ with the final retrieval kernel: DRAM traffic per launch at the 1/2/4/8-GPU
# shard sizes, ncu --set full of the kernel, launch list of the training step (same command as round 1), search bench line
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max
for n in 5600000 2800000 1400000 700000; do
  timeout 400 ncu --metrics $M --clock-control none -k regex:ip_topk --csv --log-file gpurun_out/dram_search_r02_N$n.csv python scripts/prof_search_shape.py 190000 $n > /dev/null 2>&1
  grep -v "^==" gpurun_out/dram_search_r02_N$n.csv | cut -d, -f13- | tr -d '"' | head -9
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ip_topk -o gpurun_out/r2c11_ts python scripts/prof_search_shape.py 37888 1400000 > gpurun_out/r2c11_ts.log 2>&1
python scripts/ncu_metrics.py gpurun_out/r2c11_ts.ncu-rep > gpurun_out/ncu_search_kernel_ts_r02.txt 2>&1; head -40 gpurun_out/ncu_search_kernel_ts_r02.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c11_step_launches.csv python scripts/prof_step.py 64 128 > /dev/null 2>&1
python scripts/ncu_summarize.py gpurun_out/r2c11_step_launches.csv > gpurun_out/launches_train_step_r02.txt; head -30 gpurun_out/launches_train_step_r02.txt
timeout 600 python bench.py --workload search --steps 3 --warmup 3 > gpurun_out/r2c11_bench_search.json 2> gpurun_out/r2c11_bench_search.err; cut -c1-900 gpurun_out/r2c11_bench_search.json
