#!/bin/bash
# round 2, call 17: final state -- full GPU suite, smoke(), default bench line, search bench line, DRAM counters of the
# final kernel at the four shard sizes, ncu --set full summary
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c17_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r2c17_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r2c17_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c17_bench.json'))
def show(tag,o):
    r=o.get('roofline',{})
    print(tag, round(o['value'],1), o['unit'], 'e2e', round(o['e2e']['value'],1), 'frac', round(r.get('frac',0),3), 'iso', round(r.get('isolated_frac',0) or 0,3), 'step_alg', round(r.get('frac_step_algorithmic',0) or 0,3), 'traffic', r.get('traffic'), 'cpu', o.get('cpu_baseline',{}).get('value'), o.get('clocks'))
show('train',d)
for k,v in d.get('extra',{}).items(): show(k,v)
PY
timeout 600 python bench.py --workload search --steps 3 --warmup 3 > gpurun_out/r2c17_bench_search.json 2> gpurun_out/r2c17_bench_search.err; cut -c1-300 gpurun_out/r2c17_bench_search.json
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max
for n in 5600000 2800000 1400000 700000; do
  timeout 400 ncu --metrics $M --clock-control none -k regex:ip_topk --csv --log-file gpurun_out/dram_search_r02_N$n.csv python scripts/prof_search_shape.py 190000 $n > /dev/null 2>&1
  grep -v "^==" gpurun_out/dram_search_r02_N$n.csv | cut -d, -f13- | tr -d '"' | head -9 | tr '\n' ' '; echo
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ip_topk -o gpurun_out/r2c17_ts python scripts/prof_search_shape.py 37888 1400000 > gpurun_out/r2c17_ts.log 2>&1
python scripts/ncu_metrics.py gpurun_out/r2c17_ts.ncu-rep > gpurun_out/ncu_search_kernel_ts_r02.txt 2>&1; grep -E "time_duration|tensor_cycles|hit_rate|dram__bytes_read.sum \[" gpurun_out/ncu_search_kernel_ts_r02.txt
