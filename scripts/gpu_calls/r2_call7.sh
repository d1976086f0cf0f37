#!/bin/bash
# round 2, call 7: accumulator-release arrivals without the GPU-scope fence (retrieval kernel + every CTA-pair GEMM)
mkdir -p gpurun_out
for cfg in "2 512" "2 2048" "2 4096" "1 64"; do set -- $cfg
  UNIIR_SEARCH_MODE=$1 UNIIR_SEARCH_TPC=$2 timeout 200 python scripts/prof_search_shape.py 190000 5600000 2>&1 | tail -1 | sed "s/^/mode $1 tpc $2: /"
done
UNIIR_SEARCH_MODE=2 UNIIR_SEARCH_TPC=2048 timeout 200 python scripts/prof_search_shape.py 190000 700000 2>&1 | tail -1 | sed "s/^/mode 2 shard tpc 2048: /"
UNIIR_SEARCH_MODE=2 UNIIR_SEARCH_TPC=512 timeout 200 python scripts/prof_search_shape.py 190000 700000 2>&1 | tail -1 | sed "s/^/mode 2 shard tpc 512: /"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c7_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c7_tests.log
tail -6 gpurun_out/r2c7_tests.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2c7_bench_train.json 2> gpurun_out/r2c7_bench_train.err
cut -c1-600 gpurun_out/r2c7_bench_train.json
UNIIR_SEARCH_MODE=2 UNIIR_SEARCH_TPC=2048 timeout 300 ncu --set full --import-source on --clock-control none -k regex:ip_topk -o gpurun_out/r2c7_ts_small python scripts/prof_search_shape.py 37888 1400000 > gpurun_out/r2c7_ts_small.log 2>&1
