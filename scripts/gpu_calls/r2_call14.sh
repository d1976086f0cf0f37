#!/bin/bash
# round 2, call 14: deferred row-buffer compaction (256-entry buffers) in the queries-on-chip retrieval kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py tests/test_cli_e2e_gpu.py -x -q -m gpu > gpurun_out/r2c14_search_tests.log 2>&1; echo "search tests exit $?"; tail -4 gpurun_out/r2c14_search_tests.log
for n in 700000 2800000 5600000; do for k in 1 50; do
  timeout 200 python scripts/prof_search_shape.py 190000 $n $k 2>&1 | tail -1 | sed "s/^/k=$k: /"
done; done
timeout 900 python scripts/bench_search_modes.py 190000 5600000 3 "UNIIR_SEARCH_MODE=1" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=64" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=96" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=128" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=192" 2>&1 | tee gpurun_out/r2c14_modes_full.txt
timeout 600 python scripts/bench_search_modes.py 190000 700000 7 "UNIIR_SEARCH_MODE=1" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=64" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=128" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=256" 2>&1 | tee gpurun_out/r2c14_modes_shard.txt
