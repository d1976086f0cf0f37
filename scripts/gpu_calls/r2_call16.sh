#!/bin/bash
# round 2, call 16: selection-based compaction (256-entry row buffers) -- correctness, k = 1 / 50, same-box A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py tests/test_cli_e2e_gpu.py -x -q -m gpu > gpurun_out/r2c16_search_tests.log 2>&1; echo "search tests exit $?"; tail -4 gpurun_out/r2c16_search_tests.log
timeout 900 python scripts/bench_search_modes.py 190000 5600000 3 "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_SELECT=0" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_SELECT=1" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_SELECT=1,UNIIR_SEARCH_TPC=96" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_SELECT=1,UNIIR_SEARCH_TPC=192" 2>&1 | tee gpurun_out/r2c16_modes_full.txt
timeout 600 python scripts/bench_search_modes.py 190000 700000 7 "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_SELECT=0" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_SELECT=1" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_SELECT=1,UNIIR_SEARCH_TPC=64" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_SELECT=1,UNIIR_SEARCH_TPC=256" 2>&1 | tee gpurun_out/r2c16_modes_shard.txt
for n in 700000 5600000; do for k in 1 50; do
  timeout 200 python scripts/prof_search_shape.py 190000 $n $k 2>&1 | tail -1 | sed "s/^/k=$k: /"
done; done
