#!/bin/bash
# round 2, call 9: producer reordering + next-item query prefetch; chunk sweep on one box; full GPU suite; BLIP graphs default
mkdir -p gpurun_out
timeout 900 python scripts/bench_search_modes.py 190000 5600000 3 "UNIIR_SEARCH_MODE=1" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=96" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=128" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=192" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=256" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=384" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=256,UNIIR_SEARCH_EPIPIPE=1" 2>&1 | tee gpurun_out/r2c9_modes_full.txt
timeout 600 python scripts/bench_search_modes.py 190000 700000 5 "UNIIR_SEARCH_MODE=1" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=128" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=256" 2>&1 | tee gpurun_out/r2c9_modes_shard.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c9_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c9_tests.log
tail -6 gpurun_out/r2c9_tests.log
