#!/bin/bash
# round 2, call 8: interleaved same-box A/B of the retrieval kernel variants; BLIP hard-negative + engine tests
mkdir -p gpurun_out
timeout 900 python scripts/bench_search_modes.py 190000 5600000 3 "UNIIR_SEARCH_MODE=1" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=256" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=512" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=512,UNIIR_SEARCH_EPIPIPE=0" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=1024" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=2048" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=512,UNIIR_SEARCH_ROT=0" 2>&1 | tee gpurun_out/r2c8_modes_full.txt
timeout 600 python scripts/bench_search_modes.py 190000 700000 5 "UNIIR_SEARCH_MODE=1" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=128" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=256" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=512" "UNIIR_SEARCH_MODE=2,UNIIR_SEARCH_TPC=1024" 2>&1 | tee gpurun_out/r2c8_modes_shard.txt
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_engine_gpu.py tests/test_blip_gpu.py tests/test_cli_e2e_gpu.py -x -q -m gpu > gpurun_out/r2c8_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c8_tests.log
tail -6 gpurun_out/r2c8_tests.log
