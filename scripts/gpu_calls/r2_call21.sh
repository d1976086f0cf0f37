#!/bin/bash
# round 2, call 21 (1 GPU): the full-size property tests (configs[4] retrieval, configs[1] training step)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_full_size_gpu.py -m gpu -x -q -s --durations=5 2>&1 | tail -40 | tee gpurun_out/r2c21_fullsize.txt
