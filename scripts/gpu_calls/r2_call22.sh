#!/bin/bash
# round 2, call 22 (1 GPU): the whole GPU suite with the train-mode dropout path (new kernels, BLIP step vs oracle + masks)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=8 2>&1 | tail -150 > gpurun_out/r2c22_tests.log
tail -60 gpurun_out/r2c22_tests.log
