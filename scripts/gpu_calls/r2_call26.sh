#!/bin/bash
# round 2, call 26 (1 GPU): the GPU suite after the change of the mask hash (kernels vs the integer restatement, BLIP / CLIP_FF train mode)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -40 > gpurun_out/r2c26_tests.log
tail -4 gpurun_out/r2c26_tests.log
