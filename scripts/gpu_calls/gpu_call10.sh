#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/check_graph_modes.py > gpurun_out/c10_modes.log 2>&1
cat gpurun_out/c10_modes.log | tail -20
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_clip_gpu.py::test_train_step_graphs_equal_eager_and_loss_decreases > gpurun_out/c10_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c10_tests.log
tail -15 gpurun_out/c10_tests.log
