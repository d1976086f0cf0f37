#!/bin/bash
# round 2, call 15 (2 GPUs): NCCL correctness test and the CLI under torch.distributed.run, after the fixes of call 12
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nccl_2gpu.py -x -q -m gpu > gpurun_out/r2c15_nccl_test.log 2>&1; echo "nccl test exit $?"; tail -4 gpurun_out/r2c15_nccl_test.log; grep -n "^E " gpurun_out/r2c15_nccl_test.log | head
timeout 900 python scripts/cli_e2e_multigpu.py 2 > gpurun_out/r2c15_cli_e2e_2gpu.json 2> gpurun_out/r2c15_cli_e2e_2gpu.err; echo "cli e2e 2gpu exit $?"; cat gpurun_out/r2c15_cli_e2e_2gpu.json; tail -5 gpurun_out/r2c15_cli_e2e_2gpu.err
