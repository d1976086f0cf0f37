#!/bin/bash
# call 16 (2 GPUs): data-parallel training step and sharded retrieval through torchrun / NCCL
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c16_bench_2gpu.json 2> gpurun_out/c16_bench_2gpu.err
cut -c1-400 gpurun_out/c16_bench_2gpu.json; tail -n 5 gpurun_out/c16_bench_2gpu.err
