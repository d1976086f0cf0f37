#!/bin/bash
# call 2: LN-backward load batching fix + skip-masked-towers path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_clip_gpu.py tests/test_blip_gpu.py -x -q -m gpu > gpurun_out/c2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c2_tests.log
timeout 300 python scripts/bench_rowops.py --json gpurun_out/c2_rowops.json > gpurun_out/c2_rowops.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/c2_bench_train.json 2> gpurun_out/c2_bench_train.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --mask-mix mbeir > gpurun_out/c2_bench_mix.json 2> gpurun_out/c2_bench_mix.err
timeout 600 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --mask-mix mbeir --no-tower-skip > gpurun_out/c2_bench_mix_noskip.json 2> gpurun_out/c2_bench_mix_noskip.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'layernorm_bwd_warp' -c 2 -o gpurun_out/c2_lnbwd_prof python scripts/bench_rowops.py --once > gpurun_out/c2_ncu.log 2>&1
tail -5 gpurun_out/c2_tests.log; cat gpurun_out/c2_rowops.log; for f in gpurun_out/c2_bench_*.json; do cut -c1-200 $f; echo; done
