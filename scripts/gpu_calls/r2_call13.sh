#!/bin/bash
# round 2, call 13: is the per-launch fixed cost of the retrieval kernel the top-k warm-up (survivor path)?  k = 1/10/50
# at the 700 k-row shard and at 2.8 M rows; attention-backward FOLD variant check (never run on hardware before)
mkdir -p gpurun_out
for n in 700000 2800000; do for k in 1 10 50; do
  timeout 200 python scripts/prof_search_shape.py 190000 $n $k 2>&1 | tail -1 | sed "s/^/k=$k: /"
done; done
UNIIR_ATTN_BWD_FOLD=1 timeout 600 python scripts/check_attn_bwd.py > gpurun_out/r2c13_fold_check.log 2>&1; echo "fold check exit $?"; tail -25 gpurun_out/r2c13_fold_check.log
UNIIR_ATTN_BWD_FOLD=1 timeout 200 python scripts/bench_attn.py --json gpurun_out/r2c13_attn_fold.json > gpurun_out/r2c13_attn_fold.log 2>&1; tail -12 gpurun_out/r2c13_attn_fold.log
timeout 200 python scripts/bench_attn.py --json gpurun_out/r2c13_attn_nofold.json > gpurun_out/r2c13_attn_nofold.log 2>&1; tail -12 gpurun_out/r2c13_attn_nofold.log
