#!/bin/bash
# frames per ViT launch (M = B x 4096 GEMM rows): 10 (default of the bench) vs 13 / 17 / 25
mkdir -p gpurun_out
for B in 10 13 17 25 10; do
  timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --encoder-batch $B > gpurun_out/c13_bench_b$B.log 2>&1
  echo "encoder-batch $B rc=$?: $(tail -1 gpurun_out/c13_bench_b$B.log | cut -c100-260)"
done
