#!/bin/bash
# concurrent decode chains (graph slots / streams) 4 / 8 (default) / 12 / 16, and the decoder's image-side projections back on fp32
mkdir -p gpurun_out
for S in 8 4 12 16; do
  SAMPT_DECODE_STREAMS=$S timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c14_bench_s$S.log 2>&1
  echo "decode streams $S rc=$?: $(tail -1 gpurun_out/c14_bench_s$S.log | cut -c100-260)"
done
SAMPT_DECODER_TC=0 timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c14_bench_notc.log 2>&1
echo "decoder TC off rc=$?: $(tail -1 gpurun_out/c14_bench_notc.log | cut -c100-260)"
