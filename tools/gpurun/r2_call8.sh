#!/bin/bash
# skinny clustered fp32 GEMM (PIPS mixer), persistent global-attention prep, proj GEMM in the fp8-corrected form, HQ e2e diagnostics
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pips.py tests/test_gpu_gemm.py tests/test_gpu_attention.py -x -q > gpurun_out/c8_units.log 2>&1; echo "unit tests rc=$?"; tail -3 gpurun_out/c8_units.log
timeout 900 python -m pytest tests/test_gpu_sam.py -q -k "encoder or precision3 or hq or c1_end or chain" > gpurun_out/c8_sam.log 2>&1; echo "sam tests rc=$?"; tail -4 gpurun_out/c8_sam.log
timeout 300 python tests/manual/debug_hq_e2e.py > gpurun_out/c8_hq_debug.log 2>&1; echo "hq debug rc=$?"; grep -v "^Loading" gpurun_out/c8_hq_debug.log | tail -12
SAMPT_DECODER_TC=0 timeout 300 python tests/manual/debug_hq_e2e.py > gpurun_out/c8_hq_debug_notc.log 2>&1; echo "hq debug (decoder TC off) rc=$?"; grep "frame\|flipped\|vis" gpurun_out/c8_hq_debug_notc.log | tail -8
timeout 900 python -m pytest tests/test_gpu_full_configs.py -q -s > gpurun_out/c8_full.log 2>&1; echo "full rc=$?"; grep "full:\|passed\|failed" gpurun_out/c8_full.log
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --kernel-table gpurun_out/kernel_table_c8.md > gpurun_out/c8_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/c8_bench.log | cut -c1-300
SAMPT_SGEMM_SKINNY=0 timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c8_bench_noskinny.log 2>&1; echo "bench (skinny off) rc=$?"; tail -1 gpurun_out/c8_bench_noskinny.log | cut -c1-300
SAMPT_ATTN_PREP2_GLOBAL=0 timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c8_bench_noglobal.log 2>&1; echo "bench (global prep2 off) rc=$?"; tail -1 gpurun_out/c8_bench_noglobal.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c8_bench_b.log 2>&1; echo "bench (2nd) rc=$?"; tail -1 gpurun_out/c8_bench_b.log | cut -c1-300
