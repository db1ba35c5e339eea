#!/bin/bash
# GEMM 7 stages + pipelined TMEM epilogue, row-blocked skinny GEMM, one-pass image->token attention; A/B and C3 / C5 lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pips.py tests/test_gpu_sam.py tests/test_gpu_gemm.py tests/test_gpu_cotracker.py -q > gpurun_out/c10_units.log 2>&1; echo "unit tests rc=$?"; tail -4 gpurun_out/c10_units.log
timeout 900 python -m pytest tests/test_gpu_full_configs.py -q -s > gpurun_out/c10_full.log 2>&1; echo "full rc=$?"; grep "full:\|passed\|failed" gpurun_out/c10_full.log
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --kernel-table gpurun_out/kernel_table_c10.md > gpurun_out/c10_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/c10_bench.log | cut -c1-300
SAMPT_GEMM_STAGES=6 timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c10_bench_s6.log 2>&1; echo "bench (6 stages) rc=$?"; tail -1 gpurun_out/c10_bench_s6.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c10_bench_b.log 2>&1; echo "bench (2nd) rc=$?"; tail -1 gpurun_out/c10_bench_b.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --config C3 --kernel-table gpurun_out/kernel_table_c3_c10.md > gpurun_out/c10_bench_c3.log 2>&1; echo "bench C3 rc=$?"; tail -1 gpurun_out/c10_bench_c3.log | cut -c1-300
timeout 500 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --config C5 --kernel-table gpurun_out/kernel_table_c5_c10.md > gpurun_out/c10_bench_c5.log 2>&1; echo "bench C5 rc=$?"; tail -1 gpurun_out/c10_bench_c5.log | cut -c1-300
