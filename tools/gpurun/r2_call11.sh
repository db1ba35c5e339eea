#!/bin/bash
# decoder token-side kernels (warp-per-head image->token attention, warp-per-query token self-attention), two-context test
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sam.py tests/test_gpu_registry.py tests/test_gpu_cotracker.py -q > gpurun_out/c11_units.log 2>&1; echo "unit tests rc=$?"; tail -4 gpurun_out/c11_units.log
timeout 900 python -m pytest tests/test_gpu_full_configs.py -q -s > gpurun_out/c11_full.log 2>&1; echo "full rc=$?"; grep "full:\|passed\|failed" gpurun_out/c11_full.log
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c11_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/c11_bench.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --config C3 --kernel-table gpurun_out/kernel_table_c3_c11.md > gpurun_out/c11_bench_c3.log 2>&1; echo "bench C3 rc=$?"; tail -1 gpurun_out/c11_bench_c3.log | cut -c1-300
timeout 500 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --config C5 --kernel-table gpurun_out/kernel_table_c5_c11.md > gpurun_out/c11_bench_c5.log 2>&1; echo "bench C5 rc=$?"; tail -1 gpurun_out/c11_bench_c5.log | cut -c1-300
