#!/bin/bash
# CoTracker UpdateFormer on tcgen05 (3-pass), fixed tests, C3 / C5 single-GPU lines with kernel tables
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pips.py tests/test_gpu_cotracker.py -q > gpurun_out/c9_units.log 2>&1; echo "unit tests rc=$?"; tail -3 gpurun_out/c9_units.log
timeout 600 python -m pytest tests/test_gpu_sam.py -q -k "hq" > gpurun_out/c9_hq.log 2>&1; echo "hq tests rc=$?"; tail -3 gpurun_out/c9_hq.log
timeout 900 python -m pytest tests/test_gpu_full_configs.py -q -s > gpurun_out/c9_full.log 2>&1; echo "full rc=$?"; grep "full:\|passed\|failed" gpurun_out/c9_full.log
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --config C3 --kernel-table gpurun_out/kernel_table_c3_c9.md > gpurun_out/c9_bench_c3.log 2>&1; echo "bench C3 rc=$?"; tail -1 gpurun_out/c9_bench_c3.log | cut -c1-300
SAMPT_COT_TC=0 timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --config C3 > gpurun_out/c9_bench_c3_fp32.log 2>&1; echo "bench C3 (fp32 UpdateFormer) rc=$?"; tail -1 gpurun_out/c9_bench_c3_fp32.log | cut -c1-300
timeout 500 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --config C5 --kernel-table gpurun_out/kernel_table_c5_c9.md > gpurun_out/c9_bench_c5.log 2>&1; echo "bench C5 rc=$?"; tail -1 gpurun_out/c9_bench_c5.log | cut -c1-400
SAMPT_COT_TC=0 timeout 500 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --config C5 > gpurun_out/c9_bench_c5_fp32.log 2>&1; echo "bench C5 (fp32 UpdateFormer) rc=$?"; tail -1 gpurun_out/c9_bench_c5_fp32.log | cut -c1-300
