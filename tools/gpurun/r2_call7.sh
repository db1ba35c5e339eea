#!/bin/bash
# fp8-corrected GEMM (ViT precision 6): unit tests, encoder parity, full-clip parity, A/B bench against precision 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/c7_smi.txt
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -s > gpurun_out/c7_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -3 gpurun_out/c7_gemm.log
timeout 900 python -m pytest tests/test_gpu_sam.py -q -s -k "encoder or precision or hq" > gpurun_out/c7_sam.log 2>&1; echo "sam tests rc=$?"; tail -5 gpurun_out/c7_sam.log
SAMPT_VIT_PRECISION=6 timeout 900 python -m pytest tests/test_gpu_full_configs.py -q -s > gpurun_out/c7_full_p6.log 2>&1; echo "full p6 rc=$?"; grep "full:\|passed\|failed" gpurun_out/c7_full_p6.log
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --precision 6 --kernel-table gpurun_out/kernel_table_p6.md > gpurun_out/c7_bench_p6.log 2>&1; echo "bench p6 rc=$?"; tail -1 gpurun_out/c7_bench_p6.log | cut -c1-400
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --precision 4 > gpurun_out/c7_bench_p4.log 2>&1; echo "bench p4 rc=$?"; tail -1 gpurun_out/c7_bench_p4.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --precision 6 > gpurun_out/c7_bench_p6b.log 2>&1; echo "bench p6 (2nd) rc=$?"; tail -1 gpurun_out/c7_bench_p6b.log | cut -c1-300
