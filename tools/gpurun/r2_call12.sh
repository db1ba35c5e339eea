#!/bin/bash
# two-call refine form with an empty positive set; re-validation of the decode chain and the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sam.py -q > gpurun_out/c12_sam.log 2>&1; echo "sam tests rc=$?"; tail -3 gpurun_out/c12_sam.log
timeout 900 python -m pytest tests/test_gpu_full_configs.py tests/test_gpu_cotracker.py -q -s > gpurun_out/c12_full.log 2>&1; echo "full rc=$?"; grep "full:\|passed\|failed" gpurun_out/c12_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c12_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke gpurun_out/c12_smoke.log | tail -3
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/c12_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/c12_bench.log | cut -c1-300
