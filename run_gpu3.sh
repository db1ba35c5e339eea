#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/debug_cotracker.py > gpurun_out/cot_debug.log 2>&1
echo "debug rc=$?"
tail -22 gpurun_out/cot_debug.log
timeout 420 python -m pytest tests/test_gpu_cotracker.py -q -s > gpurun_out/cot_tests.log 2>&1
echo "cot tests rc=$?"
grep -n "cotracker \|passed\|failed\|Error\|assert" gpurun_out/cot_tests.log | head -30
