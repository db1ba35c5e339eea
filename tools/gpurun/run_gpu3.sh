#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_cotracker.py -q -s -k "end_to_end" > gpurun_out/fix2_tests.log 2>&1
echo "rc=$?"; grep -E "passed|failed|SamPt \+ CoTracker|Error" gpurun_out/fix2_tests.log | head
