#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_cotracker.py tests/test_gpu_registry.py tests/test_gpu_sam.py -q -s -k "end_to_end or registry or alternate or plain_sam or predict_torch_matches or refine_chain or c1_end" > gpurun_out/fix_tests.log 2>&1
echo "rc=$?"; grep -E "passed|failed" gpurun_out/fix_tests.log | tail -2; grep -E "^FAILED|^ERROR|SamPt \+ CoTracker" gpurun_out/fix_tests.log | head
