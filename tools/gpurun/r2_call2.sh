#!/bin/bash
# round-2 call 2: whole GPU suite with the round-1 drafts on by default + the new rows (full-clip parity, query points, VOS tail)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -s -x --deselect tests/test_gpu_full_configs.py ) > $O/c2_tests.log 2>&1
echo "suite rc=$?"; grep -E "passed|failed|error" $O/c2_tests.log | tail -3; grep -E "^FAILED|^ERROR|Error" $O/c2_tests.log | head -8
( time timeout 900 python -m pytest tests/test_gpu_full_configs.py -q -s ) > $O/c2_full.log 2>&1
echo "full rc=$?"; grep -E "full:|passed|failed" $O/c2_full.log | tail -6; grep -E "^FAILED|^ERROR|Error|assert" $O/c2_full.log | head -8
( time timeout 400 python bench.py --kernel-table $O/kernel_table_c2.md ) > $O/c2_bench.log 2>&1
echo "bench rc=$?"; grep '^{' $O/c2_bench.log | cut -c1-1500
( time timeout 300 python bench.py --impl reference ) > $O/c2_bench_ref.log 2>&1
echo "ref rc=$?"; grep '^{' $O/c2_bench_ref.log | cut -c1-600; nproc
