#!/bin/bash
# round-2 first call: gates of the round-1 drafts + what each buys
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
for t in attention_v2_multi attention_v2_inside vit_skip gemm_cta_pair attention_v3; do
  SAMPT_TEST_EXPERIMENTAL=1 timeout 700 python -m pytest tests/test_gpu_experimental.py -q -x -k "$t" > $O/exp_$t.log 2>&1
  echo "exp $t rc=$?"; tail -n 25 $O/exp_$t.log | grep -E "passed|failed|Error|error|mismatch|assert" | head -8
done
timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --kernel-table $O/kernel_table_base.md > $O/bench_base.log 2>&1
echo "base rc=$?"; grep '^{' $O/bench_base.log | cut -c1-300
for f in SAMPT_ATTN_V2 SAMPT_ATTN_V3 SAMPT_VIT_SKIP_PAD SAMPT_GEMM_2CTA; do
  env $f=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > $O/bench_$f.log 2>&1
  echo "$f rc=$?"; grep '^{' $O/bench_$f.log | cut -c1-300
done
