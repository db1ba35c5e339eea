#!/bin/bash
# 2 GPUs: frame-sharded parity (PIPS + CoTracker, ragged clips, rotated ownership) and the N=2 bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L > $O/m2_smi.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 800 ) > $O/m2_tests.log 2>&1
echo "multi tests rc=$?"; grep -E "passed|failed|skipped" $O/m2_tests.log | tail -2; grep -E "^FAILED|^E  " $O/m2_tests.log | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/m2_bench_c2.log 2>&1
echo "C2 N=2 rc=$?"; grep '^{' $O/m2_bench_c2.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 --config C3 --no-cpu-baseline > $O/m2_bench_c3.log 2>&1
echo "C3 N=2 rc=$?"; grep '^{' $O/m2_bench_c3.log | cut -c1-500; grep -E "Error|error" $O/m2_bench_c3.log | head -5
# single-GPU extras riding on this call: the two fixed tests, attn_prep2 (cp.async rewrite) parity + A/B
timeout 300 python -m pytest tests/test_gpu_reinit_patch.py tests/test_gpu_sam.py -q -k "patch_similarity or hq_encoder" > $O/m2_fixed.log 2>&1
echo "fixed tests rc=$?"; grep -E "passed|failed" $O/m2_fixed.log | tail -1
SAMPT_ATTN_PREP2=1 timeout 400 python -m pytest tests/test_gpu_sam.py tests/test_gpu_full_configs.py -q -s -k "encoder or c2_full or c1_end" > $O/m2_prep2.log 2>&1
echo "prep2 parity rc=$?"; grep -E "passed|failed|full:" $O/m2_prep2.log | grep -v print | tail -3
SAMPT_ATTN_PREP2=1 timeout 300 python bench.py --no-cpu-baseline --kernel-table $O/kernel_table_prep2b.md > $O/m2_bench_prep2.log 2>&1
echo "prep2 bench:"; grep '^{' $O/m2_bench_prep2.log | cut -c1-200; grep -E "attn_prep" $O/kernel_table_prep2b.md | cut -c1-160
timeout 300 python bench.py --no-cpu-baseline > $O/m2_bench_prep1.log 2>&1
echo "prep1 bench:"; grep '^{' $O/m2_bench_prep1.log | cut -c1-200
