#!/bin/bash
# 8 GPUs: BASELINE configs[3] (8 x C2, frame-sharded + one all-gather), configs[4] (C5: ONE 100-frame 1080p clip, HQ-SAM + CoTracker, 256 pts:
# strong scaling) and 8 x C5 (weak); single-GPU C2 / C5 lines on the same box for the efficiency
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L > $O/m8_smi.txt 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > $O/m8_bench_c2.log 2>&1
echo "C2 N=8 rc=$?"; grep '^{' $O/m8_bench_c2.log | cut -c1-900; grep -E "Error" $O/m8_bench_c2.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 2 --warmup 3 --config C5 --no-cpu-baseline > $O/m8_bench_c5.log 2>&1
echo "C5 N=8 (1 clip) rc=$?"; grep '^{' $O/m8_bench_c5.log | cut -c1-900; grep -E "Error" $O/m8_bench_c5.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 2 --warmup 3 --config C5 --clips-per-step 8 --no-cpu-baseline > $O/m8_bench_c5_weak.log 2>&1
echo "C5 N=8 (8 clips) rc=$?"; grep '^{' $O/m8_bench_c5_weak.log | cut -c1-600; grep -E "Error" $O/m8_bench_c5_weak.log | head -5
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/m8_bench_c2_n1.log 2>&1
echo "C2 N=1 rc=$?"; grep '^{' $O/m8_bench_c2_n1.log | cut -c1-300
timeout 300 python bench.py --steps 2 --warmup 3 --config C5 --no-cpu-baseline > $O/m8_bench_c5_n1.log 2>&1
echo "C5 N=1 rc=$?"; grep '^{' $O/m8_bench_c5_n1.log | cut -c1-300
