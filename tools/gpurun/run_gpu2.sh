mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt 2>&1
timeout -s KILL 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 800 2>&1 | tail -15 > gpurun_out/t_multi.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_c2_n2.log 2>&1
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --mgpu-mode clip_per_gpu --no-cpu-baseline > gpurun_out/bench_c2_n2_clip.log 2>&1
cat gpurun_out/smi2.txt; tail -n 6 gpurun_out/t_multi.log; tail -n 3 gpurun_out/bench_c2_n2.log | cut -c1-1200; tail -n 2 gpurun_out/bench_c2_n2_clip.log | cut -c1-600
