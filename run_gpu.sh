mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -s KILL 400 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 120 2>&1 | tail -40 > gpurun_out/t_gemm.log
timeout -s KILL 400 python -m pytest tests/test_gpu_attention.py -q -m gpu --timeout 120 2>&1 | tail -40 > gpurun_out/t_attn.log
timeout -s KILL 900 python -m pytest tests/test_gpu_sam.py -q -m gpu --timeout 300 2>&1 | tail -80 > gpurun_out/t_sam.log
timeout -s KILL 300 python bench.py --config C1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c1.log 2>&1
timeout -s KILL 600 python bench.py --config C2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1
tail -5 gpurun_out/t_gemm.log gpurun_out/t_attn.log gpurun_out/t_sam.log gpurun_out/bench_c1.log gpurun_out/bench_c2.log
