mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -30 > gpurun_out/t_all.log
timeout -s KILL 900 python bench.py --config C2 --steps 3 --warmup 2 --no-cpu-baseline --breakdown --kernel-table gpurun_out/kernel_table_c2_p3.md > gpurun_out/bench_c2_p3.log 2>&1
SAMPT_OVERLAP=0 timeout -s KILL 900 python bench.py --config C2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c2_p3_nooverlap.log 2>&1
tail -n 5 gpurun_out/t_all.log; tail -n 1 gpurun_out/bench_c2_p3.log;  tail -n 1 gpurun_out/bench_c2_p3_nooverlap.log
