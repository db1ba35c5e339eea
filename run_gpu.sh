mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --config C2 --steps 2 --warmup 1 --no-cpu-baseline --breakdown > gpurun_out/bench_c2_p3.log 2>&1
timeout -s KILL 600 python bench.py --config C2 --steps 2 --warmup 1 --no-cpu-baseline --precision 1 > gpurun_out/bench_c2_p1.log 2>&1
timeout -s KILL 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 70000 --csv --log-file gpurun_out/launches_c2.csv python bench.py --config C2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 2 -o gpurun_out/prof_gemm python bench.py --config C2b --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
for f in gpurun_out/bench_c2_p3.log gpurun_out/bench_c2_p1.log; do tail -n 1 $f; done
ls -la gpurun_out
