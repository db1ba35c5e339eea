mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -30 > gpurun_out/t_all.log
timeout -s KILL 900 python bench.py --config C2 --steps 2 --warmup 1 --no-cpu-baseline --breakdown --kernel-table gpurun_out/kernel_table_c2_p3.md > gpurun_out/bench_c2_p3.log 2>&1
timeout -s KILL 600 python bench.py --config C2 --steps 2 --warmup 1 --no-cpu-baseline --precision 2 > gpurun_out/bench_c2_p2.log 2>&1
SAMPT_DECODE_GRAPHS=0 SAMPT_PIPS_GRAPHS=0 timeout -s KILL 900 ncu --kernel-name-base demangled -k regex:sampt --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_c2p.csv python bench.py --config C2p --steps 1 --warmup 1 --no-cpu-baseline --encoder-batch 4 > gpurun_out/ncu_bench.log 2>&1
tail -n 3 gpurun_out/t_all.log; tail -n 1 gpurun_out/bench_c2_p3.log;  tail -n 1 gpurun_out/bench_c2_p2.log; wc -l gpurun_out/launches_c2p.csv; cat gpurun_out/precision_dial_c2slice.json
