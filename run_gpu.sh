mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -30 > gpurun_out/t_all.log
timeout -s KILL 900 python bench.py --config C2 --steps 3 --warmup 2 --no-cpu-baseline --breakdown > gpurun_out/bench_c2_p3.log 2>&1
SAMPT_OVERLAP=0 timeout -s KILL 900 python bench.py --config C2 --steps 1 --warmup 1 --no-cpu-baseline --kernel-table gpurun_out/kernel_table_c2_p3.md > gpurun_out/bench_c2_p3_nooverlap.log 2>&1
timeout -s KILL 300 python tools/microbench.py > gpurun_out/microbench.log 2>&1
tail -n 5 gpurun_out/t_all.log; tail -n 1 gpurun_out/bench_c2_p3.log | cut -c1-300; tail -n 1 gpurun_out/bench_c2_p3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['breakdown_ms'], d.get('roofline_corr_gather'))"
