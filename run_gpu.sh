mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests/test_gpu_sam.py -q -m gpu --timeout 600 2>&1 | tail -30 > gpurun_out/t_sam.log
timeout -s KILL 900 python bench.py --config C2 --steps 3 --warmup 2 --no-cpu-baseline --breakdown > gpurun_out/bench_c2_p3.log 2>&1
tail -n 12 gpurun_out/t_sam.log; cat gpurun_out/precision_dial_c2slice.json; cat gpurun_out/precision_dial_c1.json; tail -n 1 gpurun_out/bench_c2_p3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['breakdown_ms'])"
