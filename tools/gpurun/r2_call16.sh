#!/bin/bash
# the default bench invocation with the in-step CUPTI cross-check of the roofline kernel
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/c16_bench.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/c16_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
print(json.dumps(d['roofline'].get('in_step'), indent=1))
"
