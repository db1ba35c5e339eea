#!/bin/bash
# round-2 call 4: event-driven MMA issue in attn_ws; parity of the full configs vs ViT precision / attention kernel; boundary tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 240 python -m pytest tests/test_gpu_attention.py -q -s ) > $O/c4_attn.log 2>&1
echo "attn unit rc=$?"; grep -E "passed|failed|assert|Error" $O/c4_attn.log | head -5
( time timeout 900 python -m pytest tests/test_gpu_pips.py tests/test_gpu_cotracker.py tests/test_gpu_query_points.py tests/test_gpu_sam.py -q -x -k "not precision_dial and not c2_slice" ) > $O/c4_units.log 2>&1
echo "units rc=$?"; grep -E "passed|failed" $O/c4_units.log | tail -2; grep -E "^FAILED|^E  " $O/c4_units.log | head -10
for cfg in "3 1" "4 1" "3 0"; do
  set -- $cfg
  SAMPT_VIT_PRECISION=$1 SAMPT_ATTN_WS=$2 timeout 600 python -m pytest tests/test_gpu_full_configs.py -q -s > $O/c4_full_p$1_ws$2.log 2>&1
  echo "full precision=$1 attn_ws=$2 rc=$?"; grep -E "full:" $O/c4_full_p$1_ws$2.log | grep -v print | head -4
  for c in C2 C3 C5s; do cp $O/full_config_parity_$c.json $O/full_config_parity_${c}_p$1_ws$2.json 2>/dev/null; done
done
( time timeout 400 python bench.py --no-cpu-baseline --kernel-table $O/kernel_table_c4.md ) > $O/c4_bench.log 2>&1
echo "bench rc=$?"; grep '^{' $O/c4_bench.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value']); ra = d.get('roofline_attention') or {}
    print({k: (round(v['ms'], 4), round(v['frac'], 3), round(v['frac_issued'], 3)) for k, v in ra.items()})"
SAMPT_VIT_PRECISION=4 timeout 300 python bench.py --no-cpu-baseline --precision 4 > $O/c4_bench_p4.log 2>&1
grep '^{' $O/c4_bench_p4.log | cut -c1-200
