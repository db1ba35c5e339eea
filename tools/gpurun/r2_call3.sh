#!/bin/bash
# round-2 call 3: the new warp-specialised attention kernel (TS MMA) -- unit test, encoder parity, timing, ncu -- + the fixed rows
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 240 python -m pytest tests/test_gpu_attention.py -q -s ) > $O/c3_attn.log 2>&1
echo "attn unit rc=$?"; grep -E "passed|failed|assert|Error" $O/c3_attn.log | head -8
if grep -q "failed\|Error" $O/c3_attn.log || ! grep -q passed $O/c3_attn.log; then
  echo "attn_ws FAILED its unit test: the rest of the call runs with SAMPT_ATTN_WS=0"; export SAMPT_ATTN_WS=0
fi
( time timeout 400 python -m pytest tests/test_gpu_sam.py -q -s -k "encoder or c1_end or c2_slice" ) > $O/c3_enc.log 2>&1
echo "encoder rc=$?"; grep -E "passed|failed" $O/c3_enc.log | tail -2; grep -E "^FAILED|Error" $O/c3_enc.log | head
( time timeout 600 python -m pytest tests/test_gpu_query_points.py tests/test_gpu_vos_tail.py tests/test_gpu_full_configs.py -q -s ) > $O/c3_rows.log 2>&1
echo "rows rc=$?"; grep -E "full:|passed|failed" $O/c3_rows.log | tail -6; grep -E "^FAILED|^E  " $O/c3_rows.log | head -12
( time timeout 400 python bench.py --no-cpu-baseline --kernel-table $O/kernel_table_c3.md ) > $O/c3_bench.log 2>&1
echo "bench rc=$?"; grep '^{' $O/c3_bench.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value']); print(json.dumps(d.get('roofline_attention'))[:1500])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ws -c 2 -f -o $O/prof_attn_ws python tools/ncu_targets.py > $O/c3_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i $O/prof_attn_ws.ncu-rep --page raw --csv > $O/prof_attn_ws_raw.csv 2>/dev/null
python - <<'PY'
import csv
try:
    rows = list(csv.reader(open("gpurun_out/prof_attn_ws_raw.csv")))
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active"]
    for r in rows[2:]:
        print({w[:48]: r[hdr.index(w)][:60] for w in want if w in hdr})
except Exception as e:
    print("no ncu raw page:", e)
PY
