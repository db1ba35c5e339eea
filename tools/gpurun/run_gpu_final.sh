#!/bin/bash
# round-1 final validation on one B200: full GPU suite, smoke, both bench arms, roofline ncu captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
( time timeout 700 python -m pytest tests -q -m gpu -s ) > $O/final_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|error" $O/final_tests.log | tail -3; grep -E "^FAILED|^ERROR" $O/final_tests.log | head
( time timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/final_smoke.log 2>&1
echo "smoke rc=$?"; grep "smoke" $O/final_smoke.log | tail -4
( time timeout 400 python bench.py ) > $O/final_bench.log 2>&1
echo "bench rc=$?"; grep '^{' $O/final_bench.log | cut -c1-1800
( time timeout 400 python bench.py --impl reference ) > $O/final_bench_ref.log 2>&1
echo "bench ref rc=$?"; grep '^{' $O/final_bench_ref.log | cut -c1-900; grep real $O/final_bench.log $O/final_bench_ref.log
timeout 200 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 4 -c 3 -f -o $O/prof_roofline_gemm python tools/ncu_targets.py > $O/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:pips_corr -s 4 -c 3 -f -o $O/prof_roofline_corr python tools/ncu_targets.py > $O/ncu_corr.log 2>&1
echo "ncu corr rc=$?"
ncu -i $O/prof_roofline_gemm.ncu-rep --page raw --csv > $O/prof_roofline_gemm_raw.csv 2>/dev/null
ncu -i $O/prof_roofline_corr.ncu-rep --page raw --csv > $O/prof_roofline_corr_raw.csv 2>/dev/null
ls -la $O/*.ncu-rep $O/*_raw.csv
SAMPT_DECODE_STREAMS=12 timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > $O/final_ds12.log 2>&1
grep '^{' $O/final_ds12.log | cut -c1-260
timeout 300 python bench.py --config C3 --steps 2 --warmup 3 > $O/final_bench_c3.log 2>&1
grep '^{' $O/final_bench_c3.log | cut -c1-700
