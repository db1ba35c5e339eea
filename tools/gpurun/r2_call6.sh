#!/bin/bash
# round-2 call 6: whole GPU suite at the final defaults (ViT precision 4, attn_ws, attn_prep2, decoder on tcgen05), smoke, bench +
# reference arm, ncu launch list + full captures of the roofline kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -s ) > $O/c6_tests.log 2>&1
echo "suite rc=$?"; grep -E "passed|failed|error" $O/c6_tests.log | tail -3; grep -E "^FAILED|^ERROR" $O/c6_tests.log | head -12; grep -E "full:|threshold" $O/c6_tests.log | grep -v print | head
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/c6_smoke.log 2>&1
echo "smoke rc=$?"; grep "smoke" $O/c6_smoke.log | tail -4
( time timeout 500 python bench.py --kernel-table $O/kernel_table_c6.md ) > $O/c6_bench.log 2>&1
echo "bench rc=$?"; grep '^{' $O/c6_bench.log | cut -c1-3000
SAMPT_ATTN_PREP2=0 timeout 300 python bench.py --no-cpu-baseline > $O/c6_bench_prep1.log 2>&1
echo "old prep:"; grep '^{' $O/c6_bench_prep1.log | cut -c1-220
( time timeout 300 python bench.py --impl reference ) > $O/c6_bench_ref.log 2>&1
echo "ref rc=$?"; grep '^{' $O/c6_bench_ref.log | cut -c1-400
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_c2p_r02.csv python bench.py --config C2p --steps 1 --warmup 1 --no-cpu-baseline > $O/c6_ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l $O/launches_c2p_r02.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'attn_ws|gemm_tc|pips_corr_kernel|attn_prep2' -c 14 -f -o $O/prof_r02_roofline python tools/ncu_targets.py > $O/c6_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i $O/prof_r02_roofline.ncu-rep --page raw --csv > $O/prof_r02_roofline_raw.csv 2>/dev/null; ls -la $O/prof_r02_roofline* | head
