#!/bin/bash
# final validation of the round: whole GPU suite, smoke, bench + reference arm, ncu launch list + full captures of the roofline kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/final_smi.txt
( time timeout 1800 python -m pytest tests -q -m gpu -s ) > $O/final_tests.log 2>&1
echo "suite rc=$?"; grep -E "passed|failed|error" $O/final_tests.log | tail -3; grep -E "^FAILED|^ERROR" $O/final_tests.log | head -12; grep -E "full:|threshold" $O/final_tests.log | grep -v print | head
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/final_smoke.log 2>&1
echo "smoke rc=$?"; grep "smoke" $O/final_smoke.log | tail -4
( time timeout 500 python bench.py --kernel-table $O/kernel_table_final.md ) > $O/final_bench.log 2>&1
echo "bench rc=$?"; grep '^{' $O/final_bench.log | cut -c1-3500
( time timeout 300 python bench.py --impl reference ) > $O/final_bench_ref.log 2>&1
echo "ref rc=$?"; grep '^{' $O/final_bench_ref.log | cut -c1-400
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_c2p_r02.csv python bench.py --config C2p --steps 1 --warmup 1 --no-cpu-baseline > $O/final_ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l $O/launches_c2p_r02.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'attn_ws|gemm_tc|pips_corr_kernel' -c 16 -f -o $O/prof_r02_roofline python tools/ncu_targets.py > $O/final_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i $O/prof_r02_roofline.ncu-rep --page raw --csv > $O/prof_r02_roofline_raw.csv 2>/dev/null; ls -la $O/prof_r02_roofline* | head
NCU_TARGET=pips timeout 300 ncu --set full --clock-control none --import-source on -k regex:'sgemm_skinny' --launch-skip 30 -c 3 -f -o $O/prof_r02_skinny python tools/ncu_targets.py > $O/final_ncu_skinny.log 2>&1
echo "ncu skinny rc=$?"
ncu -i $O/prof_r02_skinny.ncu-rep --page raw --csv > $O/prof_r02_skinny_raw.csv 2>/dev/null; ls -la $O/prof_r02_skinny* | head
