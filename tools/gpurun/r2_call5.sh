#!/bin/bash
# round-2 call 5: mask decoder image-side GEMMs on tcgen05 (parity + speed), ViT precision modes vs the full-clip goldens
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; O=gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sam.py tests/test_gpu_registry.py -q -x -k "not precision_dial and not c2_slice and not encoder" ) > $O/c5_dec.log 2>&1
echo "decoder tests rc=$?"; grep -E "passed|failed" $O/c5_dec.log | tail -2; grep -E "^FAILED|^E  " $O/c5_dec.log | head -10
for p in 5 4; do
  SAMPT_VIT_PRECISION=$p timeout 600 python -m pytest tests/test_gpu_full_configs.py -q -s > $O/c5_full_p$p.log 2>&1
  echo "full precision=$p rc=$?"; grep -E "full:|threshold" $O/c5_full_p$p.log | grep -v print | head -6; grep -E "^E  " $O/c5_full_p$p.log | head -4
  for c in C2 C3 C5s; do cp $O/full_config_parity_$c.json $O/full_config_parity_${c}_p${p}_dectc.json 2>/dev/null; done
done
for cfg in "5 1" "4 1" "5 0"; do
  set -- $cfg
  SAMPT_DECODER_TC=$2 timeout 300 python bench.py --no-cpu-baseline --precision $1 --kernel-table $O/kernel_table_c5_p$1_tc$2.md > $O/c5_bench_p$1_tc$2.log 2>&1
  echo "bench precision=$1 decoder_tc=$2 rc=$?"; grep '^{' $O/c5_bench_p$1_tc$2.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'])"
done
