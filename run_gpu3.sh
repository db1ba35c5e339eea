#!/bin/bash
# experiment run: new inorm statistics kernel (parity via the PIPS / CoTracker suites), then small tuning sweeps on C2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_pips.py tests/test_gpu_cotracker.py -q > gpurun_out/exp_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/exp_tests.log
run() { # name, env..., args
  name=$1; shift
  timeout 300 env "$@" > gpurun_out/exp_$name.log 2>&1
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
for l in open(f"gpurun_out/exp_{name}.log"):
    if l.startswith("{"):
        d=json.loads(l); print(name, "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms", round(d["ms_per_step"],1), d["clocks"])
PY
}
run base   X=1 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --kernel-table gpurun_out/exp_kernel_table_c2.md
run eb25   X=1 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --encoder-batch 25
run ds8    SAMPT_DECODE_STREAMS=8 python bench.py --no-cpu-baseline --steps 3 --warmup 3
run ds2    SAMPT_DECODE_STREAMS=2 python bench.py --no-cpu-baseline --steps 3 --warmup 3
