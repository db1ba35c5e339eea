mkdir -p gpurun_out
( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench_default.log 2>&1
( time timeout -s KILL 900 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_reference.log 2>&1
( time timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
tail -n 4 gpurun_out/bench_default.log | cut -c1-2500; tail -n 4 gpurun_out/bench_reference.log | cut -c1-900; tail -n 6 gpurun_out/smoke.log; nproc
