#!/bin/bash
# re-entry check: full GPU suite, smoke, default bench (ours + reference arm)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
tail -16 gpurun_out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_r1g.json; tail -4 gpurun_out/bench_r1g.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_r1g_ref.json 2> gpurun_out/bench_r1g_ref.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/bench_r1g_ref.json; tail -4 gpurun_out/bench_r1g_ref.err
