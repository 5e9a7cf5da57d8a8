#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
PS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_find_lambdas|k_solve_fluids' -s 20 -c 2 -o gpurun_out/prof_r1c python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
