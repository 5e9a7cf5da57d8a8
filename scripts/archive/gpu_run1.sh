#!/bin/bash
# First GPU pass of a round: parity tests, smoke, bench, ncu launch list, ncu --set full of the neighbour kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "benchref rc=$?"; cat gpurun_out/bench_ref.json
PS_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
PS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_find_lambdas|k_solve_fluids|k_radix_pass|k_reorder|k_cell_begin' -s 40 -c 8 -o gpurun_out/prof_r1a python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
