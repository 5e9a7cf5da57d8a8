#!/bin/bash
# round summary pass: full GPU suite, smoke, bench (ours + reference arm), ncu launch list, ncu --set full of the PBF kernels
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r1l.json; tail -2 gpurun_out/bench_r1l.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1l_ref.json 2> gpurun_out/bench_r1l_ref.err; echo "ref rc=$?"
PS_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
PS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_find_lambdas|k_solve_fluids|k_radix_pass|k_reorder|k_cell_begin' -s 40 -c 8 -o gpurun_out/prof_r1l python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
