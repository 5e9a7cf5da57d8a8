#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest.log
timeout 600 python bench.py --quick --steps 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench quick rc=$?"; cut -c1-700 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
timeout 600 python bench.py --workload c5 --particles 4000000 --steps 5 > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err; echo "bench c5 rc=$?"; cut -c1-400 gpurun_out/bench_c5_1gpu.json; tail -5 gpurun_out/bench_c5_1gpu.err
python tests/golden/make_stats_golden.py gpurun_out/golden > gpurun_out/stats_golden.log 2>&1; echo "stats golden rc=$?"; tail -4 gpurun_out/stats_golden.log
PS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_find_lambdas|k_solve_fluids' -s 30 -c 3 -o gpurun_out/prof_r1e python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
