#!/bin/bash
# parity + bench + ncu of the fluid kernels after the accept/interact split and the chunked cell table
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
PS_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
PS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_find_lambdas|k_solve_fluids|k_radix_pass|k_radix_hist|k_reorder|k_cell_begin' -s 40 -c 9 -o gpurun_out/prof_r1b python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
# the reference's own CUDA code on the same B200 at the bench size (its neighbour lists need 8 KB/particle)
timeout 300 oracle/_ref/ref_gpu --scene c3 --side 100 --grid 256 --max 1100000 --mode whole --steps 4 --dump-every 0 --out /tmp/ref_c3 > gpurun_out/ref_gpu_c3.log 2>&1; echo "ref_gpu c3 rc=$?"; tail -3 gpurun_out/ref_gpu_c3.log
ls -la gpurun_out
