#!/bin/bash
# r2b: ncu --set full of the staged K6 and the list K7 (C3, 1M particles)
mkdir -p gpurun_out
PS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_find_lambdas_staged|k_solve_fluids_list' -s 12 -c 2 -o gpurun_out/prof_r2b python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r2b_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r2b_ncu_full.log
ls -la gpurun_out/prof_r2b.ncu-rep
