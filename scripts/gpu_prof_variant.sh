#!/bin/bash
# ncu --set full of the two PBF kernels for one library variant: scripts/gpu_prof_variant.sh <variant|default> <out-name>
V=$1; OUT=$2
if [ "$V" != "default" ]; then export PS_LIBRARY=$PWD/particlesolver_b200/libpsolver_$V.so; fi
PS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_find_lambdas|k_solve_fluids' -s 20 -c 2 -o gpurun_out/$OUT python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_$OUT.log 2>&1; echo "ncu rc=$?"
