#!/bin/bash
# every BASELINE config on one B200, this library (psolver_cli) beside the reference's own unmodified code:
#   3-D scenes: oracle/_ref/ref_gpu (the reference's CUDA sources compiled for sm_100a) vs psolver_cli --app gpu
#   2-D scenes: oracle/_ref/ref_cpu (the reference's CPU solver, one host core) vs psolver_cli --app cpu
mkdir -p gpurun_out
OUT=gpurun_out/scene_table.jsonl
: > $OUT
CLI=particlesolver_b200/psolver_cli
for spec in "2 64 15000 100" "c2 64 66000 256" "7 64 15000 100" "8 64 15000 100" "c3 256 1001024 100"; do
  set -- $spec
  timeout 300 oracle/_ref/ref_gpu --scene $1 --mode whole --grid $2 --max $3 --side $4 --steps 30 --out /tmp/refgpu_$1 2>/dev/null | grep '^{' >> $OUT
  timeout 300 $CLI --app gpu --scene $1 --grid $2 --max-particles $3 --side $4 --steps 100 --json >> $OUT
done
for key in 6 1 2 3 7 8 0 w v; do
  timeout 300 oracle/_ref/ref_cpu --scene $key --ticks 200 --json | grep '^{' >> $OUT
  timeout 300 $CLI --app cpu --scene $key --ticks 200 --json >> $OUT
done
cat $OUT
