#!/bin/bash
# r2q: BASELINE.md section 4's scaled replicas of CPU scene 6 (432 / 1,610 / 6,348 particles) — the reference's unmodified CPU solver on one
# host core beside the 2-D GPU path; where the single-CTA level-scheduled design stops scaling
mkdir -p gpurun_out
OUT=gpurun_out/r2q_scene6_replicas.jsonl
: > $OUT
CLI=particlesolver_b200/psolver_cli
for spec in "4 6 200" "8 6x2 60" "16 6x4 12"; do
  set -- $spec
  timeout 600 oracle/_ref/ref_cpu --scene 6 --fluid-scale $1 --ticks $3 --json | grep '^{' >> $OUT
  timeout 600 $CLI --app cpu --scene $2 --ticks $3 --json >> $OUT
done
cat $OUT
