#!/bin/bash
# the fused 2-D tick (one cluster of CTAs) against the launch sequence: parity tests, sanitizer, then ms per tick per scene
#   PS2D_FUSED_MAX_N: 0 = never fused, 2048 = always (where it fits); PS2D_FUSED_CLUSTER: CTAs of the cluster (default 16)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_2d_fused.py tests/test_gpu_2d.py tests/test_gpu_2d_full.py tests/test_gpu_scenes2d.py tests/test_gpu_checkpoint.py -x -q -m gpu 2>&1 | tail -3
CLI=particlesolver_b200/psolver_cli
for k in 6 w; do
  compute-sanitizer --tool memcheck $CLI --app cpu --scene $k --ticks 12 --json 2>&1 | grep -E "ERROR SUMMARY"
  compute-sanitizer --tool racecheck $CLI --app cpu --scene $k --ticks 4 --json 2>&1 | grep -E "RACECHECK SUMMARY"
done
OUT=gpurun_out/r2zz_2d_fused_tick.jsonl
: > $OUT
for cl in 1 2 4 8 16; do
  for key in 8 6; do
    PS2D_FUSED_CLUSTER=$cl timeout 300 $CLI --app cpu --scene $key --ticks 200 --json | sed "s/^{/{\"cluster\": $cl, /" >> $OUT
  done
done
for key in n 8 d s v 7 0 6 2 w 1 3; do
  for thr in 0 2048; do
    for rep in 1 2; do
      PS2D_FUSED_MAX_N=$thr timeout 300 $CLI --app cpu --scene $key --ticks 200 --json | sed "s/^{/{\"fused_max_n\": $thr, /" >> $OUT
    done
  done
  timeout 300 oracle/_ref/ref_cpu --scene $key --ticks 200 --json | grep '^{' >> $OUT
done
PS2D_FUSED_PROFILE=1 $CLI --app cpu --scene 8 --ticks 200 --json > /dev/null 2> gpurun_out/r2zz_2d_fused_phases.txt
PS2D_FUSED_PROFILE=1 $CLI --app cpu --scene 6 --ticks 200 --json > /dev/null 2>> gpurun_out/r2zz_2d_fused_phases.txt
cat gpurun_out/r2zz_2d_fused_phases.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2zz_2d_fused_tick.jsonl'):
    d = json.loads(l)
    print(d.get('impl', 'ours'), 'scene', d.get('scene'), 'cluster', d.get('cluster'), 'max_n', d.get('fused_max_n'), 'n', d.get('particles', d.get('n')), 'ms', d.get('wall_ms_per_tick', d.get('ms_per_tick')), 'launches', d.get('launches_per_tick'))
PY
