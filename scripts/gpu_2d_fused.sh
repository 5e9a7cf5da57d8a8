#!/bin/bash
# the fused 2-D tick against the launch sequence: parity test, then ms per tick per scene at thresholds 0 (never) / 2048 (always, where it fits)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_2d_fused.py tests/test_gpu_2d.py tests/test_gpu_2d_full.py -x -q -m gpu 2>&1 | tail -5
CLI=particlesolver_b200/psolver_cli
OUT=gpurun_out/r2zz_2d_fused_tick.jsonl
: > $OUT
for key in 8 7 6 2 1 0 w v; do
  for thr in 0 2048; do
    for rep in 1 2; do
      PS2D_FUSED_MAX_N=$thr timeout 300 $CLI --app cpu --scene $key --ticks 400 --json | sed "s/^{/{\"fused_max_n\": $thr, /" >> $OUT
    done
  done
done
timeout 300 oracle/_ref/ref_cpu --scene 8 --ticks 400 --json | grep '^{' >> $OUT
timeout 300 oracle/_ref/ref_cpu --scene 8 --ticks 400 --json | grep '^{' >> $OUT
python - <<'PY'
import json
for l in open('gpurun_out/r2zz_2d_fused_tick.jsonl'):
    d = json.loads(l)
    print(d.get('impl', 'ours'), d.get('scene'), d.get('fused_max_n'), d.get('particles', d.get('n')), d.get('wall_ms_per_tick', d.get('ms_per_tick')), d.get('launches_per_tick'))
PY
