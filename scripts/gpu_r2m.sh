#!/bin/bash
# r2m: K7 variants (list rows per trip, CTA size) on the pinned-gather K7
mkdir -p gpurun_out
: > gpurun_out/r2m_variants.jsonl
for lib in default u8 u4 b128 u8b128; do
  if [ $lib = default ]; then unset PS_LIBRARY; else export PS_LIBRARY=$PWD/particlesolver_b200/libpsolver_$lib.so; fi
  for w in 5 100; do timeout 300 python bench.py --quick --steps 20 --warmup $w | tee -a gpurun_out/r2m_variants.jsonl | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['lib'][-22:], 'ms/step %.3f'%d['ms_per_step'], 'K6 %.4f K7 %.4f'%(d['stage_ms_per_launch']['lambda'], d['stage_ms_per_launch']['delta_p']))"; done
done
