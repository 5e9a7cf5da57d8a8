#!/bin/bash
# r2p: fused K6 — CTA size and resident-CTA target variants
mkdir -p gpurun_out
: > gpurun_out/r2p_variants.jsonl
for lib in default fb64 fb256 m6 m8; do
  if [ $lib = default ]; then unset PS_LIBRARY; else export PS_LIBRARY=$PWD/particlesolver_b200/libpsolver_$lib.so; fi
  for w in 5 100; do timeout 300 python bench.py --quick --steps 20 --warmup $w | tee -a gpurun_out/r2p_variants.jsonl | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['lib'][-22:], 'ms/step %.3f'%d['ms_per_step'], 'K6 %.4f K7 %.4f'%(d['stage_ms_per_launch']['lambda'], d['stage_ms_per_launch']['delta_p']))"; done
done
