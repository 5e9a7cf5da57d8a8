#!/bin/bash
# r2r: fused K6 with the row set-up's cell-table loads issued per batch of rows (9 / 5 / 3 rows), no votes in the set-up
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_headline_parity.py tests/test_gpu_slab.py -m gpu -q -x ) > gpurun_out/r2r_pytest.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/r2r_pytest.log
: > gpurun_out/r2r_variants.jsonl
for lib in default rb5 rb3 rb9m6; do
  if [ $lib = default ]; then unset PS_LIBRARY; else export PS_LIBRARY=$PWD/particlesolver_b200/libpsolver_$lib.so; fi
  for w in 5 100; do timeout 300 python bench.py --quick --steps 20 --warmup $w | tee -a gpurun_out/r2r_variants.jsonl | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['lib'][-22:], 'ms/step %.3f'%d['ms_per_step'], 'K6 %.4f K7 %.4f'%(d['stage_ms_per_launch']['lambda'], d['stage_ms_per_launch']['delta_p']))"; done
done
