#!/bin/bash
# r2k: grid build — histogram fused into the hash kernel, look-back depth 8 / 16 / 32 — at 1M (c3) and 8M (c5 on one GPU)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_headline_parity.py -m gpu -q -x ) > gpurun_out/r2k_pytest.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/r2k_pytest.log
: > gpurun_out/r2k_variants.jsonl
for lib in default L16 L8; do
  if [ $lib = default ]; then unset PS_LIBRARY; else export PS_LIBRARY=$PWD/particlesolver_b200/libpsolver_$lib.so; fi
  timeout 300 python bench.py --quick --steps 20 --warmup 5 | tee -a gpurun_out/r2k_variants.jsonl | cut -c1-520
  timeout 600 python - <<'PY' | tee -a gpurun_out/r2k_variants.jsonl
import json, os, bench
r = bench.c5_single_gpu(0, 6532.2)
print(json.dumps({"lib": os.environ.get("PS_LIBRARY", "default"), "c5_8M_ms_per_step": r.get("ms_per_step"), "sort_ms_per_launch": r.get("kernels_at_8M", {}).get("sort"), "hash": r.get("kernels_at_8M", {}).get("hash"), "err": r.get("error")}))
PY
done
