#!/bin/bash
# r2l: K7 with its position / lambda gathers pinned ahead of the bodies (6 in flight per lane), fused hash + histogram, ps_io_begin / ps_io_end
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extensions.py tests/test_gpu_full_size.py tests/test_gpu_headline_parity.py tests/test_gpu_stream_io.py tests/test_gpu_slab.py tests/test_long_run_stats.py -m gpu -q -x ) > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2l_pytest.log
: > gpurun_out/r2l_variants.jsonl
for w in 5 100; do
  timeout 300 python bench.py --quick --steps 20 --warmup $w | tee -a gpurun_out/r2l_variants.jsonl | cut -c1-520
done
timeout 600 python - <<'PY' | tee -a gpurun_out/r2l_variants.jsonl
import json, bench
r = bench.c5_single_gpu(0, 6532.2)
print(json.dumps(r))
PY
