#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py --quick --steps 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench quick rc=$?"; cut -c1-700 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
bash scripts/gpu_run10.sh
