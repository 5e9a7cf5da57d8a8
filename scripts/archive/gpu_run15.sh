#!/bin/bash
# extensions (shape matching, viscosity) + parity regression after the K6 tuning + quick bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extensions.py -m gpu -q > gpurun_out/pytest_ext.log 2>&1; echo "ext rc=$?"; tail -25 gpurun_out/pytest_ext.log
echo skip
timeout 300 python bench.py --quick --steps 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_quick.json
