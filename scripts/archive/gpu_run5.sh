#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest.log
timeout 600 python bench.py --quick --steps 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench quick rc=$?"; cut -c1-600 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
timeout 600 python bench.py --workload c5 --particles 4000000 --steps 5 > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err; echo "bench c5 rc=$?"; cat gpurun_out/bench_c5_1gpu.json; tail -5 gpurun_out/bench_c5_1gpu.err
