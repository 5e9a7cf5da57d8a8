#!/bin/bash
# r2h: full GPU suite on the current build, then both bench arms
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/r2h_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2h_pytest_all.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err; echo "ref rc=$?"; cut -c1-2200 gpurun_out/r2h_bench_ref.json; tail -4 gpurun_out/r2h_bench_ref.err
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2h_bench.json'));print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','long_run')})"; tail -4 gpurun_out/r2h_bench.err
