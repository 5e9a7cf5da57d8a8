#!/bin/bash
# r2f: streamed host I/O (ps_step_streamed) tests, full bench line (pipelined e2e, long_run, c5_8M_1gpu), reference arm with the reference's CUDA binary
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_stream_io.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2f_pytest.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"; cut -c1-1800 gpurun_out/r2f_bench_ref.json; tail -4 gpurun_out/r2f_bench_ref.err
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; cut -c1-6000 gpurun_out/r2f_bench.json; tail -4 gpurun_out/r2f_bench.err
