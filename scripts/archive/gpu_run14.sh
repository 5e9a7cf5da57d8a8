#!/bin/bash
# a19: general 2-D path vs the reference CPU solver's golden states; deviation report first, then the asserting run
mkdir -p gpurun_out
rm -f gpurun_out/ps2d_report.txt
PS2D_REPORT=gpurun_out/ps2d_report.txt timeout 600 python -m pytest tests/test_gpu_2d_full.py -m gpu -q -k scene_matches > gpurun_out/pytest_2d_report.log 2>&1; echo "report rc=$?"
cat gpurun_out/ps2d_report.txt
tail -15 gpurun_out/pytest_2d_report.log
timeout 900 python -m pytest tests/test_gpu_2d_full.py tests/test_gpu_2d.py -m gpu -q > gpurun_out/pytest_2d.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_2d.log
