#!/bin/bash
# r2a: first run of the staged (TMA) K6: smoke, K6/K7 parity tests, headline-size parity, variant A/B, memcheck
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extensions.py -m gpu -q -x ) > gpurun_out/r2a_pytest_parity.log 2>&1; echo "parity rc=$?"; tail -15 gpurun_out/r2a_pytest_parity.log
( time timeout 1200 python -m pytest tests/test_gpu_headline_parity.py -m gpu -q ) > gpurun_out/r2a_pytest_headline.log 2>&1; echo "headline rc=$?"; tail -30 gpurun_out/r2a_pytest_headline.log
: > gpurun_out/r2a_variants.jsonl
timeout 300 python bench.py --quick --steps 20 --warmup 5 | tee -a gpurun_out/r2a_variants.jsonl | cut -c1-600
for v in r1 s128 s512 c2048; do
  PS_LIBRARY=$PWD/particlesolver_b200/libpsolver_$v.so timeout 300 python bench.py --quick --steps 20 --warmup 5 | tee -a gpurun_out/r2a_variants.jsonl | cut -c1-600
done
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wrap or cap_500 or cell or paths_agree" > gpurun_out/r2a_sanitize_mem.log 2>&1; echo "sanitize mem rc=$? $(grep 'ERROR SUMMARY' gpurun_out/r2a_sanitize_mem.log | sort | uniq -c | tr '\n' ';')"; grep -E "passed|failed" gpurun_out/r2a_sanitize_mem.log | tail -1
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_gpu_headline_parity.py ) > gpurun_out/r2a_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -25 gpurun_out/r2a_pytest_all.log
