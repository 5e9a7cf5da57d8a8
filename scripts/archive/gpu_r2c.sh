#!/bin/bash
# r2c: staged K6 v2 (parallel row windows, chunking, static lists): parity, variants at two points of the C3 run, ncu
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extensions.py tests/test_gpu_full_size.py tests/test_gpu_headline_parity.py -m gpu -q -x ) > gpurun_out/r2c_pytest.log 2>&1; echo "parity rc=$?"; tail -15 gpurun_out/r2c_pytest.log
: > gpurun_out/r2c_variants.jsonl
for w in 5 100; do
  timeout 300 python bench.py --quick --steps 20 --warmup $w | tee -a gpurun_out/r2c_variants.jsonl | cut -c1-700
  for v in r1 c2048 s512; do
    PS_LIBRARY=$PWD/particlesolver_b200/libpsolver_$v.so timeout 300 python bench.py --quick --steps 20 --warmup $w | tee -a gpurun_out/r2c_variants.jsonl | cut -c1-700
  done
done
PS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_find_lambdas|k_solve_fluids' -s 16 -c 4 -o gpurun_out/prof_r2c python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r2c_ncu_full.log 2>&1; echo "ncu full rc=$?"
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wrap or cap_500 or cell or paths_agree" > gpurun_out/r2c_sanitize_mem.log 2>&1; echo "sanitize mem rc=$? $(grep 'ERROR SUMMARY' gpurun_out/r2c_sanitize_mem.log | sort | uniq -c | tr '\n' ';')"; grep -E "passed|failed" gpurun_out/r2c_sanitize_mem.log | tail -1
