#!/bin/bash
# round summary pass (r1m): full GPU suite, smoke, bench (ours + reference arm), c5 on one GPU, ncu launch list, ncu --set full of the
# PBF / grid kernels, memcheck + racecheck over the slab tests (halo / migration / ghost-lambda kernels)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r1m.json; tail -2 gpurun_out/bench_r1m.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1m_ref.json 2> gpurun_out/bench_r1m_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --workload c5 --particles 8000000 --steps 10 > gpurun_out/bench_r1m_c5_8M.json 2> gpurun_out/bench_r1m_c5_8M.err; echo "c5 rc=$?"; cut -c1-200 gpurun_out/bench_r1m_c5_8M.json
PS_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1m.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
PS_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_find_lambdas|k_solve_fluids|k_radix_pass|k_reorder|k_cell_begin' -s 40 -c 8 -o gpurun_out/prof_r1m python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_slab.py -m gpu -q -k "lambda or one_gpu or pack" > gpurun_out/sanitize_slab_mem.log 2>&1; echo "sanitize mem rc=$? $(grep 'ERROR SUMMARY' gpurun_out/sanitize_slab_mem.log | sort | uniq -c | tr '\n' ';')"; grep -E "passed|failed" gpurun_out/sanitize_slab_mem.log | tail -1
timeout 900 $CS --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_slab.py -m gpu -q -k "lambda_exchange_kernels or pack" > gpurun_out/sanitize_slab_race.log 2>&1; echo "sanitize race rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/sanitize_slab_race.log | sort | uniq -c | tr '\n' ';')"
