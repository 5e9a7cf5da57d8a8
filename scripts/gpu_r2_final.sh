#!/bin/bash
# r2 final (1 GPU): the whole GPU suite, smoke, both bench arms, scene table, ncu launch list + --set full capture, sanitizer — on the build that is committed
mkdir -p gpurun_out
T=r2z
( time timeout 1800 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/${T}_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/${T}_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2z_bench.json")); r=json.load(open("gpurun_out/r2z_bench_reference.json"))
print("ours", d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"], d["long_run"]["ms_per_step_mean"], d["c5_8M_1gpu"]["ms_per_step"])
print("ref port", r["ms_per_step"], "ref gpu", r["reference_gpu_solver"])
PY
bash scripts/gpu_scene_table.sh > /dev/null 2>&1; cp gpurun_out/scene_table.jsonl gpurun_out/${T}_scene_table.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --quick --steps 4 --warmup 3 > gpurun_out/${T}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
PS_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_find_lambdas|k_solve_fluids|k_radix_pass|k_calc_hash|k_reorder|k_cell_begin|k_collide_world|k_predict|k_velocity' -s 30 -c 14 -o gpurun_out/prof_${T} python bench.py --quick --steps 2 --warmup 3 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $CS --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream_io.py tests/test_gpu_slab.py -m gpu -q -k "wrap or cap_500 or cell or paths_agree or omega or streamed or prefetch or pack or lambda_exchange" > gpurun_out/${T}_sanitize_mem.log 2>&1; echo "sanitize mem rc=$? $(grep 'ERROR SUMMARY' gpurun_out/${T}_sanitize_mem.log | sort | uniq -c | tr '\n' ';')"; grep -E "passed|failed" gpurun_out/${T}_sanitize_mem.log | tail -1
timeout 900 $CS --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "paths_agree or cap_500" > gpurun_out/${T}_sanitize_race.log 2>&1; echo "sanitize race rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/${T}_sanitize_race.log | sort | uniq -c | tr '\n' ';')"
