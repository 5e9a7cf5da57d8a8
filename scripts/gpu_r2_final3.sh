#!/bin/bash
# r2zz final (1 GPU): whole GPU suite, smoke, bench line, scene-6 replicas on the last build
mkdir -p gpurun_out
T=r2zz
( time timeout 1800 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${T}_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/${T}_pytest_all.log
rm -f gpurun_out/${T}_ps2d_parity.txt; PS2D_REPORT=gpurun_out/${T}_ps2d_parity.txt timeout 600 python -m pytest tests/test_gpu_2d_full.py -m gpu -q > /dev/null 2>&1; echo "parity report rc=$?"; cat gpurun_out/${T}_ps2d_parity.txt | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2zz_bench.json"))
print("ours", d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"], d["long_run"]["ms_per_step_mean"], d["c5_8M_1gpu"]["ms_per_step"], d["c1_2d_path"])
PY
OUT=gpurun_out/${T}_scene6_replicas.jsonl
: > $OUT
CLI=particlesolver_b200/psolver_cli
for spec in "4 6 200" "8 6x2 60" "16 6x4 12"; do
  set -- $spec
  timeout 600 oracle/_ref/ref_cpu --scene 6 --fluid-scale $1 --ticks $3 --json | grep '^{' >> $OUT
  timeout 600 $CLI --app cpu --scene $2 --ticks $3 --json >> $OUT
done
python - <<'PY'
import json
for l in open("gpurun_out/r2zz_scene6_replicas.jsonl"):
    d=json.loads(l); print(d.get("impl","ours"), d.get("n", d.get("particles")), d.get("ms_per_tick", d.get("wall_ms_per_tick")), d.get("launches_per_tick"))
PY
