#!/bin/bash
# r2zz (1 GPU): the whole GPU suite, smoke, the bench line and the scene table on the final build (after the fused 2-D tick and the lambda sinks)
mkdir -p gpurun_out
T=r2zz
( time timeout 1800 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${T}_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/${T}_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2zz_bench.json"))
print("ours", d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"], d["long_run"]["ms_per_step_mean"], d["c5_8M_1gpu"]["ms_per_step"])
PY
bash scripts/gpu_scene_table.sh > /dev/null 2>&1; cp gpurun_out/scene_table.jsonl gpurun_out/${T}_scene_table.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r2zz_scene_table.jsonl"):
    d=json.loads(l)
    print(d.get("impl", d.get("app")), d.get("scene"), d.get("particles", d.get("n")), d.get("ms_per_tick", d.get("wall_ms_per_tick", d.get("device_ms_per_step", d.get("ms_per_step")))))
PY
