#!/bin/bash
# compute-sanitizer over small representative runs: memcheck (out-of-bounds / misaligned) and racecheck (shared-memory hazards)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; shift; timeout 900 $CS "$@" > gpurun_out/sanitize_$name.log 2>&1; echo "$name rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_$name.log) summaries: $(grep 'ERROR SUMMARY' gpurun_out/sanitize_$name.log | sort | uniq -c | tr '\n' ';')"; }
run mem3d --tool memcheck --print-limit 5 particlesolver_b200/psolver_cli --app gpu --scene 8 --steps 3
run mem3d_fluid --tool memcheck --print-limit 5 particlesolver_b200/psolver_cli --app gpu --scene 3 --steps 2
run race3d_fluid --tool racecheck --print-limit 5 particlesolver_b200/psolver_cli --app gpu --scene 7 --steps 1
run race3d_combo --tool racecheck --print-limit 5 particlesolver_b200/psolver_cli --app gpu --scene 8 --steps 1
run mem2d_w --tool memcheck --print-limit 5 particlesolver_b200/psolver_cli --app cpu --scene w --ticks 30
run mem2d_v --tool memcheck --print-limit 5 particlesolver_b200/psolver_cli --app cpu --scene v --ticks 30
run race2d_w --tool racecheck --print-limit 5 particlesolver_b200/psolver_cli --app cpu --scene w --ticks 5
run race2d_3 --tool racecheck --print-limit 5 particlesolver_b200/psolver_cli --app cpu --scene 3 --ticks 5
run sync2d_w --tool synccheck --print-limit 5 particlesolver_b200/psolver_cli --app cpu --scene w --ticks 5
run sync3d --tool synccheck --print-limit 5 particlesolver_b200/psolver_cli --app gpu --scene 8 --steps 1
for f in gpurun_out/sanitize_*.log; do echo "== $f"; grep -v "^=========$" $f | grep "=========" | head -12; done
