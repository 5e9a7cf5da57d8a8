#!/bin/bash
# r2 scaling: bench line at N GPUs (8M particles per GPU, ps_comm_step), plus — at N = 4 — the 4-rank NCCL parity tests
N=$1
mkdir -p gpurun_out
if [ "$N" = "4" ]; then
  ( time timeout 1200 python -m pytest tests/test_gpu_slab_nccl.py -m gpu -v -rA ) > gpurun_out/r2z_pytest_nccl_${N}gpu.log 2>&1; echo "nccl pytest rc=$?"; grep -E "^(PASSED|FAILED|SKIPPED|ERROR)|passed|failed" gpurun_out/r2z_pytest_nccl_${N}gpu.log | tail -12
fi
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/r2z_bench_${N}gpu.json 2> gpurun_out/r2z_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2z_bench_${N}gpu.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["ms_per_step"], d["state_check"]["particles_conserved"], d["stage_ms_per_step_rank0"])
PY
tail -3 gpurun_out/r2z_bench_${N}gpu.err
