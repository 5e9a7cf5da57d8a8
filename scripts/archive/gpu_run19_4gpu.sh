#!/bin/bash
# 4 GPUs: NCCL slab parity (all cases incl. 4 ranks) and the c5 weak-scaling point at 4 x 8M particles with the ghost-lambda exchange
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_slab_nccl.py -m gpu -x -q ) > gpurun_out/pytest_nccl_r19.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_nccl_r19.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29813 bench.py --gpus 4 --steps 10 --warmup 3 \
  > gpurun_out/bench_4gpu_exchange.json 2> gpurun_out/bench_4gpu_exchange.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_4gpu_exchange.json"))
print("ms/step", round(d["ms_per_step"], 2), "value", f'{d["value"]:.3e}', "ghosts", d["config"]["ghosts_total"], "bytes/step", d["config"]["exchange_bytes_per_step"], d["stage_ms_per_step_rank0"])
PY
