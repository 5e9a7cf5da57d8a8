#!/bin/bash
# 2 GPUs: slab decomposition over NCCL (parity vs one GPU, both ghost-lambda modes), then the c5 weak-scaling point at
# 2 x 8M particles with the ghost lambdas exchanged (default) and computed locally (the previous scheme)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_slab_nccl.py -m gpu -x -q ) > gpurun_out/pytest_nccl_r18.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_nccl_r18.log
for mode in "" "--local-ghost-lambda"; do
  tag=${mode:+local}; tag=${tag:-exchange}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 10 --warmup 3 $mode \
    > gpurun_out/bench_2gpu_$tag.json 2> gpurun_out/bench_2gpu_$tag.err; echo "bench $tag rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_2gpu_$tag.json"))
print("$tag", "ms/step", round(d["ms_per_step"], 2), "value", f'{d["value"]:.3e}', "ghosts", d["config"]["ghosts_total"], "bytes/step", d["config"]["exchange_bytes_per_step"], d["stage_ms_per_step_rank0"])
PY
done
