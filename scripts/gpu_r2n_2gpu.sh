#!/bin/bash
# r2n (2 GPUs): NCCL parity — slab.py over torch.distributed, ps_comm_step behind the C ABI, the C++ CLI over two ranks — log kept;
# then the N = 2 bench line driven by ps_comm_step and, for comparison, by slab.py
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_slab_nccl.py -m gpu -v -rA ) > gpurun_out/r2n_pytest_nccl.log 2>&1; echo "nccl pytest rc=$?"; grep -E "^(PASSED|FAILED|SKIPPED|ERROR)|passed|failed" gpurun_out/r2n_pytest_nccl.log | tail -14; grep -h "SLAB_NCCL_OK" gpurun_out/r2n_pytest_nccl.log | head
for host in c python; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --slab-host $host ) > gpurun_out/r2n_bench_2gpu_$host.json 2> gpurun_out/r2n_bench_2gpu_$host.err; echo "bench $host rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2n_bench_2gpu_$host.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"], d["state_check"]["particles_conserved"], d["config"]["exchange"][:40])
PY
tail -3 gpurun_out/r2n_bench_2gpu_$host.err
done
