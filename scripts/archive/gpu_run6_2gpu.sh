#!/bin/bash
# 2 GPUs: slab decomposition over NCCL — parity test, then the c5 bench at 64M
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_slab_nccl.py -m gpu -x -q > gpurun_out/pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"
tail -15 gpurun_out/pytest_nccl.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu rc=$?"; cat gpurun_out/bench_2gpu.json; tail -8 gpurun_out/bench_2gpu.err
