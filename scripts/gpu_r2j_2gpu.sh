#!/bin/bash
# r2j (2 GPUs): NCCL parity tests with the log kept, then the weak-scaling bench line at N = 2 (8M particles per GPU)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_slab_nccl.py tests/test_gpu_slab.py -m gpu -v -rA ) > gpurun_out/r2j_pytest_nccl.log 2>&1; echo "nccl pytest rc=$?"; grep -E "PASSED|FAILED|SKIPPED|passed|failed" gpurun_out/r2j_pytest_nccl.log | tail -20
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; echo "bench rc=$?"; cut -c1-3500 gpurun_out/r2j_bench_2gpu.json; tail -5 gpurun_out/r2j_bench_2gpu.err
