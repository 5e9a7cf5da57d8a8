#!/bin/bash
# N GPUs (N = number visible): NCCL slab parity test, then the c5 bench at 64M, then c3 on one GPU for the record
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_slab_nccl.py -m gpu -x -q > gpurun_out/pytest_nccl_$N.log 2>&1; echo "pytest nccl rc=$?"
tail -5 gpurun_out/pytest_nccl_$N.log
for n in $N; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "bench ${n}gpu rc=$?"; cat gpurun_out/bench_${n}gpu.json; tail -3 gpurun_out/bench_${n}gpu.err | cut -c1-300
done
