#!/bin/bash
# r2g: why did the reference's CUDA binary take 299 ms/step inside bench.py --impl reference (r2f) and 23 ms in round 1's scene table?
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used,memory.total,ecc.mode.current --format=csv
echo "--- as the scene table ran it (30 steps)"
REF_GPU_VERBOSE=1 timeout 300 oracle/_ref/ref_gpu --scene c3 --mode whole --grid 256 --max 1001024 --side 100 --steps 30 --dump-every 0 --out /tmp/rg1 2>&1 | tail -34
echo "--- as bench.py runs it"
REF_GPU_VERBOSE=1 OMP_NUM_THREADS=16 timeout 300 oracle/_ref/ref_gpu --scene c3 --grid 256 --side 100 --max 1004096 --iters 5 --mode whole --steps 23 --warmup 3 --out /tmp/rg2 2>&1 | tail -27
