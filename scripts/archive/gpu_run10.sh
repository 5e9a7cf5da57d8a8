#!/bin/bash
# per-kernel time + DRAM bytes at 8M particles on one GPU (c5 dam break): the streaming kernels against the HBM roofline at scale
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 400 -c 120 --csv --log-file gpurun_out/launches_c5_8m.csv python bench.py --workload c5 --particles 8000000 --steps 3 --warmup 3 > gpurun_out/ncu_c5.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_c5.log | cut -c1-300
