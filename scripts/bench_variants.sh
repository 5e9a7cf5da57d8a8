#!/bin/bash
# tuning aid: bench.py --quick for the default build and every particlesolver_b200/libpsolver_<variant>.so
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
python bench.py --quick --steps 10 --warmup 3 | tee -a gpurun_out/variants.jsonl
for lib in particlesolver_b200/libpsolver_*.so; do
  PS_LIBRARY=$PWD/$lib python bench.py --quick --steps 10 --warmup 3 | tee -a gpurun_out/variants.jsonl
done
