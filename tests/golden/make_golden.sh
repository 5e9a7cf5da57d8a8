#!/bin/bash
# Regenerates tests/golden/ref_gpu_*.npz by running the reference's UNMODIFIED GPU solver (oracle/_ref/ref_gpu,
# built in the dev container by `make -C oracle ref`) on a B200.  Run from the repo root on a GPU box:
#   bash tests/golden/make_golden.sh  [raw_dir] [out_dir]
set -e
RAW=${1:-/tmp/ref_gpu_raw}
OUT=${2:-gpurun_out/golden}
REF=oracle/_ref/ref_gpu
rm -rf "$RAW"; mkdir -p "$RAW" "$OUT"
for s in ${SCENES:-1 2 3 4 5 6 7 8 9}; do
  $REF --scene $s --mode staged --steps 1 --out "$RAW/scene$s"
done
python tests/golden/pack_golden.py "$RAW" "$OUT"
