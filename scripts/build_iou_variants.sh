#!/bin/bash
# builds gpurun_variants/iou_<name>.so for scripts/time_variants.py: each line of VARIANTS = name + nvcc -D flags
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
rm -f gpurun_variants/iou_*.so
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden,-O3 --expt-relaxed-constexpr -rdc=false -shared"
while read -r name defs; do
  [ -z "$name" ] && continue
  nvcc $FLAGS $defs r3det-pytorch_b200/csrc/api.cu r3det-pytorch_b200/csrc/iou.cu -o gpurun_variants/iou_$name.so -lcudart &
done <<< "$VARIANTS"
wait
ls -la gpurun_variants/iou_*.so
