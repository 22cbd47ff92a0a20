#!/bin/bash
# builds gpurun_variants/lib_<name>.so (the whole library with extra -D flags) for variant timing through R3G_LIB=...
# VARIANTS: one "name flags..." per line
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden,-O3 --expt-relaxed-constexpr -rdc=false"
S=r3det-pytorch_b200/csrc
B=r3det-pytorch_b200/_build
while read -r name defs; do
  [ -z "$name" ] && continue
  ( nvcc $FLAGS $defs -c $S/${UNIT:-frm}.cu -o gpurun_variants/${UNIT:-frm}_$name.o && \
    objs=""; for u in api iou nms frm transforms coder; do if [ "$u" = "${UNIT:-frm}" ]; then objs="$objs gpurun_variants/${UNIT:-frm}_$name.o"; else objs="$objs $B/$u.cu.o"; fi; done; \
    nvcc -shared -o gpurun_variants/lib_$name.so $objs -lcudart ) &
done <<< "$VARIANTS"
wait
ls -la gpurun_variants/lib_*.so
