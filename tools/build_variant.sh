#!/bin/bash
# Builds a variant of the CUDA library with extra nvcc defines into variants/<name>/libmovierec_b200.so
# (A/B runs on the GPU box: MR_LIB_PATH=variants/<name>/libmovierec_b200.so python bench.py ...).
# usage: tools/build_variant.sh <name> "<-D... flags>"
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/movierecommender-tf-trt_b200/csrc
out=$root/variants/$name
mkdir -p "$out/obj"
for f in api neumf_kernels gather radix_sort segreduce optimizer dp_exchange rank sampler dataset tc_dense tc_wgrad tc_fused small_tower head; do
  case $f in tc_dense|tc_wgrad|tc_fused|small_tower|head) ;; *) if [ -f "$src/build/$f.o" ]; then cp "$src/build/$f.o" "$out/obj/$f.o"; continue; fi;; esac
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I"$root/include" -I"$src" $flags -c "$src/$f.cu" -o "$out/obj/$f.o" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libmovierec_b200.so" "$out"/obj/*.o -lcudart
rm -rf "$out/obj"
echo "built $out/libmovierec_b200.so"
