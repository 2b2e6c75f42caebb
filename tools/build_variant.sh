#!/bin/bash
# usage: tools/build_variant.sh <name> [-DFLAG=V ...]  -> vampire_b200/_lib/variants/libvb200_<name>.so
# Builds the whole library with extra nvcc defines, for A/B runs on the GPU box:
#   VB200_LIB=$PWD/vampire_b200/_lib/variants/libvb200_<name>.so python bench.py ...
set -e
name=$1; shift
out=vampire_b200/_lib/variants; mkdir -p $out/obj_$name
for f in vampire_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -I include -c $f -o $out/obj_$name/$(basename $f .cu).o &
done
wait
nvcc -shared -o $out/libvb200_$name.so $out/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf $out/obj_$name
echo $out/libvb200_$name.so
