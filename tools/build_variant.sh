#!/bin/bash
# usage: tools/build_variant.sh <tag> <file.cu> [-DNAME=VALUE ...]
# one .cu recompiled with extra defines, linked with the standard objects into build_ab/libskidgpu_<tag>.so
# (measurement scaffolding for tools/ab_variants.py; build_ab/ is git-ignored but travels to the GPU box)
set -e
TAG=$1; SRC=$2; shift 2
mkdir -p build_ab
BASE=$(basename $SRC .cu)
EXTRA=""; [ "$BASE" = groups ] && EXTRA="-fmad=false"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $EXTRA "$@" -c $SRC -o build_ab/${TAG}_$BASE.o
OBJS=$(ls skid_b200/csrc/*.o | grep -v "/$BASE.o")
/usr/local/cuda/bin/nvcc -shared -o build_ab/libskidgpu_$TAG.so $OBJS build_ab/${TAG}_$BASE.o -gencode arch=compute_100a,code=sm_100a
rm build_ab/${TAG}_$BASE.o
echo build_ab/libskidgpu_$TAG.so
