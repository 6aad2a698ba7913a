#!/bin/bash
# Build an A/B variant of liblgcn_b200.so with extra -D flags for lgcn_spmm.cu (tuning only):
#   profiles/build_variant.sh NAME -DLGCN_RING_S=8 ...   ->  profiles/variants/NAME.so
# Run with LGCN_B200_LIB=profiles/variants/NAME.so python profiles/prof_spmm.py ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
src=${SRC:-lgcn_spmm}
mkdir -p profiles/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
  -Xcompiler -fvisibility=hidden -I include -I gcn_recommendation_b200/csrc "$@" \
  -c gcn_recommendation_b200/csrc/$src.cu -o profiles/variants/$name.o
objs=$(ls gcn_recommendation_b200/build/*.o | grep -v "/$src.o")
nvcc -shared -o profiles/variants/$name.so profiles/variants/$name.o $objs -cudart shared \
  -Xlinker -rpath=/usr/local/cuda/lib64
rm profiles/variants/$name.o
echo profiles/variants/$name.so
