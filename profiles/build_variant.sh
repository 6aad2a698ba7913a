#!/bin/bash
# Build an A/B variant of liblgcn_b200.so with extra -D flags (tuning only):
#   profiles/build_variant.sh NAME -DLGCN_RING_S=8 ...   ->  profiles/variants/NAME.so
# By default the flags go to the SpMM kernels (the five per-width translation units
# lgcn_spmm_d*.cu); SRC=lgcn_fusion_tc (any other single source) rebuilds that file instead.
# Run with LGCN_B200_LIB=profiles/variants/NAME.so python profiles/prof_spmm.py ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
if [ -n "$SRC" ]; then srcs="$SRC"; else srcs="lgcn_spmm_d16 lgcn_spmm_d32 lgcn_spmm_d64 lgcn_spmm_d128 lgcn_spmm_d256"; fi
mkdir -p profiles/variants
objs=""; skip=""
for s in $srcs; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
    -Xcompiler -fvisibility=hidden -I include -I gcn_recommendation_b200/csrc "$@" \
    -c gcn_recommendation_b200/csrc/$s.cu -o profiles/variants/$name.$s.o &
  objs="$objs profiles/variants/$name.$s.o"; skip="$skip -e /$s.o"
done
wait
rest=$(ls gcn_recommendation_b200/build/*.o | grep -v $skip)
nvcc -shared -o profiles/variants/$name.so $objs $rest -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64
rm $objs
echo profiles/variants/$name.so
