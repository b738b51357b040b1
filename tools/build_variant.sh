#!/bin/bash
# Build an A-B variant of the library with extra compile-time flags:  tools/build_variant.sh NAME -DQB_RES_EPI_WARPS=8 ...
# -> qinco_b200/variants/NAME.so (git-ignored, travels with gpurun); load it with QINCO_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../qinco_b200"
name=$1; shift
mkdir -p variants /tmp/qbv_$name
objs=""
for s in qb_api.cu qb_mlp.cu qb_kernels.cu qb_prep_tc.cu qb_ivf_tc.cu qb_pairwise.cu qb_plan.cpp; do
  o=/tmp/qbv_$name/${s%.*}.o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 "$@" -c csrc/$s -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/$name.so $objs -cudart static
ls -la variants/$name.so
