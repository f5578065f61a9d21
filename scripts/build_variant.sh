#!/bin/bash
# build_variant.sh <name> [-DFOO=1 ...]  ->  build/ab/lib_<name>.so  (same ABI; select with SPH_B200_LIB)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
mkdir -p "$ROOT/build/ab"
cd "$ROOT/cudafluidsimulator_b200/csrc"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --extended-lambda \
     -Xcompiler -fPIC -shared "$@" -I "$ROOT/include" -o "$ROOT/build/ab/lib_$NAME.so" \
     sph_api.cu sph_kernels.cu sph_sort.cu sph_cluster.cu -ldl -lpthread
