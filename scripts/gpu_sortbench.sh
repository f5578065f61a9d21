#!/bin/bash
cd scripts/microbench
for e in 0 1 2 3 4 7; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DSORT_EXP=$e -o sort_bench_$e sort_bench.cu 2>/dev/null && timeout 120 ./sort_bench_$e && { [ $e = 0 ] && timeout 120 ./sort_bench_$e random; }
done
