#!/bin/bash
cd scripts/microbench
for cfg in "3 16 0" "3 16 100" "3 16 400" "3 8 200" "3 4 200" "3 1 200" "4 8 200"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DSORT_CTAS_PER_SM=$1 -DSORT_LOOKBACK_WINDOW=$2 -DSORT_SPIN_SLEEP_NS=$3 -o sort_bench_v sort_bench.cu 2>/dev/null && echo "ctas=$1 window=$2 sleep=$3: $(timeout 120 ./sort_bench_v | cut -c1-80)"
done
