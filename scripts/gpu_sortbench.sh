#!/bin/bash
cd scripts/microbench
for cfg in "3 4 0" "3 4 1" "4 4 0" "4 4 1" "5 4 1"; do set -- $cfg
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DSORT_CTAS_PER_SM=$1 -DSORT_LOOKBACK_WINDOW=$2 -DSORT_USE_MATCH_INSTRUCTION=$3 -o sort_bench_v sort_bench.cu 2>/dev/null && echo "ctas=$1 window=$2 match=$3: $(./sort_bench_v | cut -c1-90) | $(./sort_bench_v random | cut -c12-40)"; done
