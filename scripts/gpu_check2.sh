#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== microbench"; (cd scripts/microbench && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu 2>/dev/null; timeout 120 ./pipes) 2>&1 | tee gpurun_out/microbench_pipes.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for wl in 16m_grid 1m_random; do
  echo "== bench $wl"; timeout 900 python bench.py --workload $wl 2>&1 | tail -2 | tee gpurun_out/bench100_${wl}.json
done
echo "== bench 16m morton"; timeout 600 python bench.py --workload 16m_grid --key morton --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench100_16m_morton.json
echo "== reference arm 16m"; timeout 900 python bench.py --impl reference --workload 16m_grid 2>&1 | tail -1 | tee gpurun_out/bench100_ref_16m.json
echo "== profile plain"; timeout 600 python scripts/profile_step.py --workload 16m_grid --pre 100 > gpurun_out/profile_plain.log 2>&1 && cat gpurun_out/profile_plain.log && \
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_16m_step100 python scripts/profile_step.py --workload 16m_grid --pre 100 > gpurun_out/ncu_16m.log 2>&1; tail -5 gpurun_out/ncu_16m.log
