#!/bin/bash
# Runs on the GPU box: smoke, GPU parity suite, golden-vector generation, short benches.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== golden"; timeout 300 python scripts/make_golden.py gpurun_out/golden 2>&1 | tail -10
for wl in 10k_grid 1m_random 16m_grid; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_${wl}.json
done
echo "== reference arm 1m"; timeout 600 python bench.py --impl reference --workload 1m_random --steps 20 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_ref_1m.json
