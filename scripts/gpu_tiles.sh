#!/bin/bash
# A/B of SphOptions.stage_tiles (TMA-staged dense density CTAs): parity test, then stage timings
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -rf -p no:cacheprovider -k "tma_staged or mask_handoff" 2>&1 | tail -5
for flag in "" "--stage-tiles"; do
  for pre in 3 100; do
    echo "== 16m_grid pre=$pre $flag"; timeout 600 python scripts/profile_step.py --workload 16m_grid --pre $pre --steps 5 $flag 2>&1 | tail -7
  done
  echo "== 1m_random pre=100 $flag"; timeout 600 python scripts/profile_step.py --workload 1m_random --pre 100 --steps 5 $flag 2>&1 | tail -7
done
