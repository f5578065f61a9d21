#!/bin/bash
# stage timings only (no pytest): early and developed 16M state, developed 1M state
set -u
mkdir -p gpurun_out
echo "== early"; timeout 600 python scripts/profile_step.py --workload 16m_grid --pre 3 --steps 5 "$@" 2>&1 | tail -7
echo "== step100"; timeout 600 python scripts/profile_step.py --workload 16m_grid --pre 100 --steps 5 "$@" 2>&1 | tail -7
echo "== 1m step100"; timeout 600 python scripts/profile_step.py --workload 1m_random --pre 100 --steps 5 "$@" 2>&1 | tail -7
