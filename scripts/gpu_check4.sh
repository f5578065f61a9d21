#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x --durations=8 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== profile plain"; timeout 600 python scripts/profile_step.py --workload 16m_grid --pre 100 > gpurun_out/profile_plain.log 2>&1 && tail -7 gpurun_out/profile_plain.log && \
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_density|k_force' -f -o gpurun_out/prof_16m_step100_v3 python scripts/profile_step.py --workload 16m_grid --pre 100 > gpurun_out/ncu_16m.log 2>&1; tail -3 gpurun_out/ncu_16m.log
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_density|k_force' -f -o gpurun_out/prof_16m_step3_v3 python scripts/profile_step.py --workload 16m_grid --pre 3 > gpurun_out/ncu_16m_early.log 2>&1; tail -3 gpurun_out/ncu_16m_early.log
