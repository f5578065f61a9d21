#!/bin/bash
# quick correctness + stage timing loop used while tuning kernels
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== early"; timeout 600 python scripts/profile_step.py --workload 16m_grid --pre 3 --steps 5 2>&1 | tail -7
echo "== step100"; timeout 600 python scripts/profile_step.py --workload 16m_grid --pre 100 --steps 5 2>&1 | tail -7
echo "== 1m step100"; timeout 600 python scripts/profile_step.py --workload 1m_random --pre 100 --steps 5 2>&1 | tail -7
for wl in 16m_grid 1m_random; do
  timeout 900 python bench.py --workload $wl --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl', 'value %.3e ms/step %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
