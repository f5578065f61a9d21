#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for wl in 16m_grid 1m_random; do
  echo "== bench $wl"; timeout 900 python bench.py --workload $wl --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','stages','gpu_launches')}))" | tee gpurun_out/bench3_${wl}.json
done
echo "== profile plain"; timeout 600 python scripts/profile_step.py --workload 16m_grid --pre 100 2>&1 | tail -8
echo "== profile plain early"; timeout 600 python scripts/profile_step.py --workload 16m_grid --pre 3 2>&1 | tail -8
