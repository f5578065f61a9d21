#!/bin/bash
# north_star weak-scaling point: 32M particles per GPU.
# usage: gpurun --gpus 8 -- 'bash scripts/gpu_weak32.sh 8'   (N=1 reference point: gpu_weak32.sh 1)
set -u
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" -eq 1 ]; then
  timeout 900 python bench.py --workload 32m_grid --steps 100 --warmup 3 --no-cpu 2>&1 | grep '^{' | tail -1 | tee gpurun_out/weak32_n1.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=1 value %.3e ms/step %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --workload 32m_grid --steps 100 --warmup 3 --no-cpu 2>&1 | grep '^{' | tail -1 | tee gpurun_out/weak32_n$N.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=$N value %.3e ms/step %.3f e2e %.3e n_total %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['n_total']), d['load_balance'])"
fi
