#!/bin/bash
# usage: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N [extra bench args]'
set -u
N=${1:-2}; shift
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $1 "${@:2}" 2>&1 | grep '^{' | tail -1; }
for g in 2 4 8; do
  [ $g -le $N ] || continue
  echo "== weak N=$g"; run $g --steps 100 --warmup 3 | tee gpurun_out/scale_weak_n$g.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value %.3e ms/step %.3f e2e %.3e n_total %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['n_total']), d['load_balance'])"
done
for g in 2 4 8; do
  [ $g -le $N ] || continue
  echo "== strong 64M N=$g"; run $g --steps 50 --warmup 3 --scaling strong | tee gpurun_out/scale_strong_n$g.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value %.3e ms/step %.3f n_total %d' % (d['value'], d['ms_per_step'], d['config']['n_total']), d['load_balance']['owned_per_rank'])"
done
