#!/bin/bash
# usage: gpurun --gpus 8 -- 'bash scripts/gpu_weak16_n8.sh'   the driver's weak-scaling point at N=8 (16M/GPU)
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 100 --warmup 3 --no-cpu 2>&1 | grep '^{' | tail -1 | tee gpurun_out/weak16_n8.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=8 value %.3e ms/step %.3f e2e %.3e n_total %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['n_total']))"
