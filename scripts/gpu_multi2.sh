#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "== nccl slab test"; timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q --no-header -rf -p no:cacheprovider 2>&1 | tail -4
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $1 "${@:2}" 2>&1 | grep '^{' | tail -1; }
echo "== weak N=$N"; run $N --steps 100 --warmup 3 | tee gpurun_out/scale_weak_n$N.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value %.3e ms/step %.3f e2e %.3e n_total %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['n_total']))"
