#!/bin/bash
# usage: gpurun --gpus N -- 'bash scripts/gpu_overlap.sh N'   NCCL slab test, then weak scaling with / without overlap
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q --no-header -p no:cacheprovider -k nccl 2>&1 | tail -3
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 100 --warmup 3 --no-cpu "$@" 2>&1 | grep '^{' | tail -1; }
for flag in "" "--overlap"; do
  echo "== weak N=$N $flag"; run $flag | tee "gpurun_out/overlap_n${N}${flag}.json" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value %.3e ms/step %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
