#!/bin/bash
# usage: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
echo "== nccl slab test"; timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q --no-header -rf -p no:cacheprovider -k nccl 2>&1 | tail -8
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 50 --warmup 3 2>&1 | tail -4 | tee gpurun_out/bench_slab_n$N.json
