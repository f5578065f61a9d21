#!/bin/bash
# usage: gpurun --gpus 2 -- 'bash scripts/gpu_slabtrace.sh 2'   host-side phase times of the slab protocol
set -u
N=${1:-2}
SPH_SLAB_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 100 --warmup 3 --no-cpu 2>&1 | grep -E "host ms|^\{" | cut -c1-900
