#!/bin/bash
# usage: gpurun --gpus N -- 'bash scripts/gpu_overlap.sh N'
# slab tests (single-process split + NCCL driver), then weak scaling without / with overlap + host trace
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3
run() { SPH_SLAB_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 100 --warmup 3 --no-cpu "$@" 2>&1 | grep -E 'host ms|^\{' ; }
for flag in "--no-overlap" ""; do
  echo "== weak N=$N $flag"; run $flag | tee "gpurun_out/overlap_n${N}${flag}.log" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.3e ms/step %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['load_balance'].get('speculative_hits'))
    else: print(l.strip()[:700])"
done
