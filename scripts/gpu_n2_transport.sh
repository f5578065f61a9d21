#!/bin/bash
# N=2: cluster tests (NCCL worker uses the default peer-memory transport), then weak and strong bench lines
set -u
echo "== cluster tests"; timeout 300 python -m pytest tests/test_gpu_cluster.py -q -x 2>&1 | tail -3
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500"
show() { tail -1 $1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
w = d.get('weak_scaling_reference') or {}
print('ms/step', round(d['ms_per_step'], 3), 'value %.3e' % d['value'], d['load_balance']['ms_per_step_per_rank'], 'n1', w.get('ms_per_step'), 'eff', w.get('efficiency'), 'parity', (d.get('parity_check') or {}).get('ok'))
" || tail -5 ${1%.json}.err; }
echo "== weak 32M/GPU"; timeout 300 $RUN bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_weak.json 2> gpurun_out/n2_weak.err; show gpurun_out/n2_weak.json
echo "== strong 16M total"; timeout 300 $RUN bench.py --gpus 2 --workload 16m_grid --scaling strong --total 16000000 --steps 50 --warmup 3 --no-parity --no-n1 > gpurun_out/n2_strong.json 2> gpurun_out/n2_strong.err; show gpurun_out/n2_strong.json
echo "== strong 16M total, NCCL data"; SPH_CLUSTER_NCCL_DATA=1 timeout 300 $RUN bench.py --gpus 2 --workload 16m_grid --scaling strong --total 16000000 --steps 50 --warmup 3 --no-parity --no-n1 > gpurun_out/n2_strong_nccl.json 2> gpurun_out/n2_strong_nccl.err; show gpurun_out/n2_strong_nccl.json
