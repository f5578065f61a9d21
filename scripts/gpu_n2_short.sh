#!/bin/bash
# N=2: the two-GPU transport tests, then the default weak-scaling bench line (32M per GPU, parity check)
set -u
mkdir -p gpurun_out
echo "== two-GPU cluster tests"; timeout 300 python -m pytest tests/test_gpu_cluster.py -q -x -k "two_gpus" 2>&1 | tail -3
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500"
echo "== weak 32M/GPU"; timeout 400 $RUN bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_weak.json 2> gpurun_out/n2_weak.err
tail -1 gpurun_out/n2_weak.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
w = d.get('weak_scaling_reference') or {}
print('ms/step', round(d['ms_per_step'], 3), 'value %.3e' % d['value'], d['load_balance']['ms_per_step_per_rank'], 'n1', w.get('ms_per_step'), 'eff', w.get('efficiency'), 'parity', (d.get('parity_check') or {}).get('ok'), 'migrated', d['load_balance'].get('migrated_total'))
" || tail -5 gpurun_out/n2_weak.err
