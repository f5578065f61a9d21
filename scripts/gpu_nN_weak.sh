#!/bin/bash
# scripts/gpu_nN_weak.sh <N>: the default weak-scaling bench line on N GPUs (32M per GPU, parity check)
set -u
N=$1
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_weak.json 2> gpurun_out/n${N}_weak.err
tail -1 gpurun_out/n${N}_weak.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
w = d.get('weak_scaling_reference') or {}
print('ms/step', round(d['ms_per_step'], 3), 'value %.3e' % d['value'], d['load_balance']['ms_per_step_per_rank'], 'n1', w.get('ms_per_step'), 'eff', w.get('efficiency'), 'parity', (d.get('parity_check') or {}).get('ok'), 'migrated', d['load_balance'].get('migrated_total'))
" || tail -5 gpurun_out/n${N}_weak.err
