set -u
echo "== cluster tests"; timeout 300 python -m pytest tests/test_gpu_cluster.py -q -x 2>&1 | tail -4
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500"
for MODE in 0 1; do
echo "== bench N=2 SPH_CLUSTER_NCCL_DATA=$MODE"
SPH_CLUSTER_NCCL_DATA=$MODE timeout 300 $RUN bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_$MODE.json 2> gpurun_out/n2_$MODE.err
tail -1 gpurun_out/n2_$MODE.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'], 3), 'value %.3e' % d['value'], 'e2e ms', round(d['e2e']['ms_per_step'], 3), d['load_balance']['migrated_total'], d['load_balance']['ms_per_step_per_rank'], 'n1', round(d['weak_scaling_reference']['ms_per_step'], 3), 'eff', round(d['weak_scaling_reference']['efficiency'], 4), 'parity', d['parity_check']['ok'])
" || tail -5 gpurun_out/n2_$MODE.err
done
