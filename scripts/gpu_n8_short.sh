set -u
N=8; T=r02b
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
echo "== weak, driver settings"; timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_weak32_n${N}_20.json 2> gpurun_out/${T}_weak20.err; tail -c 300 gpurun_out/${T}_bench_weak32_n${N}_20.json; echo
echo "== strong, 64 M"; timeout 300 $RUN bench.py --gpus $N --workload 16m_grid --scaling strong --total 64000000 --steps 50 --warmup 3 --no-parity --no-n1 > gpurun_out/${T}_bench_strong64_n${N}.json 2> gpurun_out/${T}_strong.err; tail -c 200 gpurun_out/${T}_bench_strong64_n${N}.json; echo
