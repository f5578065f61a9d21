#!/bin/bash
# Round-2 multi-GPU evidence on N GPUs of one box (gpurun --gpus N): weak scaling at 32 M particles per
# GPU (BASELINE configs[4]), strong scaling of a 64 M dam-break slab (configs[3]), the C++ host with
# -g N (one process, peer-to-peer copies) and the D2H ceiling with all GPUs copying.
set -u
N=${1:-8}
T=r02
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
echo "== weak, driver settings"; timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 5 --timeline > gpurun_out/${T}_bench_weak32_n${N}_20.json 2> gpurun_out/${T}_weak20.err; tail -c 1200 gpurun_out/${T}_bench_weak32_n${N}_20.json; echo
echo "== weak, 100 steps"; timeout 400 $RUN bench.py --gpus $N --steps 100 --warmup 3 --no-parity --timeline > gpurun_out/${T}_bench_weak32_n${N}.json 2> gpurun_out/${T}_weak.err; tail -c 1200 gpurun_out/${T}_bench_weak32_n${N}.json; echo
echo "== strong, 64 M"; timeout 400 $RUN bench.py --gpus $N --workload 16m_grid --scaling strong --total 64000000 --steps 50 --warmup 3 --no-parity --no-n1 > gpurun_out/${T}_bench_strong64_n${N}.json 2> gpurun_out/${T}_strong.err; tail -c 600 gpurun_out/${T}_bench_strong64_n${N}.json; echo
echo "== ./sph -g $N (one process, P2P)"; timeout 300 ./cudafluidsimulator_b200/sph -n 16000000 -b 25.6 -c 256 -g $N -m time -s 20 2>&1 | tail -6 | tee gpurun_out/${T}_sph_g${N}.txt
echo "== D2H ceiling"; timeout 200 $RUN scripts/pcie_d2h.py 2>&1 | tail -1 | tee gpurun_out/${T}_pcie_d2h_n${N}.txt
