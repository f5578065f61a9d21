#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full capture of one step
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/bench_plain.log | head -c 600; echo
PCMD="python scripts/profile_step.py --workload 16m_grid --pre 100"
$PCMD > gpurun_out/profile_plain.log 2>&1 && tail -7 gpurun_out/profile_plain.log && \
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_16m_step100_final $PCMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
