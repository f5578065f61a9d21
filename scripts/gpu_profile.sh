#!/bin/bash
# ncu evidence for profiles/ (one ncu pass per gpurun call):
#   gpu_profile.sh launches  -> launch list of the bench command
#   gpu_profile.sh full      -> --set full capture of one plainly launched step (developed state)
set -u
mkdir -p gpurun_out
case "${1:-launches}" in
launches)
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
  $CMD > gpurun_out/bench_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
  tail -2 gpurun_out/bench_plain.log | head -c 600; echo ;;
full)
  PCMD="python scripts/profile_step.py --workload 16m_grid --pre 100"
  $PCMD > gpurun_out/profile_plain.log 2>&1 && tail -7 gpurun_out/profile_plain.log && \
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_16m_step100_final $PCMD > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log ;;
esac
