#!/bin/bash
# ncu --set full of the two neighbour kernels, one plainly launched step at an early (lattice)
# and a late (floor pile-up) state of the 16M grid workload.  Usage: gpu_ncu_nb.sh <tag> [extra profile_step args]
set -u
TAG=${1:-r02}; shift || true
mkdir -p gpurun_out
for PRE in 3 100; do
  PCMD="python scripts/profile_step.py --workload 16m_grid --pre $PRE $*"
  $PCMD > gpurun_out/${TAG}_plain_$PRE.log 2>&1 && tail -6 gpurun_out/${TAG}_plain_$PRE.log && \
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_density|k_force' \
      -f -o gpurun_out/${TAG}_nb_step$PRE $PCMD > gpurun_out/${TAG}_ncu_$PRE.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_$PRE.log
done
