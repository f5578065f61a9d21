#!/bin/bash
# Round-2 confirmation of the final code (counting sort by cell) on ONE B200: GPU tests, smoke(), the
# bench line at the driver's settings and over 100 steps, ncu launch list and --set full captures of
# one plainly launched step at the lattice state and at the pile-up state.
set -u
mkdir -p gpurun_out
T=r02c
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (driver settings: 20 steps after 5)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_16m_grid_20.json 2> gpurun_out/${T}_bench20.err; tail -c 400 gpurun_out/${T}_bench_16m_grid_20.json; echo
echo "== bench (100 steps)"; timeout 900 python bench.py --steps 100 --warmup 3 --no-cpu > gpurun_out/${T}_bench_16m_grid_100.json 2> gpurun_out/${T}_bench100.err; tail -c 300 gpurun_out/${T}_bench_16m_grid_100.json; echo
echo "== bench 1m_random"; timeout 900 python bench.py --workload 1m_random --steps 100 --no-cpu > gpurun_out/${T}_bench_1m_random.json 2> gpurun_out/${T}_bench_1m.err; tail -c 300 gpurun_out/${T}_bench_1m_random.json; echo
echo "== ncu launch list"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-morton"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_raw.csv $CMD > gpurun_out/${T}_ncu_launch.log 2>&1; tail -1 gpurun_out/${T}_ncu_launch.log | cut -c1-200
echo "== ncu full, early and late state"
for PRE in 3 100; do
  PCMD="python scripts/profile_step.py --workload 16m_grid --pre $PRE"
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${T}_full_step$PRE $PCMD > gpurun_out/${T}_ncu_full_$PRE.log 2>&1; tail -1 gpurun_out/${T}_ncu_full_$PRE.log
done
