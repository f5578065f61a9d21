#!/bin/bash
# Round-2 confirmation of the committed state on ONE B200: GPU tests, smoke(), both bench arms, the
# other workloads, the variants comparison, and the ncu evidence (launch list + --set full captures
# of one plainly launched step at the lattice state and at the pile-up state).
set -u
mkdir -p gpurun_out
T=r02
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --durations=5 2>&1 | tail -12 | tee gpurun_out/${T}_pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (driver settings: 20 steps after 5)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_16m_grid_20.json 2> gpurun_out/${T}_bench20.err; tail -c 600 gpurun_out/${T}_bench_16m_grid_20.json; echo
echo "== bench (100 steps)"; timeout 900 python bench.py --steps 100 --warmup 3 > gpurun_out/${T}_bench_16m_grid.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench_16m_grid.json; echo
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference_16m_grid.json 2> gpurun_out/${T}_bench_reference.err; tail -c 500 gpurun_out/${T}_bench_reference_16m_grid.json; echo
echo "== bench 1m_random"; timeout 900 python bench.py --workload 1m_random --steps 100 > gpurun_out/${T}_bench_1m_random.json 2> gpurun_out/${T}_bench_1m.err; tail -c 300 gpurun_out/${T}_bench_1m_random.json; echo
echo "== bench 10k_grid"; timeout 900 python bench.py --workload 10k_grid --steps 100 --no-cpu > gpurun_out/${T}_bench_10k_grid.json 2> gpurun_out/${T}_bench_10k.err; tail -c 300 gpurun_out/${T}_bench_10k_grid.json; echo
echo "== bench 32m_grid"; timeout 900 python bench.py --workload 32m_grid --steps 100 --no-cpu --no-morton > gpurun_out/${T}_bench_32m_grid.json 2> gpurun_out/${T}_bench_32m.err; tail -c 300 gpurun_out/${T}_bench_32m_grid.json; echo
echo "== variants (config 2)"; timeout 900 python bench.py --compare-variants --workload 1m_random --steps 100 > gpurun_out/${T}_variants_1m.json 2> gpurun_out/${T}_variants.err; tail -c 1500 gpurun_out/${T}_variants_1m.json; echo
echo "== flat vs morton stages"; for K in flat morton; do python scripts/ab_stages.py --key $K --at 3,100 --counts > gpurun_out/${T}_stages_$K.json 2>&1; tail -c 900 gpurun_out/${T}_stages_$K.json; echo; done
echo "== ncu launch list"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-morton"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_raw.csv $CMD > gpurun_out/${T}_ncu_launch.log 2>&1; tail -1 gpurun_out/${T}_ncu_launch.log | cut -c1-200
echo "== ncu full, early and late state"
for PRE in 3 100; do
  PCMD="python scripts/profile_step.py --workload 16m_grid --pre $PRE"
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${T}_full_step$PRE $PCMD > gpurun_out/${T}_ncu_full_$PRE.log 2>&1; tail -1 gpurun_out/${T}_ncu_full_$PRE.log
done
for K in morton; do
  PCMD="python scripts/profile_step.py --workload 16m_grid --pre 3 --key morton"
  ncu --set full --clock-control none --profile-from-start off -k regex:'k_density|k_force' -f -o gpurun_out/${T}_morton_step3 $PCMD > gpurun_out/${T}_ncu_morton.log 2>&1; tail -1 gpurun_out/${T}_ncu_morton.log
done
