#!/bin/bash
# round-end confirmation of the committed state: GPU tests, smoke(), both bench arms, 1M workload
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --durations=5 2>&1 | tail -14 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_16m_grid.json 2> gpurun_out/bench_reference.err; tail -c 700 gpurun_out/bench_reference_16m_grid.json; echo
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_16m_grid.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_16m_grid.json; echo
echo "== bench 1m_random"; timeout 900 python bench.py --workload 1m_random > gpurun_out/bench_1m_random.json 2> gpurun_out/bench_1m.err; tail -c 400 gpurun_out/bench_1m_random.json; echo
echo "== bench 10k_grid"; timeout 900 python bench.py --workload 10k_grid --no-cpu > gpurun_out/bench_10k_grid.json 2> gpurun_out/bench_10k.err; tail -c 300 gpurun_out/bench_10k_grid.json; echo
