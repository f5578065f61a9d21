#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference.py -m gpu -q --no-header -p no:cacheprovider -k reconstructed 2>&1 | tail -5
timeout 900 python bench.py --workload 1m_random --compare-variants --steps 100 2>&1 | tail -1 | tee gpurun_out/variants_1m_random.json
timeout 900 python bench.py --workload 10k_grid --compare-variants --steps 100 2>&1 | tail -1 | tee gpurun_out/variants_10k_grid.json
