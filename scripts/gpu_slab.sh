#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -30 | tee gpurun_out/pytest_slab.log
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x --deselect tests/test_gpu_slab.py 2>&1 | tail -3
