#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -6
for wl in 16m_grid 1m_random; do
  timeout 900 python bench.py --workload $wl --no-cpu 2>&1 | tail -1 > gpurun_out/bench4_$wl.json; python -c "
import json
d=json.load(open('gpurun_out/bench4_$wl.json')); print('$wl', 'value %.3e ms/step %.3f | e2e %.3e (%.3f ms) blocking %.3e (%.3f ms) chk %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['blocking']['value'], d['e2e']['blocking']['ms_per_step'], d['e2e']['checksum_matches_blocking']))"
done
