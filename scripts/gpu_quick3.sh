#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -6
for wl in 16m_grid 1m_random; do
  timeout 900 python bench.py --workload $wl --no-cpu 2>&1 | tail -1 > gpurun_out/bench5_$wl.json; python -c "
import json
d=json.load(open('gpurun_out/bench5_$wl.json')); print('$wl', 'value %.3e ms/step %.3f | e2e %.3e (%.3f ms) blocking %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['blocking']['value'])); print({k:v['ms'] for k,v in d['stages'].items()})"
done
timeout 300 python bench.py --workload 10k_grid --compare-variants --steps 100 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d['variants'].items(): print('%-70s grid %.3f upd %.3f xfer %.3f' % (k, v['grid_ms'], v['sph_update_ms'], v['transfer_ms']))"
./cudafluidsimulator_b200/sph -n 10000 -i grid -m time
