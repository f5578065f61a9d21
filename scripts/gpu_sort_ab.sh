#!/bin/bash
# A/B of the step's sort on one GPU: counting sort by cell (fused / standalone count) against the radix passes.
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for mode in "SPH_FUSE_COUNT=1" "SPH_FUSE_COUNT=0" "SPH_SORT=radix"; do
  env $mode timeout 200 python scripts/ab_stages.py --at 3,100 2>&1 | tail -1 | cut -c1-900
done
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/sort_launches.csv python scripts/ab_stages.py --at 3 --total 4 > gpurun_out/sort_ncu.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/sort_launches.csv') if l.startswith('"'))]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
acc = collections.defaultdict(list)
for r in rows[1:]:
    try: acc[r[ki][:60]].append(float(r[vi].replace(',', '')))
    except ValueError: pass
for k, v in acc.items(): print(f"{k:60s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.2f} us  last={v[-1]/1e3:9.2f} us")
P
