#!/bin/bash
# per-kernel durations (ncu, serialised) of 4 steps for library variants: scripts/gpu_sort_ncu.sh main pad ...
mkdir -p gpurun_out
for L in "$@"; do
  echo "== $L"
  if [ "$L" = main ]; then unset SPH_B200_LIB; else export SPH_B200_LIB=$PWD/build/ab/lib_$L.so; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/sort_launches_$L.csv python scripts/ab_stages.py --at 3 --total 4 > gpurun_out/sort_ncu_$L.log 2>&1
  python - "$L" <<'P'
import csv, collections, sys
rows = [r for r in csv.reader(l for l in open(f'gpurun_out/sort_launches_{sys.argv[1]}.csv') if l.startswith('"'))]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
acc = collections.defaultdict(list)
for r in rows[1:]:
    try: acc[r[ki].split('(')[0][-40:]].append(float(r[vi].replace(',', '')))
    except ValueError: pass
for k, v in acc.items(): print(f"{k:42s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.2f} us  min={min(v)/1e3:9.2f} us")
P
done
