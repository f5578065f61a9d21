#!/bin/bash
# ab_libs.sh "<ab_stages args>" lib1 lib2 ...   (names under build/ab/, or "main" for the in-tree library)
ARGS=$1; shift
for L in "$@"; do
  echo "== $L"
  if [ "$L" = main ]; then unset SPH_B200_LIB; else export SPH_B200_LIB=$PWD/build/ab/lib_$L.so; fi
  python scripts/ab_stages.py $ARGS 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception:
        print(l.rstrip()); continue
    pts = {k: (v.get("sort_passes"), v.get("reorder_cellstart"), v["density"], v["force_integrate"]) for k, v in d["points"].items()}
    print(d["variant"], "(sort, reorder, density, force):", pts, "advance", d["advance_ms_per_step"], "KE", round(d["ke_after"], 1))
'
done
