#!/usr/bin/env python
"""Per-stage times of the step at several points of a run, for A/B comparisons on the GPU box.

    python scripts/ab_stages.py [--workload 16m_grid] [--at 3,50,100] [--variants exact,factored]

For every variant: set up the workload, advance to each of the `--at` step counts (graph replay),
profile `--prof` plainly launched steps there (CUDA events around every kernel) and print the
stage times; finally the device time of a whole `advance(--total)` from the initial state.
Prints one JSON object per variant.
"""
import argparse
import ctypes
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import cudafluidsimulator_b200 as sph  # noqa: E402
from bench import WORKLOADS  # noqa: E402

VARIANTS = {
    "default": {},
    "exact": dict(density_sum=1),
    "factored": dict(density_sum=2),
    "tiles": dict(stage_tiles=True),
    "nomask": dict(mask_handoff=False),
}

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="16m_grid", choices=list(WORKLOADS))
ap.add_argument("--at", default="3,50,100")
ap.add_argument("--prof", type=int, default=3)
ap.add_argument("--total", type=int, default=100)
ap.add_argument("--variants", default="default")
ap.add_argument("--key", default="flat")
ap.add_argument("--counts", action="store_true", help="also report mean C / K at each point")
a = ap.parse_args()
wl = WORKLOADS[a.workload]
points = [int(x) for x in a.at.split(",") if x]

for name in a.variants.split(","):
    kw = VARIANTS[name]
    ctypes.CDLL("libc.so.6").srand(1)
    st = sph.Settings(numParticles=wl["n"], randomInit=wl["randomInit"], boxDim=wl["boxDim"],
                      numCellsPerDim=wl["numCellsPerDim"])
    key = sph.SPH_KEY_MORTON if a.key == "morton" else sph.SPH_KEY_FLAT
    sim = sph.Simulator(st, key_mode=key, **kw)
    sim.setup()
    out = {"variant": name, "workload": a.workload, "key": a.key, "points": {}}
    done = 0
    for pt in points:
        sim.advance(pt - done)
        done = pt
        e = {}
        if a.counts:
            K, C = sim.get_neighbor_counts()
            e["meanC"], e["meanK"] = round(float(C.mean()), 2), round(float(K.mean()), 2)
        sim.profile_enable(True)
        sim.profile_read(reset=True)
        sim.advance(a.prof)
        done += a.prof
        prof = sim.profile_read(reset=True)
        sim.profile_enable(False)
        e.update({k: round(v["ms"] / a.prof, 4) for k, v in prof.items() if v["launches"]})
        e["sum"] = round(sum(v["ms"] for v in prof.values()) / a.prof, 4)
        out["points"][str(pt)] = e
    ke, mr = sim.get_stats()
    out["ke_after"], out["mean_rho_after"] = ke, mr
    sim.close()
    # whole run from the initial state, as bench.py's `value` times it
    ctypes.CDLL("libc.so.6").srand(1)
    sim = sph.Simulator(st, key_mode=key, **kw)
    sim.setup()
    sim.advance(3)
    ms = sim.advance_timed(a.total)
    out["advance_ms_per_step"] = round(ms / a.total, 4)
    out["updates_per_s"] = wl["n"] * a.total / (ms * 1e-3)
    sim.close()
    print(json.dumps(out), flush=True)
