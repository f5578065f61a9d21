#!/usr/bin/env python
"""One slab on one GPU through the cluster path (the weak-scaling reference of bench.py --gpus N):
ms per step of the 32m_grid per-GPU problem.  SPH_SORT=radix selects the radix passes.

    python scripts/cluster_n1.py [--workload 32m_grid] [--steps 20] [--warmup 5]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import cudafluidsimulator_b200 as sph  # noqa: E402
from bench import WORKLOADS, slosh_velocity, stretched_lattice  # noqa: E402
from cudafluidsimulator_b200.cluster import Cluster  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="32m_grid", choices=list(WORKLOADS))
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
a = ap.parse_args()
wl = WORKLOADS[a.workload]
p1, i1, n_glob, nz1 = stretched_lattice(wl, 1, lambda nzz: (0, nzz), None)
p1[:, 2] += np.float32(1.6)
st = sph.Settings(numParticles=n_glob, randomInit=False, boxDim=wl["boxDim"], numCellsPerDim=wl["numCellsPerDim"])
c1 = Cluster(st, world=1, rank=0, devices=[0], nz_cells=nz1 + 32, capacity=int(len(i1) * 1.05) + 65536,
             ghost_capacity=1024, emig_capacity=1024)
c1.load(0, p1, slosh_velocity(p1, i1, 1.0, 0.0, 0.5 * wl["boxDim"]), i1)
c1.advance(a.warmup)
ms = c1.advance_timed(a.steps) / a.steps
stats = c1.stats(0)
print(json.dumps({"workload": a.workload, "particles": int(n_glob), "ms_per_step": ms,
                  "updates_per_s": n_glob / (ms * 1e-3), "kinetic_energy": stats["kinetic_energy"]}))
c1.close()
