#!/usr/bin/env python
"""One plainly-launched timestep inside a cudaProfilerStart/Stop range, for ncu:

  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -o gpurun_out/prof python scripts/profile_step.py --workload 16m_grid --pre 100

`--pre` steps are replayed as a CUDA graph first so the profiled step sees a developed
state (floor pile-up), then per-launch mode is switched on and `--steps` steps run.
"""
import argparse
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import cudafluidsimulator_b200 as sph  # noqa: E402
from bench import WORKLOADS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="16m_grid", choices=list(WORKLOADS))
ap.add_argument("--pre", type=int, default=100)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--key", default="flat")
ap.add_argument("--stage-tiles", action="store_true", help="SphOptions.stage_tiles (TMA-staged dense CTAs)")
a = ap.parse_args()
wl = WORKLOADS[a.workload]
ctypes.CDLL("libc.so.6").srand(1)
sim = sph.Simulator(sph.Settings(numParticles=wl["n"], randomInit=wl["randomInit"], boxDim=wl["boxDim"],
                                 numCellsPerDim=wl["numCellsPerDim"]),
                    key_mode=sph.SPH_KEY_MORTON if a.key == "morton" else sph.SPH_KEY_FLAT,
                    stage_tiles=a.stage_tiles)
sim.setup()
sim.advance(a.pre)
K, C = sim.get_neighbor_counts()
print(f"state after {a.pre} steps: mean C {C.mean():.1f} (max {C.max()}), mean K {K.mean():.1f} (max {K.max()})")
sim.profile_enable(True)          # plain launches, one CUDA event pair per kernel
torch.cuda.synchronize()
torch.cuda.profiler.start()
sim.advance(a.steps)
torch.cuda.profiler.stop()
for k, v in sim.profile_read().items():
    if v["launches"]:
        print(f"{k:20s} {v['ms'] / a.steps:9.4f} ms/step  ({v['launches']} launches)")
sim.close()
