#!/usr/bin/env python
"""Generates tests/golden/ref_*.npz from the reference's OWN CUDA implementation
(oracle/_ref/libsph_ref.so = /root/reference/src/simulator.cu compiled unmodified).

Run on a GPU box:   gpurun -- 'python scripts/make_golden.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz into tests/golden/ and commit them.  The CPU suite
(tests/test_oracle.py) pins the C restatement against these vectors without a GPU.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from conftest import compressed_state, developed_state, lattice_state, random_state  # noqa: E402
from oracle.oracle import RefSim  # noqa: E402

CASES = {
    "lattice_sheet_4k": (lambda: lattice_state(4000), 20),
    "lattice_3d_13k": (lambda: lattice_state(109 * 109 + 1500), 20),
    "random_6k": (lambda: random_state(6000, seed=21, lo=3.0, hi=5.0, vel_scale=1.0), 20),
    "compressed_3k": (lambda: compressed_state(3000, seed=22), 10),
    "developed_grid_4k_50": (lambda: developed_state(4000, 50), 20),
}


def main(out_dir):
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    cfg = dict(h=0.1, boxDim=10.0, numCellsPerDim=100.0, timestep=0.01)
    for name, (make, steps) in CASES.items():
        pos, vel = make()
        ref = RefSim(len(pos), **cfg)
        ref.set_state(pos, vel)
        cells, flat = ref.keys()
        cnt = ref.neighbor_counts()
        ref.step()
        s1 = ref.get_state()
        for _ in range(steps - 1):
            ref.step()
        sN = ref.get_state()
        ref.close()
        ke = 0.5 * 0.02 * float((sN["vel"].astype(np.float64) ** 2).sum())
        np.savez_compressed(
            out / f"ref_{name}.npz", pos0=pos, vel0=vel, cells=cells, flat=flat, list_of=cnt["list_of"],
            K=cnt["K"], C=cnt["C"], Knz=cnt["Knz"], rho1=s1["rho"], prs1=s1["prs"], force1=s1["force"],
            pos1=s1["pos"], vel1=s1["vel"], steps=steps, ke=ke,
            mean_rho=float(sN["rho"].astype(np.float64).mean()),
            **{k: np.float32(v) for k, v in cfg.items()})
        print(name, len(pos), "K mean", cnt["K"].mean(), "p>0", float((s1["prs"] > 0).mean()),
              "Knz==K", bool((cnt["Knz"] == cnt["K"]).all()), flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "tests" / "golden"))
