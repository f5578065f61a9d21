#!/usr/bin/env python
"""What binning at h/2 would buy the neighbour loops: candidates per particle with
cells of edge h (9 x-runs of 3 cells) against cells of edge h/2 (25 x-runs of 5 cells), counted the way
the density kernel consumes them: every run widened to even slot bounds (aligned slot pairs feed the
packed f32x2 arithmetic), so each run costs on average one extra slot.  CPU only (numpy).

    python scripts/half_cell_estimate.py
"""
import numpy as np

H = np.float32(0.1)


def runs_cost(pos, edge, reach):
    """mean over interior particles of (candidates, slots after widening every x-run to even bounds)"""
    nc = int(np.ceil(pos.max() / edge)) + 1
    c = np.floor(pos / edge).astype(np.int64)
    key = (c[:, 2] * nc + c[:, 1]) * nc + c[:, 0]
    order = np.argsort(key, kind="stable")
    skey = key[order]
    start = np.searchsorted(skey, np.arange(nc ** 3 + 1))
    lo, hi = reach + 1, nc - reach - 2
    inner = np.all((c >= lo) & (c <= hi), axis=1)
    ci = c[inner]
    cand = np.zeros(len(ci))
    slots = np.zeros(len(ci))
    for dz in range(-reach, reach + 1):
        for dy in range(-reach, reach + 1):
            k0 = ((ci[:, 2] + dz) * nc + ci[:, 1] + dy) * nc + ci[:, 0] - reach
            s, e = start[k0], start[k0 + 2 * reach + 1]
            cand += e - s
            slots += np.where(e > s, (e + 1) // 2 * 2 - s // 2 * 2, 0)
    return cand.mean(), slots.mean()


def report(name, pos):
    c1, s1 = runs_cost(pos, float(H), 1)
    c2, s2 = runs_cost(pos, float(H) / 2, 2)
    d = pos[:, None, :] - pos[None, :400, :]
    print(f"{name}: cells of h: {c1:.1f} candidates, {s1:.1f} slots in 9 runs | cells of h/2: {c2:.1f} candidates, "
          f"{s2:.1f} slots in 25 runs | slots ratio {s2 / s1:.2f}")


rng = np.random.default_rng(0)
g = np.arange(40, dtype=np.float32) * np.float32(0.09) + H
lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
report("lattice 0.09 (early state, C 37)", lattice)
for per_cell in (4.3, 10.0):
    n = int(per_cell * 30 ** 3)
    report(f"uniform random, {per_cell} per cell", (rng.uniform(0, 3.0, (n, 3)) + 0.1).astype(np.float32))
