"""Runs representative workloads against the self-checking build (SPH_B200_LIB must point at
libsph_b200_checked.so) and prints the violation bits.  Own bounds checks stand in for
compute-sanitizer, which is closed on the GPU pool."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cudafluidsimulator_b200 as sph  # noqa: E402
from cudafluidsimulator_b200.cluster import Cluster, partition, slab_ranges  # noqa: E402
from conftest import compressed_state, lattice_state, random_state  # noqa: E402

total = 0
cases = {
    "lattice": lattice_state(109 * 109 * 3 + 777),
    "random": random_state(50000, seed=3, vel_scale=2.0),
    "compressed": compressed_state(9000, seed=4),
    "walls": (np.float32([[0.1, 0.1, 0.1], [9.9, 9.9, 9.9], [0.0, 0.0, 0.0], [9.99, 9.99, 9.99]]),
              np.float32([[-1, -1, -1], [1, 1, 1], [0, 0, 0], [2, 2, 2]])),
}
for name, (pos, vel) in cases.items():
    for mode in (sph.SPH_KEY_FLAT, sph.SPH_KEY_MORTON):
        sim = sph.Simulator(sph.Settings(numParticles=len(pos)), key_mode=mode, record_force=True)
        sim.setup()
        sim.set_state(pos, vel)
        for _ in range(6):
            sim.simulate()
        sim.get_neighbor_counts()
        sim.advance(4)
        sim.simulate()
        sim.moveParticles((400, 300))
        flags, checked = sim.debug_flags()
        assert checked, "not the self-checking build"
        print(f"{name} mode {mode}: flags {flags}")
        total |= flags
        sim.close()

# pipelined readback
pos, vel = cases["random"]
sim = sph.Simulator(sph.Settings(numParticles=len(pos)), pipeline_readback=True)
sim.setup()
sim.set_state(pos, vel)
for _ in range(6):
    sim.simulate()
total |= sim.debug_flags()[0]
sim.close()

# slabs (ghosts, emigrants)
rng = np.random.default_rng(5)
n = 20000
pos = (np.float32([3.0, 3.0, 2.5]) + rng.uniform(0, 1.0, (n, 3)) * np.float32([1.5, 1.5, 5.0])).astype(np.float32)
vel = (rng.standard_normal((n, 3)) * np.float32([0.5, 0.5, 6.0])).astype(np.float32)
cl = Cluster(sph.Settings(numParticles=n), world=3, devices=[0, 0, 0], capacity=n + 1024, ghost_capacity=n + 2,
             emig_capacity=n, rebalance_every=3)
for i, idx in enumerate(partition(pos, 0.1, slab_ranges(100, 3))):
    cl.load(i, pos[idx], vel[idx], idx.astype(np.uint32))
cl.advance(6)
for _ in range(4):
    cl.step()
cl.sync()
for i in range(3):
    st = cl.stats(i)
    assert st["checked_build"], "not the self-checking build"
    print("slab flags", st["debug_flags"])
    total |= st["debug_flags"]
cl.close()
print("CHECKED_BUILD_FLAGS", total)
