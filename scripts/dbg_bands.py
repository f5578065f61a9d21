import sys, numpy as np
sys.path.insert(0, '.')
import cudafluidsimulator_b200 as sph
from cudafluidsimulator_b200.cluster import Cluster, partition, slab_ranges
from bench import slosh_velocity
world, nc = int(sys.argv[1]), 100
rb = int(sys.argv[2])
h32, sp = np.float32(0.1), np.float32(0.09)
g = np.arange(109, dtype=np.float32)
x, y, z = np.meshgrid(h32 + sp * g[:50], h32 + sp * g[:60], h32 + sp * g[:108], indexing="ij")
pos = np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.float32)
ids = np.arange(len(pos), dtype=np.uint32)
vel = slosh_velocity(pos, ids, 1.0, 0.0, 2.8)
plane = 50 * 60
st = sph.Settings(numParticles=len(pos))
cl = Cluster(st, world=world, devices=[0] * world, capacity=len(pos), ghost_capacity=5 * plane + 1024, emig_capacity=3 * plane + 1024, rebalance_every=rb)
for i, idx in enumerate(partition(pos, 0.1, slab_ranges(nc, world))):
    cl.load(i, pos[idx], vel[idx], ids[idx])
for k in range(25):
    try:
        cl.advance(4)
    except Exception as e:
        print("FAILED at", 4 * (k + 1), e)
        break
    print(4 * (k + 1), [(s["n_owned"], s["ghosts_lo"], s["ghosts_hi"], s["z_cell_lo"], s["z_cell_hi"], s["rebalances"], s["migrated_total"]) for s in (cl.stats(i) for i in range(world))])
