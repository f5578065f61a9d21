"""torchrun worker of tests/test_gpu_cluster.py::test_nccl_transport_two_gpus: one process per GPU,
messages over ncclSend / ncclRecv inside the library; the id of the NCCL communicator is broadcast
over a gloo group (torch is only the launcher and the bootstrap here)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))

import cudafluidsimulator_b200 as sph  # noqa: E402
from cudafluidsimulator_b200.cluster import Cluster, nccl_id, partition, slab_ranges  # noqa: E402
from oracle.oracle import CpuOracle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
box = [nccl_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)

rng = np.random.default_rng(5)
n = 20000
pos = (np.float32([3.0, 3.0, 2.5]) + rng.uniform(0, 1.0, (n, 3)) * np.float32([1.5, 1.5, 5.0])).astype(np.float32)
vel = (rng.standard_normal((n, 3)) * np.float32([0.5, 0.5, 6.0])).astype(np.float32)
steps = 12
st = sph.Settings(numParticles=n)
cl = Cluster(st, world=world, rank=rank, devices=[local], capacity=n + 1024, ghost_capacity=n + 2,
             emig_capacity=n + 1024, nccl_id=box[0], rebalance_every=4)
idx = partition(pos, st.h, slab_ranges(100, world))[rank]
cl.load(0, pos[idx], vel[idx], idx.astype(np.uint32))
cl.advance(steps)
ids, p, v = cl.download(0, n + 1024)
stats = cl.stats(0)
parts = [None] * world
dist.all_gather_object(parts, (ids, p, stats))
if rank == 0:
    o = CpuOracle(n)
    o.set_state(pos, vel)
    for _ in range(steps):
        o.step()
    all_ids = np.concatenate([q[0] for q in parts])
    all_p = np.concatenate([q[1] for q in parts])
    order = np.argsort(all_ids)
    assert all_ids[order].tolist() == list(range(n)), "particles lost or duplicated"
    np.testing.assert_allclose(all_p[order], o.pos, rtol=3e-5, atol=3e-6)
    assert sum(q[2]["migrated_total"] for q in parts) > 0
    print("CLUSTER_NCCL_OK", [q[2]["n_owned"] for q in parts], [q[2]["migrated_total"] for q in parts],
          [q[2]["rebalances"] for q in parts], flush=True)
cl.close()
dist.barrier()
dist.destroy_process_group()
