"""CPU suite, world sizes 2 and 3 over gloo: the slab decomposition -- layer ranges, particle
partition, ghost halo exchanges A and B, migration -- as a Python model (tests/slab_model.py) with a
CPU stand-in for the library (tests/fake_slab.py), compared with the undecomposed CPU oracle.  The
product's protocol is C++/CUDA (csrc/sph_cluster.cu) and is covered on GPUs by
tests/test_gpu_cluster.py: several slabs on one GPU, peer-to-peer copies and NCCL on two."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import random_state
from cudafluidsimulator_b200.cluster import partition, slab_ranges
from slab_model import SlabDriver


def test_slab_ranges_cover_and_balance():
    for nz, w in [(100, 1), (100, 3), (256, 8), (7, 8)]:
        r = slab_ranges(nz, w)
        assert r[0][0] == 0 and r[-1][1] == nz and len(r) == w
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_partition_uses_reference_cell_rule():
    pos = np.float32([[5, 5, 0.1], [5, 5, 4.9999], [5, 5, 5.0], [5, 5, 9.9]])
    parts = partition(pos, 0.1, [(0, 50), (50, 100)])
    assert parts[0].tolist() == [0, 1] and parts[1].tolist() == [2, 3]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, pos, vel, steps, out, mode="gather"):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from fake_slab import FakeSlab, FastFakeSlab
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ranges = slab_ranges(100, world)
    mine = partition(pos, 0.1, ranges)[rank]
    cls = FakeSlab if mode == "gather" else FastFakeSlab
    b = cls(*ranges[rank], 100, capacity=len(pos), ghost_capacity=len(pos), emig_capacity=len(pos))
    b.load(pos[mine], vel[mine], mine.astype(np.uint32))
    drv = SlabDriver(b, rank, world, overlap=(mode == "overlap"))
    drv.guess_margin = 1                       # the test slabs have only a few particle CTAs
    for k in range(steps):
        if mode == "overlap" and k == 3:
            drv._guess = (0, 10 ** 6)          # a wrong guess: must be detected and redone
        drv.step()
    ids, p, v = b.download()
    out[rank] = (ids, p, v, drv.stats)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "gather"), (3, "gather"), (2, "neighbour"), (2, "overlap"), (3, "overlap")])
def test_two_slabs_match_undecomposed_oracle(world, mode):
    """mode: gather = counts through all_gather (any backend); neighbour = the library's protocol
    (device-side counts sent to the two neighbours only, _async / _finish halves); overlap = that
    plus interior / boundary parts with the speculative interior density and one injected wrong
    guess."""
    from oracle.oracle import CpuOracle
    # a blob straddling the slab boundary at z = 5.0 (and 3.4 / 6.7 for three slabs), moving in z
    rng = np.random.default_rng(5)
    n = 3000
    pos = (np.float32([3.0, 3.0, 3.0]) + rng.uniform(0, 1.0, (n, 3)) * np.float32([0.6, 0.6, 4.0])).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * np.float32([0.5, 0.5, 6.0])).astype(np.float32)
    steps = 6
    o = CpuOracle(n)
    o.set_state(pos, vel)
    for _ in range(steps):
        o.step()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), pos, vel, steps, out, mode), nprocs=world, join=True)
    ids = np.concatenate([out[r][0] for r in range(world)])
    p = np.concatenate([out[r][1] for r in range(world)])
    v = np.concatenate([out[r][2] for r in range(world)])
    assert sorted(ids.tolist()) == list(range(n)), "particles lost or duplicated by migration"
    order = np.argsort(ids)
    np.testing.assert_allclose(p[order], o.pos, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(v[order], o.vel, rtol=1e-3, atol=1e-3)
    assert sum(out[r][3]["migrated_particles"] for r in range(world)) > 0   # migration exercised
    assert sum(out[r][3]["ghost_particles"] for r in range(world)) > 0     # halos exercised
    if mode == "overlap":
        hits = [out[r][3].get("speculative_hits", 0) for r in range(world)]
        # guesses were used (slabs with enough particle CTAs), and the injected wrong one was rejected
        assert max(hits) > 0 and all(h <= steps - 2 for h in hits), hits
