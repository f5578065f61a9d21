"""GPU suite: the multi-GPU cluster entry points (sph_cluster_*, csrc/sph_cluster.cu) against the
single-GPU path and the CPU oracle.  On a one-GPU box every slab lives on device 0 (messages
move with device-to-device copies instead of NVLink P2P / NCCL; the kernels, the device-side
counts and the ordering are the same); with >= 2 GPUs the NCCL transport runs under torchrun.

  first step, density summed in reference order     bit-identical to the single-GPU step
  several steps with migration                      positions vs the oracle, rtol 3e-5 / atol 3e-6
  ids                                               every particle exactly once, in some slab
  rebalancing                                       layers change owner, nothing lost, same physics
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cudafluidsimulator_b200 as sph
from cudafluidsimulator_b200.cluster import Cluster, partition, slab_ranges
from conftest import compressed_state, lattice_state, random_state
from oracle.oracle import CpuOracle

pytestmark = pytest.mark.gpu


def make_cluster(pos, vel, world, nz=100, **kw):
    st = sph.Settings(numParticles=len(pos))
    n = len(pos)
    cl = Cluster(st, world=world, devices=[0] * world, nz_cells=nz, capacity=n + 1024,
                 ghost_capacity=n + 2, emig_capacity=n + 1024, **kw)
    for i, idx in enumerate(partition(pos, st.h, slab_ranges(nz, world))):
        cl.load(i, pos[idx], vel[idx], idx.astype(np.uint32))
    return cl


def straddling_blob(n=20000, seed=5):
    rng = np.random.default_rng(seed)
    pos = (np.float32([3.0, 3.0, 2.5]) + rng.uniform(0, 1.0, (n, 3)) * np.float32([1.5, 1.5, 5.0])).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * np.float32([0.5, 0.5, 6.0])).astype(np.float32)
    return pos, vel


def state(name):
    if name == "blob":
        return straddling_blob()
    if name == "compressed":
        pos, vel = compressed_state(8000, seed=3, origin=(2.0, 0.1, 4.6))   # sits on z = 5.0
        vel[:, 2] *= 3
        return pos, vel
    return lattice_state(109 * 109 * 2)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_first_step_bit_exact_vs_single_gpu(world):
    """Step 1 from an id-ordered state: same pairs, same order => bit-identical (with the density
    summed term by term; the factored default groups terms by slot parity, which differs between
    a slab and the whole box)."""
    pos, vel = straddling_blob()
    ref = sph.Simulator(sph.Settings(numParticles=len(pos)), density_sum=1)
    ref.setup()
    ref.set_state(pos, vel)
    ref.simulate()
    p_ref, v_ref = ref.get_state()
    ref.close()
    cl = make_cluster(pos, vel, world, density_sum=1)
    cl.advance(1)
    ids, p, v = cl.download_all(len(pos) + 1024)
    assert ids.tolist() == list(range(len(pos)))
    np.testing.assert_array_equal(p, p_ref)
    np.testing.assert_array_equal(v, v_ref)
    cl.close()


@pytest.mark.parametrize("name,world,steps", [("blob", 2, 12), ("blob", 4, 12), ("compressed", 2, 8),
                                              ("lattice", 3, 10)])
def test_multi_step_vs_oracle_with_migration(name, world, steps):
    pos, vel = state(name)
    o = CpuOracle(len(pos))
    o.set_state(pos, vel)
    for _ in range(steps):
        o.step()
    cl = make_cluster(pos, vel, world)
    cl.advance(steps)
    ids, p, v = cl.download_all(len(pos) + 1024)
    assert ids.tolist() == list(range(len(pos))), "particles lost or duplicated"
    np.testing.assert_allclose(p, o.pos, rtol=3e-5, atol=3e-6)
    st = [cl.stats(i) for i in range(world)]
    assert all(s["steps"] == steps and s["overflow"] == 0 for s in st)
    assert sum(s["n_owned"] for s in st) == len(pos)
    if name != "lattice":
        assert sum(s["migrated_total"] for s in st) > 0
        assert sum(s["ghosts_total"] for s in st) > 0
    # id-ordered positions, the cluster's getPosition()
    np.testing.assert_array_equal(cl.positions(len(pos)), p)
    cl.close()


def test_step_by_step_equals_one_advance_and_delivers_host_records():
    pos, vel = straddling_blob(12000)
    a = make_cluster(pos, vel, 3)
    a.advance(6)
    ref = a.download_all(len(pos) + 1024)
    a.close()
    b = make_cluster(pos, vel, 3)
    for _ in range(6):
        b.step()               # one step + the owned particles' records to pinned host memory
    b.sync()
    got = b.download_all(len(pos) + 1024)
    np.testing.assert_array_equal(got[0], ref[0])
    # (immigrants are appended in the order the emigrant atomics fired: rounding-level differences)
    np.testing.assert_allclose(got[1], ref[1], rtol=2e-5, atol=2e-6)
    seen = {}
    for i in range(3):
        rec = b.host_records(i).copy()
        ids = rec[:, 3].view(np.uint32)
        live = ids != 0xFFFFFFFF
        for pid, xyz in zip(ids[live], rec[live, :3]):
            assert pid not in seen, "a particle appears in two slabs' host records"
            seen[int(pid)] = xyz
    assert sorted(seen) == list(range(len(pos)))
    np.testing.assert_array_equal(np.stack([seen[i] for i in range(len(pos))]), got[1])
    b.close()


def test_reference_init_through_the_cluster_matches_single_simulator():
    """sph_cluster_setup(): the reference's lattice, same ids; 5 steps against one simulator."""
    n = 109 * 109 * 3
    st = sph.Settings(numParticles=n)
    ref = sph.Simulator(st)
    ref.setup()
    ref.advance(5)
    p_ref, _ = ref.get_state()
    ref.close()
    cl = Cluster(st, world=4, devices=[0] * 4, capacity=n, ghost_capacity=n, emig_capacity=n)
    cl.setup()
    cl.advance(5)
    p = cl.positions(n)
    assert np.isfinite(p).all()
    np.testing.assert_allclose(p, p_ref, rtol=3e-5, atol=3e-6)
    cl.close()


def test_non_cubic_global_box_weak_scaling_layout():
    """bench.py's weak-scaling layout: sub-boxes replicated along z (nz = world * nc)."""
    world, nc = 2, 100
    pos0, vel0 = random_state(20000, seed=3, lo=1.0, hi=9.0, vel_scale=1.0)
    pos = np.concatenate([pos0 + np.float32([0, 0, 10.0 * r]) for r in range(world)]).astype(np.float32)
    vel = np.concatenate([vel0] * world)
    cl = make_cluster(pos, vel, world, nz=world * nc)
    cl.advance(5)
    ids, p, v = cl.download_all(len(pos) + 1024)
    assert len(ids) == len(pos) and np.isfinite(p).all()
    assert p[:, 2].max() <= 10.0 * world - 0.1 + 1e-6 and p[:, 2].min() >= 0.1 - 1e-6
    cl.close()


def test_rebalancing_moves_layers_and_keeps_the_physics():
    """A blob that sits mostly in one slab: rebalancing hands layers to the lighter neighbours
    through the migration messages; nothing is lost and the trajectory stays the oracle's."""
    rng = np.random.default_rng(11)
    n = 24000
    pos = (np.float32([3.0, 3.0, 1.2]) + rng.uniform(0, 1.0, (n, 3)) * np.float32([1.5, 1.5, 3.3])).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * np.float32([0.3, 0.3, 1.0])).astype(np.float32)
    steps = 12
    o = CpuOracle(n)
    o.set_state(pos, vel)
    for _ in range(steps):
        o.step()
    cl = make_cluster(pos, vel, 4, rebalance_every=2)
    before = [cl.stats(i) for i in range(4)]
    cl.advance(steps)
    after = [cl.stats(i) for i in range(4)]
    assert sum(s["rebalances"] for s in after) > 0
    assert [(s["z_cell_lo"], s["z_cell_hi"]) for s in after] != [(s["z_cell_lo"], s["z_cell_hi"]) for s in before]
    # contiguous cover of the box
    assert after[0]["z_cell_lo"] == 0 and after[-1]["z_cell_hi"] == 100
    assert all(after[i]["z_cell_hi"] == after[i + 1]["z_cell_lo"] for i in range(3))
    imb = lambda ss: max(s["n_owned"] for s in ss) / (sum(s["n_owned"] for s in ss) / len(ss))
    assert imb(after) < imb(before)
    ids, p, v = cl.download_all(n + 1024)
    assert ids.tolist() == list(range(n)), "particles lost or duplicated"
    np.testing.assert_allclose(p, o.pos, rtol=3e-5, atol=3e-6)
    cl.close()


def test_capacity_overflow_is_reported():
    pos, vel = straddling_blob(8000)
    st = sph.Settings(numParticles=len(pos))
    cl = Cluster(st, world=2, devices=[0, 0], capacity=len(pos), ghost_capacity=16, emig_capacity=len(pos))
    for i, idx in enumerate(partition(pos, st.h, slab_ranges(100, 2))):
        cl.load(i, pos[idx], vel[idx], idx.astype(np.uint32))
    with pytest.raises(sph.SphError, match="capacity exceeded"):
        cl.advance(2)
    cl.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_transport_two_gpus():
    script = os.path.join(os.path.dirname(__file__), "cluster_nccl_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", script],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "CLUSTER_NCCL_OK" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_to_peer_transport_two_gpus():
    """One process, one slab per GPU: cudaMemcpyPeerAsync between the devices."""
    pos, vel = straddling_blob()
    steps = 10
    o = CpuOracle(len(pos))
    o.set_state(pos, vel)
    for _ in range(steps):
        o.step()
    st = sph.Settings(numParticles=len(pos))
    n = len(pos)
    cl = Cluster(st, world=2, devices=[0, 1], capacity=n + 1024, ghost_capacity=n + 2, emig_capacity=n + 1024)
    for i, idx in enumerate(partition(pos, st.h, slab_ranges(100, 2))):
        cl.load(i, pos[idx], vel[idx], idx.astype(np.uint32))
    cl.advance(steps)
    ids, p, v = cl.download_all(n + 1024)
    assert ids.tolist() == list(range(n))
    np.testing.assert_allclose(p, o.pos, rtol=3e-5, atol=3e-6)
    cl.close()
