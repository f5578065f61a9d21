"""GPU suite: the CUDA path (libsph_b200.so through its C ABI) against the CPU oracle
on identical seeded states.

Bars (north_star): cell keys, sorted order and neighbour counts bit-exact; single-step
density, force and position within a stated FP32 relative tolerance; multi-step
aggregates within a stated bound.

  integers (keys, order, cell table, K, C)            bit-exact
  density, pressure                                   density_sum=1 (the reference's term-by-term
                                                      sum): bit-exact vs the oracle (same
                                                      arithmetic, same visiting order);
                                                      default (factored sum): rtol 2e-6
  force            |dF| <= 1e-5 * max(|F|, sum|terms|) per component
  position         rtol 1e-5 / atol 1e-6 after one step
"""
import numpy as np
import pytest

import cudafluidsimulator_b200 as sph
from conftest import (compressed_state, developed_state, force_tolerance, lattice_state,
                      random_state)
from oracle.oracle import CpuOracle

pytestmark = pytest.mark.gpu

RTOL_POS = 1e-5
RTOL_F = 1e-5


RTOL_RHO = 2e-6   # density / pressure when the sum is not formed in the oracle's order


def make(n, key_mode=sph.SPH_KEY_FLAT, density_sum=0, sort_algo=0, **kw):
    s = sph.Settings(numParticles=n, **kw)
    sim = sph.Simulator(s, key_mode=key_mode, record_force=True, density_sum=density_sum, sort_algo=sort_algo)
    sim.setup()
    return sim


STATES = {
    "lattice_sheet_10k": lambda: lattice_state(10000),
    "lattice_3d_40k": lambda: lattice_state(109 * 109 * 3 + 777),
    "random_30k": lambda: random_state(30000, seed=7),
    "random_moving_20k": lambda: random_state(20000, seed=8, lo=2.0, hi=5.0, vel_scale=2.0),
    "compressed_6k": lambda: compressed_state(6000),
    "developed_grid_8k_60": lambda: developed_state(8000, 60),
    "ragged_4097": lambda: random_state(4097, seed=9),
    "single": lambda: (np.float32([[5, 5, 5]]), np.float32([[0.3, 0, 0]])),
    "two_coincident": lambda: (np.float32([[5, 5, 5], [5, 5, 5]]), np.zeros((2, 3), np.float32)),
    "walls": lambda: (np.float32([[0.1, 0.1, 0.1], [9.9, 9.9, 9.9], [0.1, 9.9, 5.0], [0.0, 0.0, 0.0],
                                  [9.99, 9.99, 9.99]]),
                      np.float32([[-1, -1, -1], [1, 1, 1], [-5, 5, 0], [0, 0, 0], [2, 2, 2]])),
}


@pytest.mark.parametrize("name", list(STATES))
@pytest.mark.parametrize("mode", [sph.SPH_KEY_FLAT, sph.SPH_KEY_MORTON], ids=["flat", "morton"])
def test_keys_order_counts_bit_exact(name, mode):
    pos, vel = STATES[name]()
    n = len(pos)
    o = CpuOracle(n)
    sim = make(n, key_mode=mode)
    sim.set_state(pos, vel)
    cells, ff, fi, mo = o.keys(pos)
    # (1) hashing, both forms, regardless of the sort's key mode
    np.testing.assert_array_equal(sim.get_keys(sph.SPH_KEY_FLAT), fi.astype(np.uint32))
    np.testing.assert_array_equal(sim.get_keys(sph.SPH_KEY_MORTON), mo)
    # (2) neighbour and candidate counts
    rho, prs, K, C = o.density(pos)
    Kg, Cg = sim.get_neighbor_counts()
    np.testing.assert_array_equal(Kg, K)
    np.testing.assert_array_equal(Cg, C)
    # (3) sorted order == stable sort by key of the original order; cell table
    sim.simulate()
    ids, skeys = sim.get_sorted_index()
    keys = fi.astype(np.uint32) if mode == sph.SPH_KEY_FLAT else mo
    table = 100 ** 3 if mode == sph.SPH_KEY_FLAT else 128 ** 3
    order, start = o.sort_order(keys, table)
    np.testing.assert_array_equal(ids, order.astype(np.uint32))
    np.testing.assert_array_equal(skeys, keys[order])
    np.testing.assert_array_equal(sim.get_cell_start(), start.astype(np.uint32))
    sim.close()


@pytest.mark.parametrize("name", list(STATES))
@pytest.mark.parametrize("mode,density_sum", [(sph.SPH_KEY_FLAT, 0), (sph.SPH_KEY_FLAT, 1), (sph.SPH_KEY_MORTON, 0)],
                         ids=["flat", "flat-refsum", "morton"])
def test_single_step_density_force_position(name, mode, density_sum):
    pos, vel = STATES[name]()
    n = len(pos)
    o = CpuOracle(n)
    o.set_state(pos, vel)
    o.step()
    sim = make(n, key_mode=mode, density_sum=density_sum)
    sim.set_state(pos, vel)
    sim.simulate()
    rho, prs, f = sim.get_density_pressure_force()
    if mode == sph.SPH_KEY_FLAT and density_sum == 1:
        # same arithmetic and same visiting order as the oracle: bit-exact
        np.testing.assert_array_equal(rho, o.rho)
        np.testing.assert_array_equal(prs, o.prs)
    else:
        np.testing.assert_allclose(rho, o.rho, rtol=RTOL_RHO, atol=0)
        np.testing.assert_allclose(prs, o.prs, rtol=0, atol=RTOL_RHO * float(o.rho.max()))
    tol = force_tolerance(o, pos, vel, o.rho, o.prs, rel=RTOL_F)
    err = np.abs(f - o.force)
    assert np.all(err <= tol), f"force: worst excess {np.max(err / tol):.3g}x tolerance"
    p1, v1 = sim.get_state()
    np.testing.assert_allclose(p1, o.pos, rtol=RTOL_POS, atol=1e-6)
    # velocity inherits the force tolerance: dv = dt * dF / rho
    vtol = 0.01 * tol / o.rho[:, None] + 1e-5 * np.abs(o.vel) + 1e-6
    flipped = (np.abs(o.vel) < 2e-4) | (np.abs(v1) < 2e-4)  # |v| < 1e-4 -> 0 is discontinuous
    assert np.all((np.abs(v1 - o.vel) <= vtol) | flipped)
    # the host readback is the reference's position[i], original order
    np.testing.assert_array_equal(sim.getPosition(), p1)
    sim.close()


def test_first_step_free_fall_bit_exact():
    pos, vel = lattice_state(20000)
    o = CpuOracle(len(pos))
    o.set_state(pos, vel)
    o.step()
    sim = make(len(pos))
    sim.simulate()  # setup() itself produced the lattice (ref: simulator.cu:438-453)
    np.testing.assert_array_equal(sim.getPosition(), o.pos)
    sim.close()


def test_setup_random_init_matches_reference_rand():
    import ctypes
    ctypes.CDLL("libc.so.6").srand(1)  # the reference's unseeded state
    sim = make(1000, randomInit=True)
    o = CpuOracle(1000, randomInit=True)
    o.setup()
    p, v = sim.get_state()
    np.testing.assert_array_equal(p, o.pos)
    assert not v.any()
    sim.close()


SORTS = [sph.SPH_SORT_COUNT, sph.SPH_SORT_RADIX]


@pytest.mark.parametrize("algo", SORTS, ids=["count", "radix"])
@pytest.mark.parametrize("mode", [sph.SPH_KEY_FLAT, sph.SPH_KEY_MORTON], ids=["flat", "morton"])
def test_sort_is_stable_across_steps(mode, algo):
    """After step k the storage order is step k's sorted order; step k+1 must be a
    STABLE sort of that order by the new keys (ties keep the previous order) -- by the radix
    passes, and by the counting sort by cell (whose members of a cell arrive in atomic order
    and are ranked by index in the reorder kernel)."""
    pos, vel = random_state(50000, seed=11, lo=1.0, hi=4.0, vel_scale=3.0)
    sim = make(len(pos), key_mode=mode, sort_algo=algo)
    assert sim.sort_info()["algo"] == ("count" if algo == sph.SPH_SORT_COUNT else "radix")
    sim.set_state(pos, vel)
    sim.simulate()
    prev_ids, _ = sim.get_sorted_index()
    for _ in range(3):
        keys_now = sim.get_keys(mode)                  # by original id, current positions
        sim.simulate()
        ids, skeys = sim.get_sorted_index()
        k_in_storage_order = keys_now[prev_ids]
        expect = prev_ids[np.argsort(k_in_storage_order, kind="stable")]
        np.testing.assert_array_equal(ids, expect)
        np.testing.assert_array_equal(skeys, np.sort(k_in_storage_order))
        assert np.array_equal(np.sort(ids), np.arange(len(pos), dtype=np.uint32))
        prev_ids = ids
    sim.close()


@pytest.mark.parametrize("mode", [sph.SPH_KEY_FLAT, sph.SPH_KEY_MORTON], ids=["flat", "morton"])
@pytest.mark.parametrize("name", ["random_moving_20k", "compressed_6k", "crowded_cells", "ragged_4097"])
def test_counting_sort_and_radix_sort_agree_bitwise(name, mode):
    """The two sorts of the step (SphOptions.sort_algo) give the same order, cell table and -- the
    summation order being the same -- bit-identical states, step after step.  crowded_cells: 3000
    particles in a block of 2 x 2 x 2 cells (hundreds of members per cell to rank)."""
    if name == "crowded_cells":
        rng = np.random.default_rng(5)
        pos = (np.float32([3.0, 0.1, 3.0]) + rng.uniform(0, 0.2, (3000, 3))).astype(np.float32)
        vel = rng.standard_normal((3000, 3)).astype(np.float32)
    else:
        pos, vel = STATES[name]()
    out = []
    for algo in SORTS:
        sim = make(len(pos), key_mode=mode, sort_algo=algo)
        sim.set_state(pos, vel)
        per_step = []
        for _ in range(4):
            sim.simulate()
            ids, skeys = sim.get_sorted_index()
            per_step.append((ids, skeys, sim.get_cell_start()))
        sim.advance(6)    # graph replay
        per_step.append(sim.get_state())
        out.append(per_step)
        sim.close()
    for a, b in zip(out[0][:-1], out[1][:-1]):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(out[0][-1][0], out[1][-1][0])
    np.testing.assert_array_equal(out[0][-1][1], out[1][-1][1])


def test_state_upload_between_steps_discards_the_fused_count():
    """The force kernel leaves the next step's per-cell counts behind; a state uploaded between two
    steps (or a neighbour-count query, which rebuilds the grid) must not see them."""
    pos1, vel1 = random_state(20000, seed=31, lo=1.0, hi=3.0, vel_scale=2.0)
    pos2, vel2 = compressed_state(6000, seed=4)
    pos2 = np.concatenate([pos2, pos1[:14000] + np.float32([4.0, 0.0, 4.0])]).astype(np.float32)
    vel2 = np.concatenate([vel2, vel1[:14000]]).astype(np.float32)
    sim = make(len(pos1))
    sim.set_state(pos1, vel1)
    sim.advance(3)
    sim.set_state(pos2, vel2)          # counts of pos1's step 3 are still in the table
    sim.advance(2)
    K, C = sim.get_neighbor_counts()   # rebuilds the grid: consumes the counts of step 2
    sim.advance(2)
    got = sim.get_state()
    fresh = make(len(pos2))
    fresh.set_state(pos2, vel2)
    fresh.advance(2)
    K2, C2 = fresh.get_neighbor_counts()
    fresh.advance(2)
    want = fresh.get_state()
    np.testing.assert_array_equal(K, K2)
    np.testing.assert_array_equal(C, C2)
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])
    sim.close()
    fresh.close()


@pytest.mark.parametrize("name,steps", [("lattice_3d_40k", 40), ("random_30k", 30), ("compressed_6k", 20)])
def test_multi_step_aggregates(name, steps):
    """Trajectories diverge chaotically, so after many steps only aggregates are
    compared: kinetic energy and mean density within 1 %."""
    pos, vel = STATES[name]()
    o = CpuOracle(len(pos))
    o.set_state(pos, vel)
    sim = make(len(pos))
    sim.set_state(pos, vel)
    for _ in range(steps):
        o.step()
    sim.advance(steps)
    ke_o, rho_o = o.stats()
    ke_g, rho_g = sim.get_stats()
    assert abs(ke_g - ke_o) <= 0.01 * abs(ke_o) + 1e-9
    assert abs(rho_g - rho_o) <= 0.01 * rho_o
    sim.close()


def test_graph_and_plain_launch_paths_agree_bitwise():
    pos, vel = random_state(30000, seed=5, vel_scale=1.0)
    out = []
    for use_graph in (True, False):
        s = sph.Settings(numParticles=len(pos))
        sim = sph.Simulator(s, use_graph=use_graph)
        sim.setup()
        sim.set_state(pos, vel)
        sim.advance(7)
        out.append(sim.get_state())
        sim.close()
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])


def test_flat_and_morton_sorts_give_same_physics():
    pos, vel = compressed_state(5000, seed=3)
    res = []
    for mode in (sph.SPH_KEY_FLAT, sph.SPH_KEY_MORTON):
        sim = make(len(pos), key_mode=mode)
        sim.set_state(pos, vel)
        sim.simulate()
        res.append(sim.get_density_pressure_force() + sim.get_state())
        sim.close()
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=2e-6)
    np.testing.assert_allclose(res[0][3], res[1][3], rtol=1e-5, atol=1e-6)


def test_timed_step_fills_reference_buckets():
    sim = make(10000)
    t = sph.Times()
    for _ in range(5):
        sim.simulateAndTime(t)
    assert t.iters == 5 and t.buildGrid > 0 and t.sphUpdate > 0 and t.memcpy > 0
    sim.close()


def test_mouse_push_matches_oracle():
    """The velocity increment of a click (+-2.5 / +-5 per axis, ref: simulator.cu:347-364) per
    particle, isolated by running the same step with and without the push: equal to the oracle's
    increment to 1e-5 (a missed or doubled half-push would be off by >= 2.5)."""
    pos, vel = random_state(40000, seed=13, lo=3.0, hi=7.0)
    o = CpuOracle(len(pos))
    o.set_state(pos, vel)
    o.step()
    v_before = o.vel.copy()
    o.push(pos, 400, 300)          # grid of the PRE-step positions
    d_oracle = o.vel - v_before
    out = []
    for push in (False, True):
        sim = make(len(pos))
        sim.set_state(pos, vel)
        sim.simulate()
        if push:
            sim.moveParticles((400, 300))
        out.append(sim.get_state()[1])
        sim.close()
    np.testing.assert_allclose(out[1] - out[0], d_oracle, rtol=0, atol=1e-5)
    assert (np.abs(d_oracle[:, 2]) > 4).sum() > 0 and (np.abs(d_oracle[:, 0]) > 2).sum() > 0   # it hit something
    # and the pushed state itself, within the single-step velocity tolerance
    np.testing.assert_allclose(out[1], o.vel, rtol=1e-5, atol=1e-4)


def test_scaled_domain_256_cells():
    """Config 3 geometry (h=.1, boxDim=25.6, 256 cells) at a test-sized N."""
    from oracle.oracle import CpuOracle
    n = 283 * 283 + 1000
    o = CpuOracle(n, boxDim=25.6, numCellsPerDim=256)
    o.setup()
    pos = o.pos.copy()
    o.step()
    for mode in (sph.SPH_KEY_FLAT, sph.SPH_KEY_MORTON):
        sim = make(n, key_mode=mode, boxDim=25.6, numCellsPerDim=256.0)
        cells, ff, fi, mo = o.keys(pos)
        np.testing.assert_array_equal(sim.get_keys(sph.SPH_KEY_FLAT), fi.astype(np.uint32))
        sim.simulate()
        np.testing.assert_array_equal(sim.getPosition(), o.pos)
        sim.close()


@pytest.mark.parametrize("n,init", [(1 << 20, "random"), (4_000_000, "grid")])
def test_full_size_properties(n, init):
    """At BASELINE sizes the oracle is too slow; check size-independent properties:
    sortedness, permutation, cell-table consistency, id-order readback."""
    box, cells = (10.0, 100.0) if init == "random" else (25.6, 256.0)
    sim = make(n, randomInit=(init == "random"), boxDim=box, numCellsPerDim=cells)
    p0, _ = sim.get_state()
    sim.simulate()
    ids, skeys = sim.get_sorted_index()
    assert np.all(np.diff(skeys.astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32))
    start = sim.get_cell_start()
    assert start[0] == 0 and start[-1] == n and np.all(np.diff(start.astype(np.int64)) >= 0)
    np.testing.assert_array_equal(np.diff(start.astype(np.int64)), np.bincount(skeys, minlength=len(start) - 1))
    # step 1 is free fall (F == 0): x and z untouched, y moved by g*dt*dt or clamped
    p1 = sim.getPosition()
    np.testing.assert_array_equal(p1[:, [0, 2]], p0[:, [0, 2]])
    dy = np.float32(np.float32(-9.8) * np.float32(0.01)) * np.float32(0.01)
    np.testing.assert_allclose(p1[:, 1], np.maximum(p0[:, 1] + dy, np.float32(0.1)), atol=2e-6)
    sim.advance(3)
    ke, mrho = sim.get_stats()
    assert np.isfinite(ke) and 25 < mrho < 2000
    sim.close()


def test_16m_developed_subdomain_matches_oracle():
    """BASELINE config 3 at its full size (16 M particles, 256^3 cells), developed for 60 steps so
    the floor pile-up exists.  The oracle cannot run 16 M, but interactions reach only h: a
    sub-domain plus a 2.5 h halo of the SAME state is a closed problem for the particles inside,
    so their candidate / neighbour counts must match bit-exactly and density, force and the
    next position within the single-step tolerances (order inside a cell follows the sort
    history here and the id order in the sub-problem, hence no bit-exact density)."""
    n = 16_000_000
    sim = make(n, boxDim=25.6, numCellsPerDim=256.0)
    sim.advance(60)
    pos, vel = sim.get_state()
    K, Cn = sim.get_neighbor_counts()
    sim.simulate()
    rho, prs, f = sim.get_density_pressure_force()   # of the state (pos, vel)
    p1, _ = sim.get_state()
    sim.close()

    lo = np.float32([8.0, 0.0, 12.0]); hi = np.float32([10.0, 1.2, 14.0])
    halo = np.float32(0.25)
    sub = np.flatnonzero(np.all((pos >= lo - halo) & (pos < hi + halo), axis=1))   # ascending id
    inner = np.all((pos[sub] >= lo) & (pos[sub] < hi), axis=1)
    assert inner.sum() > 5000 and len(sub) < 400_000
    o = CpuOracle(len(sub), boxDim=25.6, numCellsPerDim=256)
    o.set_state(pos[sub], vel[sub])
    orho, oprs, oK, oC = o.density()
    np.testing.assert_array_equal(K[sub][inner], oK[inner])
    np.testing.assert_array_equal(Cn[sub][inner], oC[inner])
    assert oK[inner].max() > 60                      # the dense regime is in the sample
    np.testing.assert_allclose(rho[sub][inner], orho[inner], rtol=2e-6, atol=0)
    np.testing.assert_allclose(prs[sub][inner], oprs[inner], rtol=0, atol=2e-6 * float(orho.max()))
    # forces from the full-domain density of the same particles (ours), oracle arithmetic
    of = o.forces(pos[sub], vel[sub], rho[sub], prs[sub])
    tol = force_tolerance(o, pos[sub], vel[sub], rho[sub], prs[sub], rel=RTOL_F)
    err = np.abs(f[sub] - of)
    assert np.all(err[inner] <= tol[inner]), f"force: worst excess {np.max(err[inner] / tol[inner]):.3g}x"
    o.step()
    np.testing.assert_allclose(p1[sub][inner], o.pos[inner], rtol=RTOL_POS, atol=2e-6)


def test_1m_random_developed_subdomain_matches_oracle():
    """BASELINE config 2 at its full size (1 M particles, glibc rand() seed 1, reference box),
    developed for 60 steps (the cloud has fallen 1.8 units: a dense layer on the floor under a
    sparse gas).  Same closed sub-problem argument as the 16 M test: a floor sub-domain plus a
    2.5 h halo re-solved by the oracle."""
    import ctypes
    n = 1_000_000
    ctypes.CDLL("libc.so.6").srand(1)
    s = sph.Settings(numParticles=n, randomInit=True)
    sim = sph.Simulator(s, record_force=True)
    sim.setup()
    sim.advance(60)
    pos, vel = sim.get_state()
    K, Cn = sim.get_neighbor_counts()
    sim.simulate()
    rho, prs, f = sim.get_density_pressure_force()
    p1, _ = sim.get_state()
    sim.close()

    lo = np.float32([4.0, 0.0, 4.0]); hi = np.float32([5.0, 0.8, 5.0])
    halo = np.float32(0.25)
    sub = np.flatnonzero(np.all((pos >= lo - halo) & (pos < hi + halo), axis=1))
    inner = np.all((pos[sub] >= lo) & (pos[sub] < hi), axis=1)
    assert inner.sum() > 2000 and len(sub) < 200_000
    o = CpuOracle(len(sub))
    o.set_state(pos[sub], vel[sub])
    orho, oprs, oK, oC = o.density()
    np.testing.assert_array_equal(K[sub][inner], oK[inner])
    np.testing.assert_array_equal(Cn[sub][inner], oC[inner])
    assert oK[inner].max() > 40                      # the dense floor layer is in the sample
    np.testing.assert_allclose(rho[sub][inner], orho[inner], rtol=RTOL_RHO, atol=0)
    np.testing.assert_allclose(prs[sub][inner], oprs[inner], rtol=0, atol=RTOL_RHO * float(orho.max()))
    of = o.forces(pos[sub], vel[sub], rho[sub], prs[sub])
    tol = force_tolerance(o, pos[sub], vel[sub], rho[sub], prs[sub], rel=RTOL_F)
    err = np.abs(f[sub] - of)
    assert np.all(err[inner] <= tol[inner]), f"force: worst excess {np.max(err[inner] / tol[inner]):.3g}x"
    o.step()
    np.testing.assert_allclose(p1[sub][inner], o.pos[inner], rtol=RTOL_POS, atol=2e-6)


def test_mask_handoff_is_bitwise_neutral():
    """Dense particles (C > 64) read density's in-range bit masks in the force kernel;
    pair order and arithmetic are unchanged, so results must be bit-identical to the
    path that repeats every distance test."""
    pos, vel = compressed_state(12000, seed=17)
    rng = np.random.default_rng(18)   # add a sparse halo so both paths run in one launch
    halo = rng.uniform(1.0, 9.0, size=(20000, 3)).astype(np.float32)
    pos = np.concatenate([pos, halo]); vel = np.concatenate([vel, np.zeros_like(halo)])
    out = []
    for handoff in (True, False):
        sim = sph.Simulator(sph.Settings(numParticles=len(pos)), record_force=True, mask_handoff=handoff)
        sim.setup()
        sim.set_state(pos, vel)
        sim.simulate()
        K, C = None, None
        rho, prs, f = sim.get_density_pressure_force()
        sim.advance(4)
        out.append((rho, f) + sim.get_state())
        sim.close()
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("per_cell", [30.0, 80.0, 400.0])
def test_tma_staged_tiles_are_bitwise_neutral(per_cell):
    """SphOptions.stage_tiles: dense CTAs take their candidates from shared-memory tiles filled
    by bulk TMA.  Same candidates in the same order => bit-identical density, force and state.
    per_cell 400 overflows the tile capacity for some CTAs (global-memory fallback in the tile
    kernel); the halo keeps non-qualifying CTAs in the same launch."""
    pos, vel = compressed_state(40000, seed=29, per_cell=per_cell)
    rng = np.random.default_rng(30)
    halo = rng.uniform(1.0, 9.0, size=(20000, 3)).astype(np.float32)
    pos = np.concatenate([pos, halo]); vel = np.concatenate([vel, np.zeros_like(halo)])
    out = []
    for staged in (False, True):
        sim = sph.Simulator(sph.Settings(numParticles=len(pos)), record_force=True, stage_tiles=staged)
        sim.setup()
        sim.set_state(pos, vel)
        sim.simulate()
        rho, prs, f = sim.get_density_pressure_force()
        sim.advance(4)
        out.append((rho, prs, f) + sim.get_state())
        sim.close()
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)


def test_pipelined_readback_hands_out_identical_positions():
    """SphOptions.pipeline_readback overlaps the D2H of step k with the computation of step
    k+1; every sph_step() must still return exactly the positions of the blocking mode."""
    pos, vel = random_state(60000, seed=23, lo=2.0, hi=6.0, vel_scale=1.5)
    seqs = []
    for pipe in (False, True):
        sim = sph.Simulator(sph.Settings(numParticles=len(pos)), pipeline_readback=pipe)
        sim.setup()
        sim.set_state(pos, vel)
        frames = []
        for _ in range(8):
            sim.simulate()
            frames.append(sim.getPosition().copy())
        seqs.append(frames)
        sim.close()
    for a, b in zip(*seqs):
        np.testing.assert_array_equal(a, b)
