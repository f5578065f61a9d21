"""CPU suite: pins oracle/sph_oracle.c (the C restatement of ref src/simulator.cu)
against the closed-form facts of the reference and against golden vectors produced
by the reference's own CUDA build on a B200 (tests/golden/, scripts/make_golden.py).
"""
import struct

import numpy as np
import pytest

from conftest import GOLDEN, compressed_state, force_tolerance, lattice_state, random_state
from oracle.oracle import CpuOracle, morton3_np


def bits(x):
    return struct.unpack("I", struct.pack("f", float(x)))[0]


def test_host_constants_match_reference_main():
    # ref: main.cpp:57-61; values probed from the reference build
    o = CpuOracle(1)
    assert bits(o.s.v_kernel_coeff) == 0x4B5A90E6
    assert bits(o.s.d_kernel_coeff) == 0x4EBAC352
    assert bits(np.float32(0.1) * np.float32(0.1)) == 0x3C23D70B  # h2, ref: simulator.cu:89


def test_random_init_is_glibc_rand_seed1():
    # ref: simulator.cu:430-437; first two particles probed from the reference
    o = CpuOracle(4, randomInit=True)
    o.setup()
    np.testing.assert_array_equal(o.pos[0], np.float32([7.72150183, 4.15506363, 7.26479387]))
    np.testing.assert_array_equal(o.pos[1], np.float32([7.38752031, 8.29317856, 2.58041096]))
    assert o.pos.min() >= 1.0 and o.pos.max() <= 9.0


def test_grid_init_geometry():
    # ref: simulator.cu:438-453: 0.9h lattice, nx = 109, x outer / z inner
    o = CpuOracle(10000)
    assert o.setup() == 10000
    assert np.all(o.pos[:, 0] == np.float32(0.1))          # 109*109 >= 10000: one x-plane
    np.testing.assert_array_equal(o.pos[1], np.float32([0.1, 0.1, np.float32(0.1) + np.float32(0.09) * 1]))
    iy, iz = divmod(9999, 109)
    sp = np.float32(0.9) * np.float32(0.1)
    np.testing.assert_array_equal(o.pos[-1], np.float32([0.1, np.float32(0.1) + sp * iy, np.float32(0.1) + sp * iz]))
    big = CpuOracle(109 ** 3 + 5)
    assert big.setup() == 109 ** 3  # the reference silently stops here


def test_cell_coordinate_rounding():
    # true IEEE division: (boxDim-h)/h -> 98, h/h -> 1
    o = CpuOracle(2)
    hi = np.float32(10.0) - np.float32(0.1)
    cells, ff, fi, mo = o.keys(np.float32([[hi, 0.1, 0.1], [0.1, hi, 5.0]]))
    np.testing.assert_array_equal(cells, [[98, 1, 1], [1, 98, 50]])
    np.testing.assert_array_equal(ff, fi)
    np.testing.assert_array_equal(fi, [98 + 100 * (1 + 100 * 1), 1 + 100 * (98 + 100 * 50)])


@pytest.mark.parametrize("ncells,box", [(100, 10.0), (256, 25.6)])
def test_float_and_integer_flat_keys_agree_up_to_2_24(ncells, box):
    # ref: simulator.cu:78-82 evaluates the key in float; exact while n^3 <= 2^24
    rng = np.random.default_rng(3)
    pos = rng.uniform(0.0, box * 0.9999, size=(20000, 3)).astype(np.float32)
    o = CpuOracle(len(pos), boxDim=box, numCellsPerDim=ncells)
    cells, ff, fi, mo = o.keys(pos)
    np.testing.assert_array_equal(ff, fi)
    np.testing.assert_array_equal(mo, morton3_np(cells))
    assert cells.min() >= 0 and cells.max() < ncells


def test_float_keys_break_above_2_24():
    # SURVEY 7 "hard parts": the reference's float keys collide for n > 256
    o = CpuOracle(1, boxDim=51.2, numCellsPerDim=512)
    cells, ff, fi, mo = o.keys(np.float32([[51.05, 51.05, 51.05]]))
    assert fi[0] == 510 + 512 * (510 + 512 * 510)
    assert ff[0] != fi[0]


def test_sort_order_is_stable_counting_sort():
    rng = np.random.default_rng(4)
    keys = rng.integers(0, 1000, size=5000).astype(np.uint32)
    o = CpuOracle(1)
    order, start = o.sort_order(keys, 1000)
    np.testing.assert_array_equal(order, np.argsort(keys, kind="stable"))
    np.testing.assert_array_equal(start, np.searchsorted(np.sort(keys), np.arange(1001)))


def test_lone_and_lattice_density():
    # lone particle 31.33, pressure 0 everywhere at t=0
    o = CpuOracle(1)
    rho, prs, K, C = o.density(np.float32([[5, 5, 5]]))
    assert abs(rho[0] - 31.3336) < 1e-3 and prs[0] == 0 and K[0] == 1 and C[0] == 1
    pos, _ = lattice_state(109 * 109 * 3)
    o = CpuOracle(len(pos))
    rho, prs, K, C = o.density(pos)
    assert np.all(prs == 0) and rho.max() < 40
    assert K.max() == 7  # 6 lattice neighbours at 0.9h + self


def test_first_step_is_pure_free_fall():
    # F == 0 at t=0, so step 1 is v = g*dt, y += v*dt (bit exact)
    pos, vel = lattice_state(5000)
    o = CpuOracle(len(pos))
    o.set_state(pos, vel)
    o.step()
    assert np.all(o.force == 0)
    dt, g = np.float32(0.01), np.float32(-9.8)
    vy = np.float32(g * dt)
    expect_y = np.maximum(pos[:, 1] + vy * dt, np.float32(0.1))
    np.testing.assert_allclose(o.pos[:, 1], expect_y, rtol=0, atol=1e-7)
    np.testing.assert_array_equal(o.pos[:, [0, 2]], pos[:, [0, 2]])
    floor = pos[:, 1] + vy * dt < np.float32(0.1)
    np.testing.assert_array_equal(o.vel[floor, 1], np.float32(vy * np.float32(-0.5)))


def test_compressed_state_exercises_pressure_and_viscosity():
    pos, vel = compressed_state(6000)
    o = CpuOracle(len(pos))
    rho, prs, K, C = o.density(pos)
    assert (prs > 0).mean() > 0.5 and K.mean() > 50
    f = o.forces(pos, vel, rho, prs)
    scale = o.forces(pos, vel, rho, prs, abs_mode=True)
    assert np.all(np.abs(f) <= scale * (1 + 1e-5) + 1e-6)
    assert np.isfinite(f).all() and np.abs(f).max() > 0


def test_walls_and_velocity_floor():
    # ref: simulator.cu:279-314
    o = CpuOracle(3)
    pos = np.float32([[0.1, 5, 5], [9.9, 5, 5], [5, 5, 5]])
    vel = np.float32([[-3, 0, 0], [4, 0, 0], [0.00005, 0.098, 0]])
    o.set_state(pos, vel)
    o.step()
    assert o.pos[0, 0] == np.float32(0.1) and o.vel[0, 0] == np.float32(1.5)
    assert o.pos[1, 0] == np.float32(10.0) - np.float32(0.1) and o.vel[1, 0] == np.float32(-2.0)
    assert o.vel[2, 0] == 0.0          # |v| < 1e-4 -> 0
    assert abs(o.vel[2, 1]) < 1e-6 or o.vel[2, 1] == 0.0


def test_push_matches_reference_semantics():
    # ref: simulator.cu:329-367: centre column gets -z, ring gets +-x/y by 5/dx, 5/dy
    o = CpuOracle(3)
    # click at window centre -> x=y=5 -> cell (50, 100-50=50); place particles around it
    pos = np.float32([[5.05, 5.05, 3.05], [5.15, 5.05, 3.05], [5.05, 4.85, 3.05]])
    o.set_state(pos)
    o.push(pos, 400, 300)
    np.testing.assert_array_equal(o.vel[0], [0, 0, -5])
    np.testing.assert_array_equal(o.vel[1], [5, 0, 0])
    np.testing.assert_array_equal(o.vel[2], [0, -2.5, 0])


# ---- golden vectors from the reference's own CUDA build (generated on a B200) -----
GOLDEN_FILES = sorted(GOLDEN.glob("ref_*.npz"))


@pytest.mark.skipif(not GOLDEN_FILES, reason="golden vectors not generated yet")
@pytest.mark.parametrize("path", GOLDEN_FILES, ids=lambda p: p.stem)
def test_oracle_against_reference_golden(path):
    g = np.load(path)
    cfg = {k: float(g[k]) for k in ("h", "boxDim", "numCellsPerDim", "timestep")}
    pos, vel = g["pos0"], g["vel0"]
    o = CpuOracle(len(pos), **cfg)
    # integers: bit exact
    cells, ff, fi, mo = o.keys(pos)
    np.testing.assert_array_equal(cells, g["cells"])
    np.testing.assert_array_equal(ff, g["flat"])
    np.testing.assert_array_equal(fi, g["list_of"])       # list membership == cell key
    rho, prs, K, C = o.density(pos)
    np.testing.assert_array_equal(K, g["K"])
    np.testing.assert_array_equal(C, g["C"])
    # one step: density / pressure / force / position / velocity
    o.set_state(pos, vel)
    o.step()
    np.testing.assert_allclose(o.rho, g["rho1"], rtol=2e-6, atol=0)
    np.testing.assert_allclose(o.prs, g["prs1"], rtol=0, atol=2e-6 * np.maximum(g["rho1"], 1.0).max())
    tol = force_tolerance(o, pos, vel, g["rho1"], g["prs1"])
    assert np.all(np.abs(o.force - g["force1"]) <= tol)
    np.testing.assert_allclose(o.pos, g["pos1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(o.vel, g["vel1"], rtol=1e-5, atol=1e-4)
    # aggregates after `steps` steps (trajectories diverge chaotically: loose bound)
    steps = int(g["steps"])
    for _ in range(steps - 1):
        o.step()
    ke, mrho = o.stats()
    assert abs(ke - float(g["ke"])) <= 0.02 * max(abs(float(g["ke"])), 1e-9)
    assert abs(mrho - float(g["mean_rho"])) <= 0.02 * float(g["mean_rho"])
