"""GPU suite: the CUDA path AND the C oracle against the reference's own CUDA
implementation -- /root/reference/src/simulator.cu compiled unmodified into
oracle/_ref/libsph_ref.so (oracle/ref_harness.cu) -- on identical initial states.
This is the ground truth the north_star names; the prebuilt library travels to the
GPU box, the reference sources do not.
"""
import numpy as np
import pytest

import cudafluidsimulator_b200 as sph
from conftest import compressed_state, developed_state, force_tolerance, lattice_state, random_state
from oracle.oracle import REF_SO, CpuOracle, RefSim

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not REF_SO.exists(), reason="oracle/_ref/libsph_ref.so not built")]

STATES = {
    "lattice_sheet_10k": lambda: lattice_state(10000),
    "lattice_3d_40k": lambda: lattice_state(109 * 109 * 3 + 777),
    "random_30k": lambda: random_state(30000, seed=7),
    "random_moving_20k": lambda: random_state(20000, seed=8, lo=2.0, hi=5.0, vel_scale=2.0),
    "compressed_6k": lambda: compressed_state(6000),
    "developed_grid_8k_60": lambda: developed_state(8000, 60),
}


def test_reference_struct_sizes():
    import ctypes
    L = ctypes.CDLL(str(REF_SO))
    assert L.ref_sizeof_particle() == 56 and L.ref_sizeof_settings() == 32


@pytest.mark.parametrize("name", list(STATES))
def test_integers_bit_exact_vs_reference(name):
    pos, vel = STATES[name]()
    n = len(pos)
    ref = RefSim(n)
    ref.set_state(pos, vel)
    cells_r, flat_r = ref.keys()
    cnt = ref.neighbor_counts()
    ref.close()
    # the reference's list membership is its own flattened key
    np.testing.assert_array_equal(cnt["list_of"], flat_r)
    # oracle
    o = CpuOracle(n)
    cells, ff, fi, mo = o.keys(pos)
    np.testing.assert_array_equal(cells, cells_r)
    np.testing.assert_array_equal(ff, flat_r)
    rho, prs, K, C = o.density(pos)
    np.testing.assert_array_equal(K, cnt["K"])
    np.testing.assert_array_equal(C, cnt["C"])
    # CUDA path
    sim = sph.Simulator(sph.Settings(numParticles=n))
    sim.setup()
    sim.set_state(pos, vel)
    np.testing.assert_array_equal(sim.get_keys(sph.SPH_KEY_FLAT), flat_r.astype(np.uint32))
    Kg, Cg = sim.get_neighbor_counts()
    np.testing.assert_array_equal(Kg, cnt["K"])
    np.testing.assert_array_equal(Cg, cnt["C"])
    sim.close()


@pytest.mark.parametrize("name", list(STATES))
def test_single_step_vs_reference(name):
    pos, vel = STATES[name]()
    n = len(pos)
    ref = RefSim(n)
    ref.set_state(pos, vel)
    ref.step()
    r = ref.get_state()
    rpos_host = ref.positions()
    ref.close()
    np.testing.assert_array_equal(rpos_host, r["pos"])
    o = CpuOracle(n)
    tol = force_tolerance(o, pos, vel, r["rho"], r["prs"], rel=1e-5)

    def check(rho, prs, f, p1, who):
        # the reference sums in CAS-race order: density within 2e-6 relative
        np.testing.assert_allclose(rho, r["rho"], rtol=2e-6, atol=0, err_msg=who)
        np.testing.assert_allclose(prs, r["prs"], rtol=0, atol=2e-6 * float(r["rho"].max()), err_msg=who)
        err = np.abs(f - r["force"])
        assert np.all(err <= tol), f"{who}: force worst excess {np.max(err / tol):.3g}x"
        np.testing.assert_allclose(p1, r["pos"], rtol=1e-5, atol=1e-6, err_msg=who)

    o.set_state(pos, vel)
    o.step()
    check(o.rho, o.prs, o.force, o.pos, "oracle")
    sim = sph.Simulator(sph.Settings(numParticles=n), record_force=True)
    sim.setup()
    sim.set_state(pos, vel)
    sim.simulate()
    rho, prs, f = sim.get_density_pressure_force()
    check(rho, prs, f, sim.getPosition(), "cuda")
    sim.close()


@pytest.mark.parametrize("n,random_init", [(10000, False), (50000, True)])
def test_100_step_aggregates_vs_reference(n, random_init):
    """`./sph -n N -i grid|random -m time` semantics: setup() then 100 steps; compare
    kinetic energy and mean density (chaotic divergence => 2 % bound)."""
    import ctypes
    ref = RefSim(n, randomInit=random_init)     # its own setup(): rand() seed 1 / lattice
    p0 = ref.get_state()["pos"]
    for _ in range(100):
        ref.step()
    r = ref.get_state()
    ref.close()
    ke_r = 0.5 * 0.02 * float((r["vel"].astype(np.float64) ** 2).sum())
    rho_r = float(r["rho"].astype(np.float64).mean())
    ctypes.CDLL("libc.so.6").srand(1)
    sim = sph.Simulator(sph.Settings(numParticles=n, randomInit=random_init))
    sim.setup()
    np.testing.assert_array_equal(sim.get_state()[0], p0)   # identical initial state
    for _ in range(100):
        sim.simulate()
    ke, mrho = sim.get_stats()
    sim.close()
    assert abs(ke - ke_r) <= 0.02 * abs(ke_r) + 1e-9, (ke, ke_r)
    assert abs(mrho - rho_r) <= 0.02 * rho_r, (mrho, rho_r)


def test_mouse_push_vs_reference():
    """Click increments against the reference's own kernelMoveParticles, isolated by differencing a
    step with and a step without the click on both sides: 1e-5."""
    pos, vel = random_state(40000, seed=13, lo=3.0, hi=7.0)
    r = []
    for click in (False, True):
        ref = RefSim(len(pos))
        ref.set_state(pos, vel)
        if click:
            ref.step_click(400, 300)
        else:
            ref.step()
        r.append(ref.get_state()["vel"])
        ref.close()
    out = []
    for click in (False, True):
        sim = sph.Simulator(sph.Settings(numParticles=len(pos)))
        sim.setup()
        sim.set_state(pos, vel)
        sim.simulate()
        if click:
            sim.moveParticles((400, 300))
        out.append(sim.get_state()[1])
        sim.close()
    np.testing.assert_allclose(out[1] - out[0], r[1] - r[0], rtol=0, atol=1e-5)
    assert (np.abs(r[1] - r[0])[:, 2] > 4).sum() > 0


@pytest.mark.parametrize("morton", [False, True], ids=["index_sort", "z_index_sort"])
def test_reconstructed_sorted_variants_agree_with_reference(morton):
    """The benchmark comparators (reconstructed from README.md:5, source absent) must at least be
    the same physics: 20 steps of `-n 20000 -i random` against the real linked-list reference."""
    from oracle.oracle import RECON_SO, ReconSim
    if not RECON_SO.exists():
        pytest.skip("oracle/_ref/libsph_recon.so not built")
    n = 20000
    ref = RefSim(n, randomInit=True)
    rec = ReconSim(n, morton, randomInit=True)
    for _ in range(20):
        ref.step()
        rec.step_timed()
    np.testing.assert_allclose(rec.positions(), ref.positions(), rtol=2e-4, atol=2e-4)
    ref.close()
    rec.close()
