"""CPU suite: the C-ABI library loads, exports every symbol include/sph_b200.h
declares, keeps the reference's struct layouts, validates settings, and fails loudly
(no CPU fallback) when there is no GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

from conftest import ROOT, has_gpu
import cudafluidsimulator_b200 as sph
from cudafluidsimulator_b200 import _native as N


def declared_symbols():
    text = (ROOT / "include" / "sph_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = N.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in sph_b200.h but not exported"
    assert set(names) == set(N.SYMBOLS), "ctypes table and header disagree"
    assert lib.sph_abi_version() == 1


def test_struct_layouts_match_reference():
    # ref: simulator.h:19-31 (sizeof Settings == 32), times.h:5-10 (sizeof Times == 32)
    assert C.sizeof(N.SphSettings) == 32
    assert N.SphSettings.numParticles.offset == 4 and N.SphSettings.h.offset == 8
    assert N.SphSettings.timestep.offset == 28
    assert C.sizeof(N.SphTimes) == 32 and N.SphTimes.iters.offset == 24
    assert C.sizeof(N.SphOptions) == 64


def test_settings_defaults_are_reference_main():
    s = sph.Settings()
    assert (s.numParticles, s.randomInit, s.boxDim, s.numCellsPerDim, s.timestep) == (1000, False, 10.0, 100.0, 0.01)
    assert s.v_kernel_coeff == 14323942.0 and s.d_kernel_coeff == 1566681344.0


@pytest.mark.parametrize("kw", [dict(numParticles=-1), dict(h=0.0), dict(numCellsPerDim=0.0),
                                dict(numCellsPerDim=100.5), dict(numCellsPerDim=2048.0),
                                dict(boxDim=-1.0), dict(timestep=0.0)])
def test_create_rejects_bad_settings(kw):
    with pytest.raises(sph.SphError) as e:
        sph.Simulator(sph.Settings(**kw))
    assert e.value.code == -1


def test_create_rejects_bad_key_mode():
    with pytest.raises(sph.SphError):
        sph.Simulator(sph.Settings(), key_mode=7)


def test_calls_before_setup_are_state_errors():
    sim = sph.Simulator(sph.Settings(numParticles=8))
    with pytest.raises(sph.SphError) as e:
        sim.simulate()
    assert e.value.code == -2
    sim.close()


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu():
    sim = sph.Simulator(sph.Settings(numParticles=8))
    with pytest.raises(sph.SphError) as e:
        sim.setup()
    assert e.value.code > 0  # a cudaError_t, not a silent CPU path
    sim.close()


def test_product_does_not_use_the_oracle():
    """The product path may not import, link, dlopen or call anything under oracle/
    (building the checkers from build.py is not using them)."""
    banned = ("liboracle", "libsph_ref", "import oracle", "from oracle", "oracle_", "ref_harness")
    for path in (ROOT / "cudafluidsimulator_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".cpp", ".h"}:
            text = path.read_text()
            for b in banned:
                assert b not in text, f"{path} mentions {b}"
    for path in (ROOT / "include").glob("*.h"):
        assert "oracle" not in path.read_text().lower(), path
