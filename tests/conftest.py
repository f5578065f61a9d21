import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """The product library is built in-tree; CPU tests only load it / check symbols."""
    from cudafluidsimulator_b200 import build
    build.build_library()
    return True


# ---- seeded states shared by the CPU and GPU suites -------------------------------
def lattice_state(n, h=0.1, box=10.0):
    """The reference's grid init (ref: simulator.cu:438-453) through the oracle."""
    from oracle.oracle import CpuOracle
    o = CpuOracle(n, h=h, boxDim=box, numCellsPerDim=round(box / h))
    o.setup()
    return o.pos.copy(), np.zeros_like(o.pos)


def random_state(n, seed=0, lo=1.0, hi=9.0, vel_scale=0.0):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * vel_scale).astype(np.float32)
    return pos, vel


def compressed_state(n, seed=1, per_cell=80.0, origin=(2.0, 0.1, 3.0), vel_scale=1.0, h=0.1):
    """Dense blob (~per_cell particles per h^3 cell; rho > 1000 needs >= ~50) with random
    velocities so that pressure > 0 and viscosity are exercised."""
    rng = np.random.default_rng(seed)
    edge = h * (n / per_cell) ** (1.0 / 3.0)
    pos = (np.asarray(origin, np.float32) + rng.uniform(0, edge, size=(n, 3))).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * vel_scale).astype(np.float32)
    return pos, vel


def developed_state(n, steps, random_init=False):
    """State after `steps` oracle steps from the reference's own initial condition."""
    from oracle.oracle import CpuOracle
    o = CpuOracle(n, randomInit=random_init)
    o.setup()
    for _ in range(steps):
        o.step()
    return o.pos.copy(), o.vel.copy()


def force_tolerance(oracle, pos, vel, rho, prs, rel=1e-5):
    """|dF| <= rel * max(|F|, sum|terms|) per component + a floor."""
    scale = oracle.forces(pos, vel, rho, prs, abs_mode=True)
    return rel * np.maximum(scale, 1e-3)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
