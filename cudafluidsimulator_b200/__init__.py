"""cudafluidsimulator_b200 -- B200-native (sm_100a) SPH timestep, drop-in behind the
reference's simulator.h interface.  See DESIGN.md and INTEGRATION.md.

The product is libsph_b200.so (hand-written CUDA + C ABI, include/sph_b200.h) with a
C++ `class Simulator` / `./sph` CLI on top (include/simulator.h, host/).  This
Python package is the ctypes mirror of that interface used by tests and bench.py.
"""
from ._native import (SPH_KEY_FLAT, SPH_KEY_MORTON, SPH_SORT_COUNT, SPH_SORT_RADIX,  # noqa: F401
                      SphError, load)
from .simulator import Settings, Simulator, Times, kernel_coefficients  # noqa: F401

__all__ = ["Settings", "Simulator", "Times", "kernel_coefficients", "SPH_KEY_FLAT",
           "SPH_KEY_MORTON", "SPH_SORT_COUNT", "SPH_SORT_RADIX", "SphError", "load"]
