"""Python mirror of the reference interface (ref: src/simulator.h:19-74, src/times.h)
over the C ABI -- same names, argument meaning and call order as `class Simulator`,
so parity tests read like the reference's callers (main.cpp:62-76, display.cpp:35-64).

All compute happens in libsph_b200.so on the GPU; numpy is only the container for
arrays crossing the boundary.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _native as N
from ._native import SPH_KEY_FLAT, SPH_KEY_MORTON, SphError  # noqa: F401  (re-export)

PI = np.float32(3.14159265)  # ref: simulator.h:6


def kernel_coefficients(h: float):
    """v_kernel_coeff, d_kernel_coeff exactly as main.cpp:57-61 computes them
    (pow(float,int) is the double overload, rounded back to float)."""
    h = np.float32(h)
    h6 = np.float32(math.pow(float(h), 6))
    h9 = np.float32(math.pow(float(h), 9))
    vk = np.float32(45.0) / (PI * h6)
    dk = np.float32(315.0) / (np.float32(64.0) * PI * h9)
    return np.float32(vk), np.float32(dk)


@dataclass
class Settings:
    """ref: simulator.h:19-31.  Defaults are main.cpp:57-63."""
    randomInit: bool = False
    numParticles: int = 1000
    h: float = 0.1
    v_kernel_coeff: float | None = None
    d_kernel_coeff: float | None = None
    boxDim: float = 10.0
    numCellsPerDim: float = 100.0
    timestep: float = 0.01

    def __post_init__(self):
        if self.v_kernel_coeff is None or self.d_kernel_coeff is None:
            vk, dk = kernel_coefficients(self.h)
            self.v_kernel_coeff = float(vk) if self.v_kernel_coeff is None else self.v_kernel_coeff
            self.d_kernel_coeff = float(dk) if self.d_kernel_coeff is None else self.d_kernel_coeff

    def to_c(self) -> N.SphSettings:
        s = N.SphSettings()
        s.randomInit = 1 if self.randomInit else 0
        s.numParticles = int(self.numParticles)
        s.h = self.h
        s.v_kernel_coeff = self.v_kernel_coeff
        s.d_kernel_coeff = self.d_kernel_coeff
        s.boxDim = self.boxDim
        s.numCellsPerDim = self.numCellsPerDim
        s.timestep = self.timestep
        return s


@dataclass
class Times:
    """ref: times.h:5-10"""
    buildGrid: float = 0.0
    sphUpdate: float = 0.0
    memcpy: float = 0.0
    iters: int = 0


def _f32(a, n):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.shape != (n, 3):
        raise ValueError(f"expected a ({n}, 3) float32 array, got {a.shape}")
    return a


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


class Simulator:
    """ref: simulator.h:53-74.  Construction only records the settings
    (simulator.cu:370-375); `setup()` allocates and initialises."""

    def __init__(self, settings: Settings, *, key_mode: int = SPH_KEY_FLAT, device: int = 0,
                 record_force: bool = False, use_graph: bool = True, mask_handoff: bool = True,
                 pipeline_readback: bool = False, stage_tiles: bool = False, density_sum: int = 0,
                 sort_algo: int = 0):
        self.settings = settings
        self._lib = N.load()
        opt = N.SphOptions()
        opt.device = device
        opt.key_mode = key_mode
        opt.record_force = 1 if record_force else 0
        opt.use_graph = 1 if use_graph else 2
        opt.no_mask_handoff = 0 if mask_handoff else 1
        opt.pipeline_readback = 1 if pipeline_readback else 0
        opt.stage_tiles = 1 if stage_tiles else 0
        opt.density_sum = int(density_sum)
        opt.sort_algo = int(sort_algo)   # 0 default, 1 counting sort by cell, 2 radix passes
        h = C.c_void_p()
        cs = settings.to_c()
        N.check(self._lib.sph_create_ex(C.byref(cs), C.byref(opt), C.byref(h)))
        self._h = h
        self.n = int(settings.numParticles)

    # -- reference surface -------------------------------------------------------
    def setup(self) -> None:
        N.check(self._lib.sph_setup(self._h))

    def getPosition(self) -> np.ndarray:
        """Host positions, (N,3) float32 in ORIGINAL particle order; a view of the
        simulator-owned pinned buffer, refreshed by every simulate()."""
        p = self._lib.sph_positions_host(self._h)
        if not p or self.n == 0:
            return np.zeros((0, 3), np.float32)
        return np.ctypeslib.as_array(p, shape=(self.n, 3))

    def simulate(self) -> None:
        N.check(self._lib.sph_step(self._h))

    def simulateAndTime(self, times: Times) -> None:
        t = N.SphTimes(times.buildGrid, times.sphUpdate, times.memcpy, times.iters)
        N.check(self._lib.sph_step_timed(self._h, C.byref(t)))
        times.buildGrid, times.sphUpdate, times.memcpy, times.iters = (
            t.buildGrid, t.sphUpdate, t.memcpy, t.iters)

    def moveParticles(self, mouse_pos) -> None:
        """Declared but never defined in the reference (simulator.h:73); here it
        applies the mouse push of simulator.cu:329-367 to the step that just ran."""
        N.check(self._lib.sph_push(self._h, int(mouse_pos[0]), int(mouse_pos[1])))

    # -- additive: device-resident stepping, state and parity hooks -------------------
    def advance(self, steps: int) -> None:
        N.check(self._lib.sph_advance(self._h, int(steps)))

    def advance_timed(self, steps: int) -> float:
        """Device milliseconds (CUDA events on the simulator's stream) for `steps` steps."""
        ms = C.c_float()
        N.check(self._lib.sph_advance_timed(self._h, int(steps), C.byref(ms)))
        return float(ms.value)

    def readback(self) -> np.ndarray:
        N.check(self._lib.sph_readback(self._h))
        return self.getPosition()

    def set_state(self, pos, vel=None) -> None:
        pos = _f32(pos, self.n)
        vel = _f32(vel, self.n) if vel is not None else None
        N.check(self._lib.sph_set_state(self._h, _ptr(pos, C.c_float), _ptr(vel, C.c_float)))

    def get_state(self):
        pos = np.empty((self.n, 3), np.float32)
        vel = np.empty((self.n, 3), np.float32)
        N.check(self._lib.sph_get_state(self._h, _ptr(pos, C.c_float), _ptr(vel, C.c_float)))
        return pos, vel

    def get_keys(self, key_mode: int = SPH_KEY_FLAT) -> np.ndarray:
        k = np.empty(self.n, np.uint32)
        N.check(self._lib.sph_get_keys(self._h, key_mode, _ptr(k, C.c_uint32)))
        return k

    def get_sorted_index(self):
        ids = np.empty(self.n, np.uint32)
        keys = np.empty(self.n, np.uint32)
        N.check(self._lib.sph_get_sorted_index(self._h, _ptr(ids, C.c_uint32), _ptr(keys, C.c_uint32)))
        return ids, keys

    @property
    def table_size(self) -> int:
        """Number of distinct cell keys of the sort (flat: cells; Morton: 8^bits)."""
        size = C.c_uint32()
        N.check(self._lib.sph_get_cell_start(self._h, None, C.byref(size)))
        return int(size.value)

    def get_cell_start(self) -> np.ndarray:
        size = C.c_uint32()
        N.check(self._lib.sph_get_cell_start(self._h, None, C.byref(size)))
        start = np.empty(size.value + 1, np.uint32)
        N.check(self._lib.sph_get_cell_start(self._h, _ptr(start, C.c_uint32), C.byref(size)))
        return start

    def get_neighbor_counts(self):
        K = np.empty(self.n, np.int32)
        Cn = np.empty(self.n, np.int32)
        N.check(self._lib.sph_get_neighbor_counts(self._h, _ptr(K, C.c_int32), _ptr(Cn, C.c_int32)))
        return K, Cn

    def get_density_pressure_force(self, force: bool = True):
        rho = np.empty(self.n, np.float32)
        prs = np.empty(self.n, np.float32)
        f = np.empty((self.n, 3), np.float32) if force else None
        N.check(self._lib.sph_get_density_pressure_force(
            self._h, _ptr(rho, C.c_float), _ptr(prs, C.c_float), _ptr(f, C.c_float)))
        return rho, prs, f

    def get_stats(self):
        ke, mr = C.c_double(), C.c_double()
        N.check(self._lib.sph_get_stats(self._h, C.byref(ke), C.byref(mr)))
        return ke.value, mr.value

    def debug_flags(self):
        """(violation bits, is_checked_build) of the self-checking build."""
        f, c = C.c_uint32(), C.c_int()
        N.check(self._lib.sph_debug_flags(self._h, C.byref(f), C.byref(c)))
        return int(f.value), bool(c.value)

    def profile_enable(self, on: bool = True) -> None:
        N.check(self._lib.sph_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset: bool = False):
        ms = (C.c_double * N.SPH_STAGE_COUNT)()
        ln = (C.c_int64 * N.SPH_STAGE_COUNT)()
        N.check(self._lib.sph_profile_read(self._h, ms, ln, 1 if reset else 0))
        names = [self._lib.sph_stage_name(i).decode() for i in range(N.SPH_STAGE_COUNT)]
        return {names[i]: {"ms": ms[i], "launches": ln[i]} for i in range(N.SPH_STAGE_COUNT)}

    @property
    def launch_count(self) -> int:
        return int(self._lib.sph_launch_count(self._h))

    def sort_info(self) -> dict:
        """{"algo": "count" | "radix", "kernels": sort kernels per step, "count_fused", "radix_passes"}"""
        v = [C.c_int32() for _ in range(4)]
        N.check(self._lib.sph_sort_info(self._h, *[C.byref(x) for x in v]))
        return {"algo": "count" if v[0].value == N.SPH_SORT_COUNT else "radix", "kernels": v[1].value,
                "count_fused": bool(v[2].value), "radix_passes": v[3].value}

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.sph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
