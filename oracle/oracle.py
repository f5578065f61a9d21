"""ctypes access to the two checkers -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

    CpuOracle   oracle/liboracle.so        serial C restatement (sph_oracle.c)
    RefSim      oracle/_ref/libsph_ref.so  the unmodified reference CUDA build
                                           (ref_harness.cu; needs a GPU to run)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "liboracle.so"
REF_SO = HERE / "_ref" / "libsph_ref.so"
RECON_SO = HERE / "_ref" / "libsph_recon.so"

_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int32)
_U = C.POINTER(C.c_uint32)


class OracleSettings(C.Structure):
    _fields_ = [("randomInit", C.c_int), ("numParticles", C.c_int), ("h", C.c_float),
                ("v_kernel_coeff", C.c_float), ("d_kernel_coeff", C.c_float),
                ("boxDim", C.c_float), ("numCellsPerDim", C.c_float), ("timestep", C.c_float)]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _ensure_oracle_built():
    src = HERE / "sph_oracle.c"
    if not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "--no-print-directory", "liboracle.so"], check=True)


class CpuOracle:
    """Serial CPU restatement of the reference step."""

    def __init__(self, n, *, h=0.1, boxDim=10.0, numCellsPerDim=100.0, timestep=0.01,
                 randomInit=False, vk=None, dk=None):
        _ensure_oracle_built()
        L = C.CDLL(str(ORACLE_SO))
        self.L = L
        L.oracle_constants.argtypes = [C.c_float, _F, _F]
        L.oracle_init.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, _F]
        L.oracle_init.restype = C.c_int
        L.oracle_keys.argtypes = [_F, C.c_int, C.c_float, C.c_float, _I, _I, _I, _U]
        L.oracle_sort_order.argtypes = [_U, C.c_int, C.c_uint32, _I, _I]
        L.oracle_sort_order.restype = C.c_int
        sp = C.POINTER(OracleSettings)
        L.oracle_density.argtypes = [_F, C.c_int, sp, _F, _F, _I, _I]
        L.oracle_forces.argtypes = [_F, _F, _F, _F, C.c_int, sp, _F, C.c_int]
        L.oracle_integrate.argtypes = [_F, _F, _F, _F, C.c_int, sp]
        L.oracle_step.argtypes = [_F, _F, _F, _F, _F, C.c_int, sp]
        L.oracle_push.argtypes = [_F, _F, C.c_int, sp, C.c_int, C.c_int]
        L.oracle_stats.argtypes = [_F, _F, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        for f in ("oracle_density", "oracle_forces", "oracle_step", "oracle_push"):
            getattr(L, f).restype = C.c_int
        if vk is None or dk is None:
            a, b = C.c_float(), C.c_float()
            L.oracle_constants(C.c_float(h), C.byref(a), C.byref(b))
            vk = a.value if vk is None else vk
            dk = b.value if dk is None else dk
        self.s = OracleSettings(int(randomInit), int(n), h, vk, dk, boxDim, numCellsPerDim, timestep)
        self.n = int(n)
        self.pos = np.zeros((self.n, 3), np.float32)
        self.vel = np.zeros((self.n, 3), np.float32)
        self.force = np.zeros((self.n, 3), np.float32)
        self.rho = np.zeros(self.n, np.float32)
        self.prs = np.zeros(self.n, np.float32)

    # -- reference surface --------------------------------------------------------
    def setup(self):
        """ref: simulator.cu:430-453 initial positions (velocities zero)."""
        got = self.L.oracle_init(self.s.randomInit, self.n, self.s.h, self.s.boxDim,
                                 _p(self.pos, C.c_float))
        self.vel[:] = 0
        return got

    def set_state(self, pos, vel=None):
        self.pos[:] = np.asarray(pos, np.float32)
        self.vel[:] = 0 if vel is None else np.asarray(vel, np.float32)

    def step(self):
        rc = self.L.oracle_step(_p(self.pos, C.c_float), _p(self.vel, C.c_float),
                                _p(self.force, C.c_float), _p(self.rho, C.c_float),
                                _p(self.prs, C.c_float), self.n, C.byref(self.s))
        if rc:
            raise RuntimeError(f"oracle_step failed: {rc}")

    def push(self, bin_pos, x, y):
        bp = np.ascontiguousarray(bin_pos, np.float32)
        rc = self.L.oracle_push(_p(bp, C.c_float), _p(self.vel, C.c_float), self.n,
                                C.byref(self.s), int(x), int(y))
        if rc:
            raise RuntimeError(f"oracle_push failed: {rc}")

    # -- pieces ---------------------------------------------------------------------
    def keys(self, pos=None):
        pos = self.pos if pos is None else np.ascontiguousarray(pos, np.float32)
        n = len(pos)
        cells = np.empty((n, 3), np.int32)
        ff = np.empty(n, np.int32)
        fi = np.empty(n, np.int32)
        mo = np.empty(n, np.uint32)
        self.L.oracle_keys(_p(pos, C.c_float), n, self.s.h, self.s.numCellsPerDim,
                           _p(cells, C.c_int32), _p(ff, C.c_int32), _p(fi, C.c_int32),
                           _p(mo, C.c_uint32))
        return cells, ff, fi, mo

    def sort_order(self, keys, nkeys):
        keys = np.ascontiguousarray(keys, np.uint32)
        order = np.empty(len(keys), np.int32)
        start = np.empty(int(nkeys) + 1, np.int32)
        rc = self.L.oracle_sort_order(_p(keys, C.c_uint32), len(keys), int(nkeys),
                                      _p(order, C.c_int32), _p(start, C.c_int32))
        if rc:
            raise RuntimeError(f"oracle_sort_order failed: {rc}")
        return order, start

    def density(self, pos=None, counts=True):
        pos = self.pos if pos is None else np.ascontiguousarray(pos, np.float32)
        n = len(pos)
        rho = np.empty(n, np.float32)
        prs = np.empty(n, np.float32)
        K = np.empty(n, np.int32) if counts else None
        Cn = np.empty(n, np.int32) if counts else None
        s = OracleSettings.from_buffer_copy(self.s)
        s.numParticles = n
        rc = self.L.oracle_density(_p(pos, C.c_float), n, C.byref(s), _p(rho, C.c_float),
                                   _p(prs, C.c_float), _p(K, C.c_int32), _p(Cn, C.c_int32))
        if rc:
            raise RuntimeError(f"oracle_density failed: {rc}")
        return rho, prs, K, Cn

    def forces(self, pos, vel, rho, prs, abs_mode=False):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        rho = np.ascontiguousarray(rho, np.float32)
        prs = np.ascontiguousarray(prs, np.float32)
        n = len(pos)
        f = np.empty((n, 3), np.float32)
        s = OracleSettings.from_buffer_copy(self.s)
        s.numParticles = n
        rc = self.L.oracle_forces(_p(pos, C.c_float), _p(vel, C.c_float), _p(rho, C.c_float),
                                  _p(prs, C.c_float), n, C.byref(s), _p(f, C.c_float), int(abs_mode))
        if rc:
            raise RuntimeError(f"oracle_forces failed: {rc}")
        return f

    def stats(self):
        ke, mr = C.c_double(), C.c_double()
        self.L.oracle_stats(_p(self.vel, C.c_float), _p(self.rho, C.c_float), self.n,
                            C.byref(ke), C.byref(mr))
        return ke.value, mr.value


def morton3_np(cells):
    """numpy Morton interleave (x bit 0) used to cross-check the C oracle."""
    def spread(v):
        v = v.astype(np.uint32) & 0x3FF
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    c = np.asarray(cells)
    return spread(c[:, 0]) | (spread(c[:, 1]) << 1) | (spread(c[:, 2]) << 2)


class RefSim:
    """The reference's own CUDA implementation (unmodified simulator.cu) behind the
    headless harness.  Needs a GPU; raises if the prebuilt library is absent."""

    def __init__(self, n, *, h=0.1, boxDim=10.0, numCellsPerDim=100.0, timestep=0.01,
                 randomInit=False, vk=None, dk=None):
        if not REF_SO.exists():
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle` where /root/reference is mounted")
        L = C.CDLL(str(REF_SO))
        self.L = L
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_int] + [C.c_float] * 6
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_state.argtypes = [C.c_void_p, _F, _F]
        L.ref_get_state.argtypes = [C.c_void_p, _F, _F, _F, _F, _F]
        L.ref_step.argtypes = [C.c_void_p]
        L.ref_step_click.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_step_timed.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ref_positions.restype = _F
        L.ref_positions.argtypes = [C.c_void_p]
        L.ref_keys.argtypes = [C.c_void_p, _I, _I]
        L.ref_neighbor_counts.argtypes = [C.c_void_p, _I, _I, _I, _I]
        if vk is None or dk is None:
            hh = np.float32(h)
            pi = np.float32(3.14159265)
            vk = float(np.float32(45.0) / (pi * np.float32(math.pow(float(hh), 6)))) if vk is None else vk
            dk = float(np.float32(315.0) / (np.float32(64.0) * pi * np.float32(math.pow(float(hh), 9)))) if dk is None else dk
        self.n = int(n)
        self.h = L.ref_create(int(randomInit), self.n, h, vk, dk, boxDim, numCellsPerDim, timestep)
        if not self.h:
            raise RuntimeError("ref_create failed (no GPU?)")
        self.buckets = (C.c_double * 3)(0, 0, 0)
        self.iters = C.c_int(0)

    def _chk(self, rc, what):
        if rc:
            raise RuntimeError(f"{what} failed with CUDA error {rc}")

    def set_state(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32) if vel is not None else None
        self._chk(self.L.ref_set_state(self.h, _p(pos, C.c_float), _p(vel, C.c_float)), "ref_set_state")

    def get_state(self):
        n = self.n
        out = [np.empty((n, 3), np.float32) for _ in range(3)] + [np.empty(n, np.float32) for _ in range(2)]
        self._chk(self.L.ref_get_state(self.h, *[_p(a, C.c_float) for a in out]), "ref_get_state")
        return dict(zip(("pos", "vel", "force", "rho", "prs"), out))

    def step(self):
        self._chk(self.L.ref_step(self.h), "ref_step")

    def step_click(self, x, y):
        self._chk(self.L.ref_step_click(self.h, int(x), int(y)), "ref_step_click")

    def step_timed(self):
        self._chk(self.L.ref_step_timed(self.h, self.buckets, C.byref(self.iters)), "ref_step_timed")
        return tuple(self.buckets), self.iters.value

    def positions(self):
        p = self.L.ref_positions(self.h)
        return np.ctypeslib.as_array(p, shape=(self.n, 3)).copy()

    def keys(self):
        cells = np.empty((self.n, 3), np.int32)
        flat = np.empty(self.n, np.int32)
        self._chk(self.L.ref_keys(self.h, _p(cells, C.c_int32), _p(flat, C.c_int32)), "ref_keys")
        return cells, flat

    def neighbor_counts(self):
        a = [np.empty(self.n, np.int32) for _ in range(4)]
        self._chk(self.L.ref_neighbor_counts(self.h, *[_p(x, C.c_int32) for x in a]), "ref_neighbor_counts")
        return dict(zip(("list_of", "K", "C", "Knz"), a))

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None


class ReconSim:
    """RECONSTRUCTED `index_sort` (morton=False) / `z_index_sort` (morton=True) variants of the
    reference (oracle/recon_variants.cu; source absent from the reference tree -- reconstructed
    from README.md:5, parity unpinned).  Benchmark comparator only."""

    def __init__(self, n, morton, *, h=0.1, boxDim=10.0, numCellsPerDim=100.0, timestep=0.01,
                 randomInit=False):
        if not RECON_SO.exists():
            raise FileNotFoundError(f"{RECON_SO} missing: run `make -C oracle` where /root/reference is mounted")
        L = C.CDLL(str(RECON_SO))
        self.L = L
        L.recon_create.restype = C.c_void_p
        L.recon_create.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_float] * 6
        L.recon_step_timed.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.recon_positions.restype = _F
        L.recon_positions.argtypes = [C.c_void_p]
        L.recon_destroy.argtypes = [C.c_void_p]
        hh, pi = np.float32(h), np.float32(3.14159265)
        vk = float(np.float32(45.0) / (pi * np.float32(math.pow(float(hh), 6))))
        dk = float(np.float32(315.0) / (np.float32(64.0) * pi * np.float32(math.pow(float(hh), 9))))
        self.n = int(n)
        self.h = L.recon_create(int(morton), int(randomInit), self.n, h, vk, dk, boxDim, numCellsPerDim, timestep)
        if not self.h:
            raise RuntimeError("recon_create failed (no GPU?)")
        self.buckets = (C.c_double * 3)(0, 0, 0)

    def step_timed(self):
        rc = self.L.recon_step_timed(self.h, self.buckets)
        if rc:
            raise RuntimeError(f"recon_step_timed failed with CUDA error {rc}")
        return tuple(self.buckets)

    def positions(self):
        return np.ctypeslib.as_array(self.L.recon_positions(self.h), shape=(self.n, 3)).copy()

    def close(self):
        if self.h:
            self.L.recon_destroy(self.h)
            self.h = None
