"""ctypes caller of the multi-GPU cluster entry points (include/sph_b200.h, sph_cluster_*).

The whole slab protocol -- z-slab decomposition, per-step ghost halo exchange, particle
migration, load rebalancing -- lives in the library (csrc/sph_cluster.cu); this module only
marshals arguments.  Two ways to run:

  * one process drives every slab (`Cluster(settings, world=W, devices=[...])`): messages move
    with peer-to-peer copies; several slabs may share one GPU (tests on a single-GPU box);
  * one process per GPU (torchrun): `Cluster(settings, world=W, rank=r, devices=[local_rank],
    nccl_id=...)`, messages move with ncclSend / ncclRecv.  `nccl_id()` on rank 0 creates the
    id, the caller broadcasts its bytes (e.g. over a gloo group).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def nccl_id() -> bytes:
    buf = (C.c_uint8 * N.SPH_NCCL_ID_BYTES)()
    N.check(N.load().sph_cluster_nccl_id(buf))
    return bytes(buf)


def slab_ranges(nz: int, world: int):
    """Initial layer ranges [(zlo, zhi)] the library gives the slabs (near-equal, contiguous)."""
    base, extra = divmod(nz, world)
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((z, z + n))
        z += n
    return out


def partition(pos, h, ranges):
    """Indices of the particles each slab owns: global z cell (IEEE divide, truncate --
    ref: simulator.cu:69) in [zlo, zhi)."""
    cz = (np.asarray(pos, np.float32)[:, 2] / np.float32(h)).astype(np.int64)
    return [np.nonzero((cz >= lo) & (cz < hi))[0] for lo, hi in ranges]


class Cluster:
    def __init__(self, settings, world: int, devices, rank: int = 0, nz_cells: int = 0,
                 capacity: int = 0, ghost_capacity: int = 0, emig_capacity: int = 0,
                 density_sum: int = 0, rebalance_every: int = 0, nccl_id: bytes | None = None):
        self._lib = N.load()
        devices = list(devices)
        o = N.SphClusterOptions()
        o.world, o.first_rank, o.local_count = int(world), int(rank), len(devices)
        for i, d in enumerate(devices):
            o.devices[i] = int(d)
        o.nz_cells, o.capacity = int(nz_cells), int(capacity)
        o.ghost_capacity, o.emig_capacity = int(ghost_capacity), int(emig_capacity)
        o.density_sum, o.rebalance_every = int(density_sum), int(rebalance_every)
        if nccl_id is not None:
            C.memmove(o.nccl_id, nccl_id, N.SPH_NCCL_ID_BYTES)
        cs = settings.to_c()
        h = C.c_void_p()
        N.check(self._lib.sph_cluster_create(C.byref(cs), C.byref(o), C.byref(h)))
        self._h = h
        self.world, self.rank, self.local_count = int(world), int(rank), len(devices)
        self.settings = settings

    # -- particles ------------------------------------------------------------------------
    def setup(self) -> None:
        """The reference's initialisation of the whole box, split over the slabs."""
        N.check(self._lib.sph_cluster_setup(self._h))

    def load(self, local_index: int, pos, vel, ids) -> None:
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        vel = np.ascontiguousarray(vel, np.float32).reshape(-1, 3) if vel is not None else None
        ids = np.ascontiguousarray(ids, np.uint32)
        P = lambda a, t: a.ctypes.data_as(C.POINTER(t)) if a is not None else None
        N.check(self._lib.sph_cluster_load(self._h, int(local_index), len(ids), P(pos, C.c_float),
                                           P(vel, C.c_float), P(ids, C.c_uint32)))

    # -- stepping -------------------------------------------------------------------------
    def advance(self, steps: int) -> None:
        N.check(self._lib.sph_cluster_advance(self._h, int(steps)))

    def advance_timed(self, steps: int) -> float:
        ms = C.c_float()
        N.check(self._lib.sph_cluster_advance_timed(self._h, int(steps), C.byref(ms)))
        return float(ms.value)

    def step(self) -> None:
        """One step; every owned particle's {x, y, z, id} record travels to pinned host memory."""
        N.check(self._lib.sph_cluster_step(self._h))

    def sync(self) -> None:
        N.check(self._lib.sph_cluster_sync(self._h))

    def rebalance(self) -> None:
        N.check(self._lib.sph_cluster_rebalance(self._h))

    # -- results --------------------------------------------------------------------------
    def host_records(self, local_index: int) -> np.ndarray:
        """(count, 4) float32 view of the slab's pinned host buffer: x, y, z, id (bit pattern)."""
        p = C.POINTER(C.c_float)()
        n = C.c_int()
        N.check(self._lib.sph_cluster_host_records(self._h, int(local_index), C.byref(p), C.byref(n)))
        if not p or n.value <= 0:
            return np.zeros((0, 4), np.float32)
        return np.ctypeslib.as_array(p, shape=(n.value, 4))

    def positions(self, n_global: int) -> np.ndarray:
        """getPosition() for the ids this process holds: (n_global, 3), NaN where an id lives in
        another process."""
        out = np.full((int(n_global), 3), np.nan, np.float32)
        N.check(self._lib.sph_cluster_positions(self._h, out.ctypes.data_as(C.POINTER(C.c_float)),
                                                C.c_int64(int(n_global))))
        return out

    def download(self, local_index: int, capacity: int):
        ids = np.empty(capacity, np.uint32)
        pos = np.empty((capacity, 3), np.float32)
        vel = np.empty((capacity, 3), np.float32)
        n = C.c_int()
        P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        N.check(self._lib.sph_cluster_download(self._h, int(local_index), P(ids, C.c_uint32),
                                               P(pos, C.c_float), P(vel, C.c_float), C.byref(n)))
        return ids[:n.value].copy(), pos[:n.value].copy(), vel[:n.value].copy()

    def download_all(self, capacity: int):
        """Every local slab's particles, sorted by id."""
        parts = [self.download(i, capacity) for i in range(self.local_count)]
        ids = np.concatenate([p[0] for p in parts])
        order = np.argsort(ids)
        return (ids[order], np.concatenate([p[1] for p in parts])[order],
                np.concatenate([p[2] for p in parts])[order])

    def stats(self, local_index: int) -> dict:
        s = N.SphSlabStats()
        N.check(self._lib.sph_cluster_stats(self._h, int(local_index), C.byref(s)))
        return {name: getattr(s, name) for name, _ in N.SphSlabStats._fields_}

    @property
    def launch_count(self) -> int:
        return int(self._lib.sph_cluster_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.sph_cluster_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
