"""Slab decomposition of the SPH step across GPUs (north_star: spatial slabs along a
horizontal axis, per-step ghost-particle halo exchange and particle migration).

One process per GPU (torch.distributed).  Rank r owns the global cell layers
[zlo_r, zhi_r) along z -- gravity is -y and the reference's grid init fills x-planes
first, so z-slabs start balanced.  Per timestep:

    build      sort owned particles by (local) flat key              [library]
    exchange A boundary layers' pos/vel -> neighbours' ghost slots    [send/recv]
    density    ghost cell ranges + density/pressure of owned          [library]
    exchange B boundary layers' {p, a}  -> neighbours' ghost slots    [send/recv]
    force      force + integrate owned, emigrants packed              [library]
    migrate    emigrants -> neighbours' particle arrays               [send/recv]

All payloads are contiguous slot ranges of the library's own device buffers (boundary
layers are contiguous in z-major flat-key order), so transfers are zero-copy views; only
a few counts cross the host.  The driver is backend-agnostic: `SlabBackend` is the B200
library, tests drive the same protocol over gloo with a CPU stand-in.

There is no reference equivalent (the reference is single-GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import time
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


def slab_ranges(nz: int, world: int):
    """Contiguous, near-equal layer ranges [(zlo, zhi)] covering [0, nz)."""
    base, extra = divmod(nz, world)
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((z, z + n))
        z += n
    return out


@dataclass
class SlabInfo:
    n_owned: int = 0
    n_total: int = 0
    slot0: int = 0
    lo_first: int = 0
    lo_count: int = 0
    hi_first: int = 0
    hi_count: int = 0
    emig_down: int = 0
    emig_up: int = 0
    overflow: int = 0


class _DevView:
    """Exposes a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 3}


def _tensor(ptr, rows, cols, device):
    t = torch.as_tensor(_DevView(ptr, rows * cols * 4), device=device)
    return t.view(torch.float32).view(rows, cols)


class SlabBackend:
    """The B200 library in slab mode (libsph_b200.so, sph_slab_* entry points)."""

    def __init__(self, settings, zlo, zhi, nz, capacity, device=0, ghost_capacity=0,
                 emig_capacity=0, density_sum=0):
        from . import _native as N
        self.N = N
        self.lib = N.load()
        opt = N.SphOptions()
        opt.device = device
        opt.key_mode = N.SPH_KEY_FLAT
        opt.use_graph = 2
        opt.capacity = int(capacity)
        opt.z_cell_lo, opt.z_cell_hi, opt.nz_cells = int(zlo), int(zhi), int(nz)
        opt.ghost_capacity, opt.emig_capacity = int(ghost_capacity), int(emig_capacity)
        opt.density_sum = int(density_sum)
        cs = settings.to_c()
        cs.numParticles = 0
        h = C.c_void_p()
        N.check(self.lib.sph_create_ex(C.byref(cs), C.byref(opt), C.byref(h)))
        self.h = h
        N.check(self.lib.sph_setup(h))
        torch.cuda.set_device(device)
        N.check(self.lib.sph_set_stream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        b = N.SphSlabBuffers()
        N.check(self.lib.sph_slab_buffers(h, C.byref(b)))
        dev = torch.device("cuda", device)
        scap = b.capacity + 2 * b.ghost_capacity
        self.srt_pos = _tensor(b.srt_pos, scap, 4, dev)
        self.srt_vel = _tensor(b.srt_vel, scap, 4, dev)
        self.pa = _tensor(b.pa, scap, 2, dev)
        self.cur_pos = _tensor(b.cur_pos, b.capacity, 4, dev)
        self.cur_vel = _tensor(b.cur_vel, b.capacity, 4, dev)
        self.emig_pos = [_tensor(b.emig_pos[i], b.emig_capacity, 4, dev) for i in range(2)]
        self.emig_vel = [_tensor(b.emig_vel[i], b.emig_capacity, 4, dev) for i in range(2)]
        self.counts = torch.as_tensor(_DevView(b.counts, 8 * 4), device=dev).view(torch.int32)
        self.capacity, self.ghost_capacity, self.emig_capacity = b.capacity, b.ghost_capacity, b.emig_capacity
        self.device = dev

    def _info(self, i):
        return SlabInfo(i.n_owned, i.n_total, i.slot0, i.lo_first, i.lo_count, i.hi_first,
                        i.hi_count, i.emig_count[0], i.emig_count[1], i.overflow)

    def load(self, pos, vel, ids):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32) if vel is not None else None
        ids = np.ascontiguousarray(ids, np.uint32)
        P = lambda a, t: a.ctypes.data_as(C.POINTER(t)) if a is not None else None
        self.N.check(self.lib.sph_slab_load(self.h, len(ids), P(pos, C.c_float), P(vel, C.c_float),
                                            P(ids, C.c_uint32)))

    def build(self):
        i = self.N.SphSlabInfo()
        self.N.check(self.lib.sph_slab_build(self.h, C.byref(i)))
        return self._info(i)

    def build_async(self):
        self.N.check(self.lib.sph_slab_build_async(self.h))

    def build_finish(self):
        i = self.N.SphSlabInfo()
        self.N.check(self.lib.sph_slab_build_finish(self.h, C.byref(i)))
        return self._info(i)

    def force_async(self):
        self.N.check(self.lib.sph_slab_force_async(self.h))

    def force_finish(self):
        i = self.N.SphSlabInfo()
        self.N.check(self.lib.sph_slab_force_finish(self.h, C.byref(i)))
        return self._info(i)

    def density(self, g_lo, g_hi):
        self.N.check(self.lib.sph_slab_density(self.h, int(g_lo), int(g_hi)))

    def interior_ctas(self):
        a, b = C.c_int(), C.c_int()
        self.N.check(self.lib.sph_slab_interior_ctas(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def density_part(self, part, ctas, g_lo=0, g_hi=0):
        self.N.check(self.lib.sph_slab_density_part(self.h, int(part), int(ctas[0]), int(ctas[1]),
                                                    int(g_lo), int(g_hi)))

    def force_part(self, part, ctas):
        self.N.check(self.lib.sph_slab_force_part(self.h, int(part), int(ctas[0]), int(ctas[1])))

    def force(self):
        i = self.N.SphSlabInfo()
        self.N.check(self.lib.sph_slab_force(self.h, C.byref(i)))
        return self._info(i)

    def append(self, count):
        self.N.check(self.lib.sph_slab_append(self.h, int(count)))

    def download(self):
        cap = self.capacity
        ids = np.empty(cap, np.uint32)
        pos = np.empty((cap, 3), np.float32)
        vel = np.empty((cap, 3), np.float32)
        n = C.c_int()
        P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        self.N.check(self.lib.sph_slab_download(self.h, P(ids, C.c_uint32), P(pos, C.c_float),
                                                P(vel, C.c_float), C.byref(n)))
        return ids[:n.value].copy(), pos[:n.value].copy(), vel[:n.value].copy()

    def profile_enable(self, on=True):
        self.N.check(self.lib.sph_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, reset=False):
        N = self.N
        ms = (C.c_double * N.SPH_STAGE_COUNT)()
        ln = (C.c_int64 * N.SPH_STAGE_COUNT)()
        N.check(self.lib.sph_profile_read(self.h, ms, ln, 1 if reset else 0))
        return {self.lib.sph_stage_name(i).decode(): {"ms": ms[i], "launches": ln[i]}
                for i in range(N.SPH_STAGE_COUNT)}

    @property
    def launch_count(self):
        return int(self.lib.sph_launch_count(self.h))

    def close(self):
        if self.h:
            self.lib.sph_destroy(self.h)
            self.h = None


class SlabDriver:
    """The per-step exchange protocol between neighbouring slabs (backend-agnostic)."""

    def __init__(self, backend, rank: int, world: int, group=None, overlap: bool = True):
        self.b, self.rank, self.world, self.group = backend, rank, world, group
        # overlap: the interior CTAs (sph_slab_*_part) run under the halo exchanges, and the interior
        # density is launched on a guess from the previous step before the counts of this step have
        # crossed the host (checked afterwards, redone if wrong).  Same results.
        self.overlap = overlap
        # host-side time per protocol phase (ms, accumulated) when SPH_SLAB_TRACE=1: shows where the
        # host, not the GPU, paces the step
        self.host_ms = {} if os.environ.get("SPH_SLAB_TRACE") else None
        self._t0 = 0.0
        self._guess = None        # (cta_a, cta_b) expected to be interior in the next step
        self.guess_margin = 8     # particle CTAs (1024 particles) the boundary layers may grow per step
        self.down = rank - 1 if rank > 0 else None          # owner of lower z
        self.up = rank + 1 if rank < world - 1 else None    # owner of higher z
        self.stats = {"ghost_particles": 0, "migrated_particles": 0, "steps": 0}

    def _mark(self, name=None):
        if self.host_ms is None:
            return
        now = time.perf_counter()
        if name is not None:
            self.host_ms[name] = self.host_ms.get(name, 0.0) + (now - self._t0) * 1e3
        self._t0 = now

    # -- plumbing -------------------------------------------------------------------
    def _counts(self, mine):
        """All ranks' small integer tuples (one all_gather, one host sync)."""
        dev = self.b.srt_pos.device
        t = torch.tensor(mine, dtype=torch.int64, device=dev)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return [o.tolist() for o in out]

    def _exchange_start(self, sends, recvs):
        """sends / recvs: lists of (tensor, peer); empty tensors are skipped on both sides.
        Returns the requests; work enqueued before _exchange_wait() overlaps the transfer."""
        ops = [dist.P2POp(dist.irecv, t, peer, group=self.group) for t, peer in recvs if t.numel()]
        ops += [dist.P2POp(dist.isend, t, peer, group=self.group) for t, peer in sends if t.numel()]
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def _exchange_wait(reqs):
        for req in reqs:
            req.wait()

    def _exchange(self, sends, recvs):
        self._exchange_wait(self._exchange_start(sends, recvs))

    def _neighbour_counts(self, lo, hi):
        """My device counts[lo:hi] -> both neighbours, theirs -> a pinned host mirror
        (row 0 = from below, row 1 = from above); asynchronous, complete after the next
        synchronisation of the stream."""
        b = self.b
        if not hasattr(self, "_nb_dev"):
            self._nb_dev = torch.zeros((2, 8), dtype=torch.int32, device=b.counts.device)
            self._nb_host = torch.zeros((2, 8), dtype=torch.int32)
            if b.counts.is_cuda:
                self._nb_host = self._nb_host.pin_memory()
        mine = b.counts[lo:hi]
        sends, recvs = [], []
        if self.down is not None:
            sends.append((mine, self.down)); recvs.append((self._nb_dev[0, lo:hi], self.down))
        if self.up is not None:
            sends.append((mine, self.up)); recvs.append((self._nb_dev[1, lo:hi], self.up))
        self._exchange(sends, recvs)
        self._nb_host.copy_(self._nb_dev, non_blocking=True)

    # -- one timestep ------------------------------------------------------------------
    def step(self):
        b = self.b
        fast = hasattr(b, "build_async")   # device-side counts: neighbours only, one sync per phase
        self._mark()
        split = fast and self.overlap and hasattr(b, "density_part")
        spec = None   # interior CTAs guessed from the previous step, launched before the counts are known
        if fast:
            b.build_async()
            self._mark("build issue")
            self._neighbour_counts(0, 4)
            counts_ready = torch.cuda.Event() if b.counts.is_cuda else None   # (CPU stand-in: synchronous)
            if counts_ready is not None:
                counts_ready.record()
            if split and self._guess is not None and self._guess[1] > self._guess[0]:
                spec = self._guess
                b.density_part(0, spec)   # keeps the GPU busy across the host round trip below
            self._mark("count exchange 1 issue")
            if counts_ready is not None:
                counts_ready.synchronize()
            info = b.build_finish()
            self._mark("sync 1 (build + counts)")
            nb = self._nb_host
            g_lo = int(nb[0, 3] - nb[0, 2]) if self.down is not None else 0   # their highest layer
            g_hi = int(nb[1, 1] - nb[1, 0]) if self.up is not None else 0     # their lowest layer
        else:
            info = b.build()
            counts = self._counts([info.lo_count, info.hi_count])
            g_lo = counts[self.down][1] if self.down is not None else 0
            g_hi = counts[self.up][0] if self.up is not None else 0
        n, s0 = info.n_owned, info.slot0
        if g_lo > b.ghost_capacity or g_hi > b.ghost_capacity:
            raise RuntimeError(f"rank {self.rank}: ghost layer ({g_lo}, {g_hi}) exceeds capacity {b.ghost_capacity}")
        lo = slice(info.lo_first, info.lo_first + info.lo_count)      # my lowest owned layer
        hi = slice(info.hi_first, info.hi_first + info.hi_count)      # my highest owned layer
        glo = slice(s0 - g_lo, s0)                                    # ghosts from below
        ghi = slice(s0 + n, s0 + n + g_hi)                            # ghosts from above

        def halo(arrs):
            sends, recvs = [], []
            for a in arrs:
                if self.down is not None:
                    sends.append((a[lo], self.down)); recvs.append((a[glo], self.down))
                if self.up is not None:
                    sends.append((a[hi], self.up)); recvs.append((a[ghi], self.up))
            return self._exchange_start(sends, recvs)

        # interior CTAs (no particle of a boundary layer) need no ghosts: they run under the exchanges
        ctas = None
        if split:
            ctas = b.interior_ctas()
            total = (n + 127) // 128
            if spec is not None and spec[0] >= ctas[0] and spec[1] <= ctas[1] and spec[1] <= total:
                ctas = spec                       # the guess holds: its density is already running
            else:
                spec = None                       # no / wrong guess: launch the true interior now
            self.stats["speculative_hits"] = self.stats.get("speculative_hits", 0) + (spec is not None)
            # next step's guess: this step's interior shrunk by a margin on both sides
            m = self.guess_margin
            self._guess = (ctas[0] + m, ctas[1] - m) if ctas[1] - ctas[0] > 2 * m else None
        self._mark("slices")
        reqs = halo([b.srt_pos, b.srt_vel])   # exchange A
        if split and spec is None:
            b.density_part(0, ctas)
        self._exchange_wait(reqs)
        self._mark("exchange A issue")
        if split:
            b.density_part(1, ctas, g_lo, g_hi)
        else:
            b.density(g_lo, g_hi)
        self._mark("density issue")
        reqs = halo([b.pa])                    # exchange B
        if split:
            b.force_part(0, ctas)
        self._exchange_wait(reqs)
        self._mark("exchange B issue")

        # migration: my emigrants -> neighbours; theirs are appended behind my particles
        if fast:
            if split:
                b.force_part(1, ctas)
            else:
                b.force_async()
            self._mark("force issue")
            self._neighbour_counts(4, 6)
            self._mark("count exchange 2 issue")
            f = b.force_finish()
            self._mark("sync 2 (density + force + counts)")
            nb = self._nb_host
            # (senders cap at their emigrant buffer; all slabs are created with the same capacity)
            in_dn = min(int(nb[0, 5]), b.emig_capacity) if self.down is not None else 0   # from below, moving up
            in_up = min(int(nb[1, 4]), b.emig_capacity) if self.up is not None else 0     # from above, moving down
        else:
            f = b.force()
            em = self._counts([f.emig_down, f.emig_up])
            in_dn = em[self.down][1] if self.down is not None else 0
            in_up = em[self.up][0] if self.up is not None else 0
        at = f.n_total
        if at + in_dn + in_up > b.capacity:
            raise RuntimeError(f"rank {self.rank}: {at}+{in_dn}+{in_up} particles exceed capacity {b.capacity}")
        sends, recvs = [], []
        for src, dst in ((b.emig_pos, b.cur_pos), (b.emig_vel, b.cur_vel)):
            if self.down is not None:
                sends.append((src[0][:f.emig_down], self.down))
                recvs.append((dst[at:at + in_dn], self.down))
            if self.up is not None:
                sends.append((src[1][:f.emig_up], self.up))
                recvs.append((dst[at + in_dn:at + in_dn + in_up], self.up))
        self._exchange(sends, recvs)
        b.append(in_dn + in_up)
        self._mark("migration issue")
        if f.overflow:
            raise RuntimeError(f"rank {self.rank}: slab capacity overflow flags {f.overflow}")
        self.stats["ghost_particles"] += g_lo + g_hi
        self.stats["migrated_particles"] += in_dn + in_up
        self.stats["steps"] += 1
        self.last = {"n_owned": n, "ghosts": g_lo + g_hi, "immigrants": in_dn + in_up}
        return self.last


def partition(pos, h, ranges):
    """Indices of the particles each slab owns: global z cell (IEEE divide, truncate --
    ref: simulator.cu:69) in [zlo, zhi)."""
    cz = (np.asarray(pos, np.float32)[:, 2] / np.float32(h)).astype(np.int64)
    return [np.nonzero((cz >= lo) & (cz < hi))[0] for lo, hi in ranges]


class LocalSlabCluster:
    """All slabs driven by ONE process: the same protocol as SlabDriver with direct tensor
    copies between the slabs' buffers (device-to-device, peer-to-peer when the slabs live
    on different GPUs).  Used to test slab mode on a single GPU and as the one-process /
    many-devices mode."""

    def __init__(self, backends, split: bool = False):
        self.b = list(backends)
        self.split = split   # interior / boundary parts, in the order SlabDriver issues them
        self.stats = {"ghost_particles": 0, "migrated_particles": 0, "steps": 0}

    def step(self):
        B = self.b
        W = len(B)
        info = [b.build() for b in B]
        g_lo = [info[r - 1].hi_count if r > 0 else 0 for r in range(W)]
        g_hi = [info[r + 1].lo_count if r < W - 1 else 0 for r in range(W)]

        def halo(name):
            for r in range(W):
                dst, i = getattr(B[r], name), info[r]
                if r > 0:       # ghosts from below = highest owned layer of r-1
                    src, j = getattr(B[r - 1], name), info[r - 1]
                    dst[i.slot0 - g_lo[r]:i.slot0].copy_(src[j.hi_first:j.hi_first + j.hi_count])
                if r < W - 1:   # ghosts from above = lowest owned layer of r+1
                    src, j = getattr(B[r + 1], name), info[r + 1]
                    dst[i.slot0 + i.n_owned:i.slot0 + i.n_owned + g_hi[r]].copy_(
                        src[j.lo_first:j.lo_first + j.lo_count])

        ctas = [b.interior_ctas() for b in B] if self.split else None
        if self.split:
            for r in range(W):
                B[r].density_part(0, ctas[r])
        halo("srt_pos")
        halo("srt_vel")
        for r in range(W):
            if self.split:
                B[r].density_part(1, ctas[r], g_lo[r], g_hi[r])
                B[r].force_part(0, ctas[r])
            else:
                B[r].density(g_lo[r], g_hi[r])
        halo("pa")
        if self.split:
            for r in range(W):
                B[r].force_part(1, ctas[r])
            f = [b.force_finish() for b in B]
        else:
            f = [b.force() for b in B]
        for r in range(W):
            at, n_in = f[r].n_total, 0
            for src_rank, side, cnt in ((r - 1, 1, f[r - 1].emig_up if r > 0 else 0),
                                        (r + 1, 0, f[r + 1].emig_down if r < W - 1 else 0)):
                if cnt:
                    B[r].cur_pos[at + n_in:at + n_in + cnt].copy_(B[src_rank].emig_pos[side][:cnt])
                    B[r].cur_vel[at + n_in:at + n_in + cnt].copy_(B[src_rank].emig_vel[side][:cnt])
                    n_in += cnt
            B[r].append(n_in)
            self.stats["migrated_particles"] += n_in
            self.stats["ghost_particles"] += g_lo[r] + g_hi[r]
            if f[r].overflow:
                raise RuntimeError(f"slab {r}: capacity overflow flags {f[r].overflow}")
        self.stats["steps"] += 1

    def download(self):
        parts = [b.download() for b in self.b]
        ids = np.concatenate([p[0] for p in parts])
        order = np.argsort(ids)
        return (ids[order], np.concatenate([p[1] for p in parts])[order],
                np.concatenate([p[2] for p in parts])[order])
