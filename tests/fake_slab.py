"""CPU stand-in for a slab of the multi-GPU decomposition (the buffers and the call protocol
tests/slab_model.py drives), built on the CPU oracle.  Lets the halo /
migration protocol of SlabDriver run under gloo without a GPU.  Test infrastructure."""
import numpy as np
import torch

from slab_model import SlabInfo
from oracle.oracle import CpuOracle

DEAD = np.uint32(0xFFFFFFFF)


class FakeSlab:
    def __init__(self, zlo, zhi, nz, capacity, ghost_capacity, emig_capacity, h=0.1, box=10.0, nc=100):
        self.zlo, self.zhi, self.nz, self.h, self.nc = zlo, zhi, nz, np.float32(h), nc
        self.capacity, self.ghost_capacity = capacity, ghost_capacity
        scap = capacity + 2 * ghost_capacity
        z = lambda r, c: torch.zeros(r, c, dtype=torch.float32)
        self.srt_pos, self.srt_vel, self.pa = z(scap, 4), z(scap, 4), z(scap, 2)
        self.cur_pos, self.cur_vel = z(capacity, 4), z(capacity, 4)
        self.emig_pos = [z(emig_capacity, 4), z(emig_capacity, 4)]
        self.emig_vel = [z(emig_capacity, 4), z(emig_capacity, 4)]
        self.n_total, self.dead = 0, np.zeros(capacity, bool)
        self.o = CpuOracle(1, h=h, boxDim=box, numCellsPerDim=nc)
        self.slot0 = ghost_capacity

    def _cz(self, pos):
        return (pos[:, 2] / self.h).astype(np.int64)

    def load(self, pos, vel, ids):
        n = len(ids)
        self.cur_pos[:n, :3] = torch.from_numpy(np.asarray(pos, np.float32))
        self.cur_pos[:n, 3] = torch.from_numpy(np.asarray(ids, np.uint32).view(np.float32))
        self.cur_vel[:n, :3] = torch.from_numpy(np.asarray(vel, np.float32))
        self.n_total, self.dead[:] = n, False

    def build(self):
        live = np.nonzero(~self.dead[:self.n_total])[0]
        pos = self.cur_pos[live].numpy()
        cells, ff, fi, mo = self.o.keys(pos[:, :3].copy())
        order = np.argsort(fi, kind="stable")
        n, s0 = len(live), self.slot0
        self.srt_pos[s0:s0 + n] = self.cur_pos[live][order]
        self.srt_vel[s0:s0 + n] = self.cur_vel[live][order]
        cz = cells[order, 2]
        lo = np.nonzero(cz == self.zlo)[0]
        hi = np.nonzero(cz == self.zhi - 1)[0]
        self.n_owned, self.n_total = n, n
        self.dead[:] = False
        rng = lambda a: (int(a[0]) + s0, len(a)) if len(a) else (s0, 0)
        (lf, lc), (hf, hc) = rng(lo), rng(hi)
        return SlabInfo(n, n, s0, lf, lc, hf, hc)

    def _combined(self, g_lo, g_hi):
        s0, n = self.slot0, self.n_owned
        return slice(s0 - g_lo, s0 + n + g_hi), g_lo

    def density(self, g_lo, g_hi):
        sl, off = self._combined(g_lo, g_hi)
        self.g = (g_lo, g_hi)
        pos = self.srt_pos[sl, :3].numpy().copy()
        rho, prs, _, _ = self.o.density(pos, counts=False)
        own = slice(off, off + self.n_owned)
        pa = np.stack([prs[own], np.float32(-0.01) / rho[own]], 1).astype(np.float32)
        self.pa[self.slot0:self.slot0 + self.n_owned] = torch.from_numpy(pa)
        self.rho = rho[own]

    def force(self):
        g_lo, g_hi = self.g
        sl, off = self._combined(g_lo, g_hi)
        n, s0 = self.n_owned, self.slot0
        pos = self.srt_pos[sl, :3].numpy().copy()
        vel = self.srt_vel[sl, :3].numpy().copy()
        pa = self.pa[sl].numpy()
        prs, rho = pa[:, 0].copy(), (np.float32(-0.01) / pa[:, 1]).astype(np.float32)
        rho[off:off + n] = self.rho            # exact own densities (ghost ones via 1/a)
        f = self.o.forces(pos, vel, rho, prs)[off:off + n]
        p1, v1 = pos[off:off + n].copy(), vel[off:off + n].copy()
        oi = CpuOracle(n)
        oi.L.oracle_integrate(p1.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)),
                              v1.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)),
                              np.ascontiguousarray(f).ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)),
                              self.rho.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)),
                              n, __import__("ctypes").byref(self.o.s))
        self.cur_pos[:n, :3] = torch.from_numpy(p1)
        self.cur_pos[:n, 3] = self.srt_pos[s0:s0 + n, 3]
        self.cur_vel[:n, :3] = torch.from_numpy(v1)
        cz = self._cz(p1)
        down, up = np.nonzero(cz < self.zlo)[0], np.nonzero(cz >= self.zhi)[0]
        for side, idx in ((0, down), (1, up)):
            self.emig_pos[side][:len(idx)] = self.cur_pos[idx]
            self.emig_vel[side][:len(idx)] = self.cur_vel[idx]
        self.dead[:] = False
        self.dead[down] = True
        self.dead[up] = True
        return SlabInfo(n - len(down) - len(up), n, s0, emig_down=len(down), emig_up=len(up))

    def append(self, count):
        self.n_total += count

    def download(self):
        live = np.nonzero(~self.dead[:self.n_total])[0]
        p = self.cur_pos[live].numpy()
        return p[:, 3].copy().view(np.uint32), p[:, :3].copy(), self.cur_vel[live, :3].numpy().copy()


class FastFakeSlab(FakeSlab):
    """The split protocol of the library on the CPU: device-side counts (`counts`), build / force
    in _async + _finish halves, and density / force in interior + boundary parts over particle
    CTAs of 128 (sph_slab_interior_ctas, sph_slab_density_part, sph_slab_force_part).  A part only
    commits the particles of its own CTAs and works with the ghost data present at that moment,
    so a driver that runs a part too early, or with a wrong interior range, gets wrong physics."""
    CTA = 128

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.counts = torch.zeros(8, dtype=torch.int32)
        self.emig_capacity = len(self.emig_pos[0])

    # -- build -------------------------------------------------------------------------
    def build_async(self):
        self._info = self.build()
        i = self._info
        self.counts[:4] = torch.tensor([i.lo_first, i.lo_first + i.lo_count, i.hi_first, i.hi_first + i.hi_count],
                                       dtype=torch.int32)
        # nothing of this step has arrived yet: stale ghost slots must not be trusted
        self.srt_pos[:self.slot0] = float("nan")
        self.srt_pos[self.slot0 + self.n_owned:] = float("nan")
        self.pa[:self.slot0] = float("nan")
        self.pa[self.slot0 + self.n_owned:] = float("nan")
        self._rho = np.full(self.n_owned, np.nan, np.float32)
        self._f = np.full((self.n_owned, 3), np.nan, np.float32)

    def build_finish(self):
        return self._info

    def interior_ctas(self):
        i, n, C = self._info, self.n_owned, self.CTA
        lo_end = i.lo_first + i.lo_count - self.slot0 if i.lo_count else 0
        hi_first = i.hi_first - self.slot0 if i.hi_count else n     # (an empty layer starts at the end)
        total = (n + C - 1) // C
        a = min((max(lo_end, 0) + C - 1) // C, total)
        b = max(min(hi_first, n), 0) // C
        return a, max(a, b)

    def _members(self, part, ctas):
        n, C = self.n_owned, self.CTA
        total = (n + C - 1) // C
        a, b = min(ctas[0], total), min(max(ctas[1], ctas[0]), total)
        inside = np.zeros(n, bool)
        inside[a * C:min(b * C, n)] = True
        return inside if part == 0 else ~inside

    # -- density -----------------------------------------------------------------------
    def density_part(self, part, ctas, g_lo=0, g_hi=0):
        if part == 1:
            self.g = (g_lo, g_hi)
        g = (0, 0) if part == 0 else (g_lo, g_hi)          # the interior part may not look at ghosts
        sl, off = self._combined(*g)
        pos = self.srt_pos[sl, :3].numpy().copy()
        rho, prs, _, _ = self.o.density(pos, counts=False)
        own = slice(off, off + self.n_owned)
        m = self._members(part, ctas)
        pa = np.stack([prs[own], np.float32(-0.01) / rho[own]], 1).astype(np.float32)
        idx = torch.from_numpy(np.nonzero(m)[0] + self.slot0)
        self.pa[idx] = torch.from_numpy(pa[m])
        self._rho[m] = rho[own][m]

    # -- force -------------------------------------------------------------------------
    def force_part(self, part, ctas):
        g = (0, 0) if part == 0 else self.g
        sl, off = self._combined(*g)
        n = self.n_owned
        pos = self.srt_pos[sl, :3].numpy().copy()
        vel = self.srt_vel[sl, :3].numpy().copy()
        pa = self.pa[sl].numpy()
        prs, rho = pa[:, 0].copy(), (np.float32(-0.01) / pa[:, 1]).astype(np.float32)
        rho[off:off + n] = self._rho
        m = self._members(part, ctas)
        self._f[m] = self.o.forces(pos, vel, rho, prs)[off:off + n][m]
        if part == 1:
            self._pending = self._integrate()

    def _integrate(self):
        import ctypes as C
        n, s0 = self.n_owned, self.slot0
        assert not np.isnan(self._f).any() and not np.isnan(self._rho).any(), "a part was skipped or ran too early"
        p1 = self.srt_pos[s0:s0 + n, :3].numpy().copy()
        v1 = self.srt_vel[s0:s0 + n, :3].numpy().copy()
        P = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(C.c_float))
        f, rho = np.ascontiguousarray(self._f), np.ascontiguousarray(self._rho)
        CpuOracle(n).L.oracle_integrate(P(p1), P(v1), P(f), P(rho), n, C.byref(self.o.s))
        self.cur_pos[:n, :3] = torch.from_numpy(p1)
        self.cur_pos[:n, 3] = self.srt_pos[s0:s0 + n, 3]
        self.cur_vel[:n, :3] = torch.from_numpy(v1)
        cz = self._cz(p1)
        down, up = np.nonzero(cz < self.zlo)[0], np.nonzero(cz >= self.zhi)[0]
        for side, idx in ((0, down), (1, up)):
            self.emig_pos[side][:len(idx)] = self.cur_pos[idx]
            self.emig_vel[side][:len(idx)] = self.cur_vel[idx]
        self.dead[:] = False
        self.dead[down] = True
        self.dead[up] = True
        self.counts[4:6] = torch.tensor([len(down), len(up)], dtype=torch.int32)
        return SlabInfo(n - len(down) - len(up), n, s0, emig_down=len(down), emig_up=len(up))

    def force_async(self):   # whole-slab form (driver without overlap)
        self.density_rho_from_whole()
        self._pending = None
        info = self.force()
        self.counts[4:6] = torch.tensor([info.emig_down, info.emig_up], dtype=torch.int32)
        self._pending = info

    def density_rho_from_whole(self):
        pass   # FakeSlab.density() already left self.rho for FakeSlab.force()

    def force_finish(self):
        return self._pending
