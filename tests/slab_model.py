"""CPU model of the slab decomposition protocol -- test infrastructure only.

The product's multi-GPU step lives in csrc/sph_cluster.cu (C++/CUDA, NCCL or peer-to-peer
copies) and needs GPUs.  This Python model states the SAME decomposition -- z-slabs, ghost halo
exchange A (pos/vel) and B (pressure terms), emigrant migration, dead entries dropped by the next
sort -- over torch.distributed so that tests/test_slab_gloo.py can run it on CPU with world sizes 2
and 3 (gloo) against the undecomposed oracle, with tests/fake_slab.py standing in for the library.
It is not imported by the package.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass

import torch
import torch.distributed as dist


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
