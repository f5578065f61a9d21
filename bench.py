#!/usr/bin/env python
"""bench.py -- particle-updates/s of the SPH timestep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 16m_grid|1m_random|10k_grid]
                    [--key flat|morton] [--impl b200|reference]

One "step" = one full timestep (hash, radix sort, cell ranges, reorder, density/pressure,
force + integrate) of every particle.  Prints ONE JSON line (rank 0).

  value        N*K / device time of K graph-replayed steps, state resident in HBM
               (CUDA events on the simulator's own stream, max over ranks)
  e2e          the same K steps through the reference-facing call sph_step()
               (== Simulator::simulate()): every step ends with the blocking
               device->host copy of all N positions into the simulator's pinned host
               buffer, exactly the reference's per-step contract (ref: simulator.cu:478-480).
               The path has no per-step host input (state is device-resident by the
               reference's own design, uploaded once in setup()), so h2d is 0 per step and
               the one-off setup upload is reported as setup_h2d_bytes.
  roofline     dominant kernel, measured live with CUDA events around every launch in a
               separate profiled pass (plain launches) right after the timed region
  cpu_baseline serial C port of the step (oracle/sph_oracle.c), 1 thread, bounded sample
  --impl reference   the reference's own implementation: its unmodified CUDA code
               (oracle/_ref/libsph_ref.so) on the same GPU -- the reference ships no CPU
               path; falls back to the serial C port if that library
               is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[2]: 16M particles, grid init, scaled box
    "16m_grid": dict(n=16_000_000, randomInit=False, boxDim=25.6, numCellsPerDim=256.0),
    # BASELINE.json configs[4] per-GPU share: 32M particles (needs a 320-cell box: 283^3 < 32M)
    "32m_grid": dict(n=32_000_000, randomInit=False, boxDim=32.0, numCellsPerDim=320.0),
    # BASELINE.json configs[1]: 1M particles, random init, reference box
    "1m_random": dict(n=1_000_000, randomInit=True, boxDim=10.0, numCellsPerDim=100.0),
    # BASELINE.json configs[0]: ./sph -n 10000 -i grid -m time
    "10k_grid": dict(n=10_000, randomInit=False, boxDim=10.0, numCellsPerDim=100.0),
}


# ---- clocks during the timed region --------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
        0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):   # may be entered several times: samples of all timed regions are pooled
        if self.nv:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [],
                    "note": "nvml unavailable" if not self.nv else "no samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


def workload_config(name, wl):
    """`config` of the JSON line: the same keys and values in both arms (b200 and reference)."""
    return {"workload": name, "n": wl["n"], "boxDim": wl["boxDim"], "numCellsPerDim": wl["numCellsPerDim"],
            "h": 0.1, "timestep": 0.01,
            "init": "random(glibc rand seed 1)" if wl["randomInit"] else "grid lattice (dam-break column)"}


def dist_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


# ---- CPU baseline: the serial C port on a bounded sample -----------------------------
def cpu_baseline(wl, budget_s=20.0):
    from oracle.oracle import CpuOracle
    o = CpuOracle(wl["n"], boxDim=wl["boxDim"], numCellsPerDim=wl["numCellsPerDim"],
                  randomInit=wl["randomInit"])
    o.setup()
    steps, t0 = 0, time.perf_counter()
    while True:
        o.step()
        steps += 1
        el = time.perf_counter() - t0
        if el > budget_s or steps >= 100 or el + el / steps > 1.5 * budget_s:
            break
    return {"value": wl["n"] * steps / el, "unit": "particle-updates/s", "cores": 1, "kind": "port",
            "sample": f"first {steps} step(s) of the same workload ({wl['n']} particles) from its initial "
                      f"state, {el:.1f} s on one host thread; early steps are the sparsest of the run, "
                      f"which favours the CPU",
            "host_cores_available": os.cpu_count()}


# ---- the reference arm ------------------------------------------------------------------
def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    from oracle.oracle import REF_SO
    line = {"impl": "reference", "metric": "particle-updates/s", "unit": "particle-updates/s",
            "n_gpus": args.gpus, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, wl)}
    budget = 150.0
    if REF_SO.exists():
        import torch
        from oracle.oracle import RefSim
        torch.cuda.init()
        ref = RefSim(wl["n"], boxDim=wl["boxDim"], numCellsPerDim=wl["numCellsPerDim"],
                     randomInit=wl["randomInit"])
        for _ in range(args.warmup):
            ref.step_timed()
        ref.buckets[0] = ref.buckets[1] = ref.buckets[2] = 0.0
        torch.cuda.synchronize()
        steps, t0 = 0, time.perf_counter()
        while steps < args.steps and time.perf_counter() - t0 < budget:
            ref.step_timed()     # Simulator::simulateAndTime(): its own three wall-clock buckets
            steps += 1
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        b = list(ref.buckets)
        dev = b[0] + b[1]
        line.update({
            "value": wl["n"] * steps / dev, "steps": steps, "ms_per_step": 1e3 * dev / steps,
            "cpu_baseline": {"value": wl["n"] * steps / dev, "unit": "particle-updates/s", "cores": 0,
                             "kind": "reference",
                             "sample": f"{steps} steps of the reference's unmodified CUDA code on this GPU "
                                       "(the reference has no CPU implementation); value = N*steps / "
                                       "(Grid construction + SPH update buckets of simulateAndTime)"},
            "e2e": {"value": wl["n"] * steps / wall, "unit": "particle-updates/s",
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": wl["n"] * 12,
                    "note": "all-in wall clock of simulateAndTime incl. its per-step D2H of N*12 B into "
                            "pageable memory and its grid reset"},
            "reference_buckets_s": {"buildGrid": b[0], "sphUpdate": b[1], "memcpy": b[2]},
        })
        ref.close()
    else:
        cb = cpu_baseline(wl, budget_s=60.0)
        cb["kind"] = "port"
        line.update({"value": cb["value"], "steps": args.steps, "ms_per_step": None, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "particle-updates/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0},
                     "note": "oracle/_ref/libsph_ref.so absent: serial C port used"})
    print(json.dumps(line), flush=True)


# ---- our arm -------------------------------------------------------------------------
def flops_per_particle(C, K):
    # density 9C+6K+3, force 9C+32K, integrate 30
    return (9 * C + 6 * K + 3), (9 * C + 32 * K), 30.0


def slosh_velocity(pos, ids, drift, sigma, y_mid):
    """z-velocity of the multi-GPU workloads: two counter-flowing streams.  The upper half of the
    fluid column (y >= y_mid) drifts towards +z, the lower half towards -z, each as a rigid block, so
    EVERY slab face is crossed in both directions from the first steps on (every rank receives
    migrants) while the lattice structure inside a stream -- the per-particle work of the N=1
    workload -- is preserved; the streams shear past each other at mid-height and pile up against
    the two end walls.  sigma adds a per-particle component that depends on the particle id only (a
    multiplicative hash mapped to a uniform variable of that standard deviation)."""
    vel = np.zeros((len(ids), 3), np.float32)
    vz = np.where(pos[:, 1] >= np.float32(y_mid), drift, -drift).astype(np.float64)
    if sigma:
        u = ((ids.astype(np.uint64) * np.uint64(2654435761)) % np.uint64(2 ** 32)).astype(np.float64) / 2.0 ** 32
        vz = vz + sigma * np.sqrt(3.0) * (2.0 * u - 1.0)
    vel[:, 2] = vz.astype(np.float32)
    return vel


def stretched_lattice(wl, world, rank_range, total=None):
    """Global problem of the multi-GPU runs = the N=1 workload stretched `world` times along z: the
    reference lattice (0.9h spacing from (h,h,h), x outer / z inner; ref: simulator.cu:438-453) with
    the same number of x-planes and world x as many lattice points along z -- one continuous fluid
    column across all slabs.  Returns what `rank_range` = (first global z cell, one past the last)
    owns: positions, ids, and the global particle count and cell layers."""
    h32, box = np.float32(0.1), wl["boxDim"]
    sp = np.float32(0.9) * h32
    nl = int(np.floor((np.float32(box) - np.float32(2) * h32) / sp) + 1)          # lattice points per box edge
    planes = int(np.ceil(wl["n"] / (nl * nl)))                                     # x-planes of the N=1 case
    nlz = (nl - 2) * world
    if total:                                                                      # strong scaling: fixed global problem
        nlz = (nl - 2) * 8
        planes = int(np.ceil(total / (nl * nlz)))
    zs = h32 + sp * np.arange(nlz, dtype=np.float32)
    nz = int(np.floor(zs[-1] / h32)) + 2                                           # + wall layer
    if rank_range is None:
        return None, None, planes * nl * nlz, nz
    zlo, zhi = rank_range(nz)
    cz = (zs / h32).astype(np.int64)
    iz = np.nonzero((cz >= zlo) & (cz < zhi))[0]
    ix, iy = np.arange(planes, dtype=np.int64), np.arange(nl, dtype=np.int64)
    pos = np.empty((planes, nl, len(iz), 3), np.float32)
    pos[..., 0] = (h32 + sp * ix.astype(np.float32))[:, None, None]
    pos[..., 1] = (h32 + sp * iy.astype(np.float32))[None, :, None]
    pos[..., 2] = zs[iz][None, None, :]
    pos = pos.reshape(-1, 3)
    ids = ((ix[:, None, None] * nl + iy[None, :, None]) * nlz + iz[None, None, :]).astype(np.uint32).ravel()
    return pos, ids, planes * nl * nlz, nz


def cluster_parity_check(world, rank, local_rank, nccl_id_fn, gather):
    """A reduced problem (cubic 128-cell box, a lattice block with the sloshing velocities) run
    through the SAME multi-GPU path (world slabs, NCCL) and on one GPU; 20 steps; kinetic energy and
    mean density must agree (SPH trajectories diverge chaotically, north_star: aggregate bound)."""
    import cudafluidsimulator_b200 as sph
    from cudafluidsimulator_b200.cluster import Cluster, partition, slab_ranges
    nc, box, steps = 128, 12.8, 20
    h32, sp = np.float32(0.1), np.float32(0.09)
    g = np.arange(100, dtype=np.float32)
    zs = h32 + sp * np.arange(138, dtype=np.float32)
    x, y, z = np.meshgrid(h32 + sp * g[:40], h32 + sp * g[:60], zs, indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.float32)
    ids = np.arange(len(pos), dtype=np.uint32)
    vel = slosh_velocity(pos, ids, 1.0, 0.5, 2.8)
    st = sph.Settings(numParticles=len(pos), boxDim=box, numCellsPerDim=float(nc))
    cl = Cluster(st, world=world, rank=rank, devices=[local_rank], capacity=len(pos) // world * 2 + 65536,
                 ghost_capacity=65536, emig_capacity=65536, nccl_id=nccl_id_fn(), rebalance_every=8)
    mine = partition(pos, 0.1, slab_ranges(nc, world))[rank]
    cl.load(0, pos[mine], vel[mine], ids[mine])
    cl.advance(steps)
    s = cl.stats(0)
    parts = gather((s["kinetic_energy"], s["density_sum"], s["n_owned"], s["migrated_total"]))
    cl.close()
    if rank != 0:
        return None
    ke = sum(q[0] for q in parts)
    mrho = sum(q[1] for q in parts) / len(pos)
    one = sph.Simulator(st, device=local_rank)
    one.setup()
    one.set_state(pos, vel)
    one.advance(steps)
    ke1, mrho1 = one.get_stats()
    one.close()
    tol_ke, tol_rho = 1e-3, 1e-4
    return {"problem": f"{len(pos)} particles, {nc}^3 cells, {steps} steps, z-sloshing velocities, {world} slabs vs 1 GPU",
            "kinetic_energy": [ke, ke1], "mean_density": [mrho, mrho1],
            "rel_diff": [abs(ke - ke1) / ke1, abs(mrho - mrho1) / mrho1], "tolerance": [tol_ke, tol_rho],
            "particles_conserved": sum(q[2] for q in parts) == len(pos),
            "migrated": int(sum(q[3] for q in parts)),
            "ok": bool(abs(ke - ke1) <= tol_ke * ke1 and abs(mrho - mrho1) <= tol_rho * mrho1
                       and sum(q[2] for q in parts) == len(pos))}


def trace(msg):
    if os.environ.get("SPH_BENCH_TRACE"):
        print(f"[bench rank {os.environ.get('RANK', 0)} +{time.perf_counter():.1f}s] {msg}", file=sys.stderr, flush=True)


def run_cluster(args, wl, rank, local_rank, world):
    """N > 1: one process per GPU (torchrun), one z-slab per process; the slab protocol -- ghost
    halo exchange (pos/vel, then pressure terms), particle migration and rebalancing -- runs inside
    libsph_b200.so over ncclSend/ncclRecv (csrc/sph_cluster.cu).  torch.distributed (gloo) only
    carries the NCCL id, the barriers and the statistics."""
    import torch
    import torch.distributed as dist
    import cudafluidsimulator_b200 as sph
    from cudafluidsimulator_b200.cluster import Cluster, nccl_id, slab_ranges

    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")

    def fresh_id():
        box = [nccl_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    total = args.total if args.scaling == "strong" else None
    rr = lambda nz: slab_ranges(nz, world)[rank]
    my_pos, my_ids, n_glob, nz = stretched_lattice(wl, world, rr, total)
    y_mid = 0.5 * wl["boxDim"]
    my_vel = slosh_velocity(my_pos, my_ids, args.drift, args.sigma, y_mid)
    n = len(my_ids)
    n_max = max(gather(n))
    st = sph.Settings(numParticles=n_glob if n_glob < 2 ** 31 else 2 ** 31 - 1, randomInit=False,
                      boxDim=wl["boxDim"], numCellsPerDim=wl["numCellsPerDim"])
    # a cell layer (0.1) holds one or two lattice planes (0.09 apart); the capacities leave room for more
    plane = int(np.ceil(n_max / max(1.0, (0.1 * nz / world) / 0.09)))
    # (a boundary layer holds at most two planes, and at most one plane crosses a face in a step)
    caps = dict(capacity=int(n_max * 1.15) + 65536, ghost_capacity=3 * plane + 16384, emig_capacity=2 * plane + 16384)

    def make():
        cl = Cluster(st, world=world, rank=rank, devices=[local_rank], nz_cells=nz, nccl_id=fresh_id(),
                     rebalance_every=args.rebalance_every, **caps)
        cl.load(0, my_pos, my_vel, my_ids)
        return cl

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # -- the same per-GPU problem on ONE GPU (one slab, no neighbours): the weak-scaling reference --
    n1 = None
    trace(f"problem built: {n} particles, nz {nz}, caps {caps}")
    if rank == 0 and args.scaling == "weak" and not args.no_n1:
        p1, i1, n1_glob, nz1 = stretched_lattice(wl, 1, lambda nzz: (0, nzz), None)
        # (16 cell layers of headroom at either end: the two streams never reach a wall, i.e. the
        # undisturbed per-particle work -- the conservative reference)
        p1[:, 2] += np.float32(1.6)
        c1 = Cluster(st, world=1, rank=0, devices=[local_rank], nz_cells=nz1 + 32, capacity=int(len(i1) * 1.05) + 65536,
                     ghost_capacity=1024, emig_capacity=1024)
        c1.load(0, p1, slosh_velocity(p1, i1, args.drift, args.sigma, y_mid), i1)
        trace("n1 loaded")
        c1.advance(args.warmup)
        trace("n1 warm")
        n1 = {"particles": int(n1_glob), "ms_per_step": c1.advance_timed(args.steps) / args.steps}
        c1.close()
        del p1, i1
    dist.barrier()
    trace("n1 done")

    # -- device-resident timed region ---------------------------------------------------
    cl = make()
    trace("cluster loaded")
    cl.advance(args.warmup)
    trace("warm")
    start = gather(cl.stats(0))
    l0 = cl.launch_count
    barrier()
    with ClockSampler(local_rank) as clocks:
        local_ms = cl.advance_timed(args.steps)     # CUDA events on the slab's stream, sync at the end
        barrier()
    launches = cl.launch_count - l0
    ms = max(gather(local_ms))
    trace(f"timed: {local_ms / args.steps:.3f} ms/step")
    end = gather({**cl.stats(0), "ms_per_step": local_ms / args.steps})
    cl.close()

    # -- e2e: every step's owned particle records {x, y, z, id} reach pinned host memory ----
    cl = make()
    for _ in range(args.warmup):
        cl.step()
    cl.sync()
    barrier()
    with clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cl.step()
        cl.sync()
        barrier()
        e2e_s = max(gather(time.perf_counter() - t0))
    trace("e2e done")
    rec = cl.host_records(0)
    live = rec[:, 3].view(np.uint32) != 0xFFFFFFFF
    rec_counts = gather((int(len(rec)), int(live.sum())))
    cl.close()

    # -- untimed: imbalance over the run (a third pass, so that the timed regions have no sync inside)
    timeline = None
    if args.timeline:
        cl = make()
        cl.advance(args.warmup)
        timeline = []
        chunks = 4
        for k in range(chunks + 1):
            owned = [q["n_owned"] for q in gather(cl.stats(0))]
            timeline.append({"step": args.warmup + k * (args.steps // chunks),
                             "imbalance_max_over_mean": round(max(owned) / (sum(owned) / world), 5)})
            if k < chunks:
                cl.advance(args.steps // chunks)
        cl.close()

    parity = cluster_parity_check(world, rank, local_rank, fresh_id, gather) if not args.no_parity else None
    if rank == 0:
        owned = [r["n_owned"] for r in end]
        line = {
            "metric": "particle-updates/s", "value": n_glob * args.steps / (ms * 1e-3),
            "unit": "particle-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "n": wl["n"], "boxDim": wl["boxDim"],
                       "numCellsPerDim": wl["numCellsPerDim"], "h": 0.1, "timestep": 0.01,
                       "init": "grid lattice (dam-break column)"},
            "run": {"n_per_gpu": n_glob // world, "n_total": n_glob, "global_cells_z": nz,
                    "init": "N=1 workload stretched along z: one continuous fluid body across all slabs; its upper half "
                            f"streams towards +z and its lower half towards -z at {args.drift} (rigid blocks), plus an "
                            f"id-hashed uniform z-component of std {args.sigma}: every slab face is crossed in both "
                            "directions from the first steps on",
                    "parallelism": f"{world} z-slabs, one process per GPU; ghost halo exchange (pos/vel, then pressure "
                                   "terms) + particle migration per step inside libsph_b200.so: the receiving kernels read "
                                   "the neighbour's message buffers in place over NVLink (CUDA IPC peer memory, seq/ack "
                                   "flags in the message headers; SPH_CLUSTER_NCCL_DATA=1: ncclSend/ncclRecv instead), "
                                   "counts device-resident (no host round trip inside a step); slab boundaries "
                                   f"rebalanced every {args.rebalance_every} steps",
                    "l2": f"state evolves step to step; working set {n * 124 / 1e6:.0f} MB per GPU vs 126 MB L2"},
            "clocks": clocks.summary(),
            "e2e": {"value": n_glob * args.steps / e2e_s, "unit": "particle-updates/s",
                    "ms_per_step": 1e3 * e2e_s / args.steps, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": int(sum(c[0] for c in rec_counts)) * 16,
                    "api": "sph_cluster_step(): one step, then every owned particle's record {x, y, z, id} from the "
                           "slab's GPU into that slab's pinned host buffer (16 B per particle; the copy of step k "
                           "overlaps step k+1; all K copies complete inside the timed region); "
                           "sph_cluster_positions() is the id-ordered getPosition() on demand",
                    "records_last_step": [c[1] for c in rec_counts]},
            "gpu_launches": int(launches) * world,
            "load_balance": {"owned_per_rank_start": [r["n_owned"] for r in start], "owned_per_rank_end": owned,
                             "imbalance_max_over_mean_start": max(r["n_owned"] for r in start) / (sum(r["n_owned"] for r in start) / world),
                             "imbalance_max_over_mean": max(owned) / (sum(owned) / world),
                             "layers_per_rank_end": [[r["z_cell_lo"], r["z_cell_hi"]] for r in end],
                             "rebalances_per_rank": [r["rebalances"] for r in end],
                             "ghosts_last_step": [r["ghosts_lo"] + r["ghosts_hi"] for r in end],
                             "ms_per_step_per_rank": [round(r["ms_per_step"], 4) for r in end],
                             "migrated_total": [int(r["migrated_total"]) for r in end],
                             "timeline": timeline},
            "parity_check": parity,
            "weak_scaling_reference": None if n1 is None else {
                **n1, "updates_per_s": n1["particles"] / (n1["ms_per_step"] * 1e-3),
                "what": "the same per-GPU problem (one band of the column) on one GPU through the same library "
                        "path, timed in this run on rank 0's GPU before the multi-GPU run",
                "efficiency": (n_glob * args.steps / (ms * 1e-3)) / (world * n1["particles"] / (n1["ms_per_step"] * 1e-3))},
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_ours(args, wl, rank, local_rank, world):
    if world > 1:
        return run_cluster(args, wl, rank, local_rank, world)

    import torch
    import cudafluidsimulator_b200 as sph

    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.init()
    key_mode = sph.SPH_KEY_MORTON if args.key == "morton" else sph.SPH_KEY_FLAT
    n = wl["n"]
    st = sph.Settings(numParticles=n, randomInit=wl["randomInit"], boxDim=wl["boxDim"],
                      numCellsPerDim=wl["numCellsPerDim"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)  # -i random: the reference's unseeded glibc state, reproducibly

    # -- device-resident timed region ------------------------------------------------
    t_setup = time.perf_counter()
    sim = sph.Simulator(st, key_mode=key_mode, device=local_rank)
    sim.setup()
    setup_s = time.perf_counter() - t_setup
    sim.advance(args.warmup)
    l0 = sim.launch_count
    barrier()
    with ClockSampler(local_rank) as clocks:
        ms = sim.advance_timed(args.steps)
        barrier()
    launches = sim.launch_count - l0
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * n * args.steps / (ms * 1e-3)

    # -- stage times + neighbour statistics for the rooflines (state after the timed region)
    K, C = sim.get_neighbor_counts()
    meanK, meanC = float(K.mean()), float(C.mean())
    table_size = sim.table_size
    sort_info = sim.sort_info()   # the library says how it sorts: counting sort by cell, or radix passes
    sort_passes = sort_info["radix_passes"]
    prof_steps = max(3, min(10, args.steps))
    sim.advance(1)   # (the neighbour-count query rebuilt the grid: one step brings the fused count back)
    sim.profile_enable(True)
    sim.profile_read(reset=True)
    sim.advance(prof_steps)
    prof = sim.profile_read(reset=True)
    sim.profile_enable(False)
    stage_ms = {k: v["ms"] / prof_steps for k, v in prof.items() if v["launches"]}
    sim.close()

    # -- BASELINE configs[2] names "Morton-sorted cells": the same steps with Morton keys, same run
    morton = None
    if args.key == "flat" and not args.no_morton:
        libc.srand(1)
        sm = sph.Simulator(st, key_mode=sph.SPH_KEY_MORTON, device=local_rank)
        sm.setup()
        sm.advance(args.warmup)
        mms = sm.advance_timed(args.steps)
        sm.close()
        morton = {"ms_per_step": mms / args.steps, "value": n * args.steps / (mms * 1e-3),
                  "relative_to_flat": (mms / args.steps) / (ms / args.steps),
                  "why": "Morton order breaks the 3 x-neighbours of a row into separate runs (27 single-cell runs "
                         "per stencil instead of 9 contiguous x-runs), see profiles/r02_morton_vs_flat.md"}

    # -- the step's sort: the default (counting sort by cell) against the 8-bit radix passes, same run
    sort_radix = None
    if sort_info["algo"] == "count" and not args.no_morton:
        libc.srand(1)
        sr = sph.Simulator(st, key_mode=key_mode, device=local_rank, sort_algo=sph.SPH_SORT_RADIX)
        sr.setup()
        sr.advance(args.warmup)
        rms = sr.advance_timed(args.steps)
        sr.profile_enable(True)
        sr.profile_read(reset=True)
        sr.advance(prof_steps)
        rprof = sr.profile_read(reset=True)
        sr.close()
        sort_radix = {"ms_per_step": rms / args.steps, "value": n * args.steps / (rms * 1e-3),
                      "relative_to_default": (rms / args.steps) / (ms / args.steps),
                      "stages_ms": {k: round(v["ms"] / prof_steps, 4) for k, v in rprof.items()
                                    if v["launches"] and k in ("histogram", "sort_passes", "reorder_cellstart")},
                      "default_stages_ms": {k: round(v, 4) for k, v in stage_ms.items()
                                            if k in ("histogram", "sort_passes", "reorder_cellstart")}}

    # -- e2e: same workload through sph_step() with the per-step D2H of positions --------
    def e2e_run(pipeline):
        libc.srand(1)
        sim = sph.Simulator(st, key_mode=key_mode, device=local_rank, pipeline_readback=pipeline)
        sim.setup()
        for _ in range(args.warmup):
            sim.simulate()
        barrier()
        with clocks:
            t0 = time.perf_counter()
            for _ in range(args.steps):
                sim.simulate()
            barrier()
            dt = time.perf_counter() - t0
        chk = float(sim.getPosition()[:: max(1, n // 1000)].sum())
        sim.close()
        return dt, chk
    e2e_blocking_s, checksum_blocking = e2e_run(False)
    e2e_s, checksum = e2e_run(True)
    if rank != 0:
        return
    hbm_peak, sm_max_mhz, peak_kind = measured_peaks()
    clk = clocks.summary()
    f_den, f_force, f_int = flops_per_particle(meanC, meanK)
    passes = sort_passes          # 8-bit onesweep passes a radix sort runs for this key range
    cells_per_particle = table_size / n
    if sort_info["algo"] == "count":
        # counting sort by cell: table scan (sums: R 4; apply: R 4 + W 4 cell_start + W 4 cleared count, per
        # cell) + scatter (R 8 tagged + R 4 cell_start + W 8 pair); the count itself is part of the force
        # kernel (+ 8 B tagged pair per particle there) unless it runs as the "histogram" stage (R 4 + W 8)
        sort_bytes = 16 * cells_per_particle + 20
        hist_bytes = 12
        reorder_bytes = 8 + 32 + 32 + 12              # pair, gather, sorted copy, pair-interleaved copy (the cell's
        #                                               range and members are re-reads of lines the warp already has)
        force_extra = 8 if sort_info["count_fused"] else 0
    else:
        sort_bytes = 4 + 8 + 16 * (passes - 1)
        hist_bytes = 4
        reorder_bytes = 8 + 32 + 32 + 12 + 4 * cells_per_particle
        force_extra = 0
    bytes_stage = {  # algorithmic bytes per particle
        "hash": 20, "histogram": hist_bytes, "sort_passes": sort_bytes,
        "reorder_cellstart": reorder_bytes,
        "density": 16 + 12 + 8, "force_integrate": 16 + 16 + 8 + 4 + 8 + 16 + 16 + 4 + 12 + force_extra,
    }
    ncu_kernel = {"density": "k_density_flat", "force_integrate": "k_force_integrate_flat",
                  "reorder_cellstart": "k_reorder", "sort_passes": "k_onesweep<0>", "histogram": "k_histogram"}
    if sort_info["algo"] == "count":
        ncu_kernel.update({"sort_passes": "k_cell_scatter", "histogram": "k_cell_count"})

    def kernel_traffic(stage):   # DRAM bytes per launch of that stage's kernel from the committed ncu capture
        if not traffic or stage not in ncu_kernel:
            return None
        for name, b in traffic["bytes_per_launch"].items():
            if name.startswith(ncu_kernel[stage]):
                return b
        return None
    # DRAM bytes per launch from the committed ncu --set full captures of the same workload; two
    # states are on file (lattice: up to ~step 40; floor pile-up: around step 100) and the one that
    # matches the state of the timed kernels is used
    traffic = None
    state_step = args.warmup + args.steps
    tfile = ROOT / "profiles" / ("r02_traffic_early.json" if state_step <= 40 else "r02_traffic_late.json")
    if tfile.exists() and args.workload == "16m_grid" and args.key == "flat":
        traffic = json.loads(tfile.read_text())
    flops_stage = {"density": f_den, "force_integrate": f_force + f_int}
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    fp32_peak_tflops = sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    stages = {}
    for name, t_ms in stage_ms.items():
        e = {"ms": round(t_ms, 4), "share": round(t_ms / sum(stage_ms.values()), 4)}
        if name in bytes_stage:
            e["algorithmic_GBps"] = round(bytes_stage[name] * n / (t_ms * 1e-3) / 1e9, 1)
            e["hbm_frac"] = round(e["algorithmic_GBps"] / hbm_peak, 4)
        if name in flops_stage:
            e["algorithmic_TFLOPs"] = round(flops_stage[name] * n / (t_ms * 1e-3) / 1e12, 3)
            e["fp32_frac"] = round(e["algorithmic_TFLOPs"] / fp32_peak_tflops, 4)
        stages[name] = e
    dom = max(stage_ms, key=stage_ms.get)
    d = stages[dom]
    if dom in flops_stage:
        roof = {"kernel": dom, "bound": "fp32", "achieved": d["algorithmic_TFLOPs"], "peak": round(fp32_peak_tflops, 2),
                "unit": "TFLOP/s", "frac": d["fp32_frac"],
                "traffic": kernel_traffic(dom),
                "traffic_source": traffic["source"] if traffic else None,
                "algorithmic_bytes_per_launch": bytes_stage[dom] * n,
                "peak_source": f"SMs*128 lanes*2*clocks.max.sm ({sm_count} SMs, {sm_max_mhz:.0f} MHz from "
                               f"MEASURED_PEAKS.json, {peak_kind}); no tensor cores on this path",
                "flops_per_particle": round(flops_stage[dom], 1), "mean_candidates_C": round(meanC, 2),
                "mean_neighbours_K": round(meanK, 2), "hbm_GBps": d.get("algorithmic_GBps"),
                "hbm_frac": d.get("hbm_frac")}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": d["algorithmic_GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": d["hbm_frac"],
                "traffic": kernel_traffic(dom),
                "traffic_source": traffic["source"] if traffic else None,
                "algorithmic_bytes_per_launch": bytes_stage[dom] * n,
                "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})"}

    line = {
        "metric": "particle-updates/s", "value": value, "unit": "particle-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.workload, wl),
        "run": {"key": args.key, "parallelism": "single GPU", "sort": sort_info,
                "state": f"steps {args.warmup + 1}..{args.warmup + args.steps} from the initial condition",
                "density_sum": "factored (library default)",
                "l2": "each step consumes the previous step's output (no repeated input); "
                      f"working set {n * 108 / 1e6:.0f} MB vs 126 MB L2"},
        "clocks": clk,
        "e2e": {"value": world * n * args.steps / e2e_s, "unit": "particle-updates/s",
                "ms_per_step": 1e3 * e2e_s / args.steps, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": n * 12, "setup_h2d_bytes": n * 32, "setup_s": round(setup_s, 3),
                "api": "sph_step() == Simulator::simulate() with SphOptions.pipeline_readback: every call returns "
                       "step k's positions in pinned host memory; their D2H overlaps the computation of step k+1",
                "blocking": {"value": world * n * args.steps / e2e_blocking_s,
                             "ms_per_step": 1e3 * e2e_blocking_s / args.steps,
                             "api": "sph_step() without overlap: step, then D2H, then return"},
                "checksum": checksum, "checksum_matches_blocking": checksum == checksum_blocking},
        "gpu_launches": int(launches),
        "roofline": roof,
        "stages": stages,
        "key_morton": morton,
        "sort_radix": sort_radix,
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(wl)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_variants(args, wl):
    """BASELINE.json configs[1]: neighbour search of this repo against all three reference variants
    on one GPU -- the real lock-free linked lists, and the two sorted variants RECONSTRUCTED from the
    README (their source is not in the reference tree; parity unpinned)."""
    import ctypes
    import torch
    import cudafluidsimulator_b200 as sph
    from oracle.oracle import RECON_SO, REF_SO, ReconSim, RefSim
    torch.cuda.init()
    kw = dict(boxDim=wl["boxDim"], numCellsPerDim=wl["numCellsPerDim"], randomInit=wl["randomInit"])
    n, steps = wl["n"], args.steps
    rows = {}

    def buckets_row(b, label):
        rows[label] = {"grid_ms": 1e3 * b[0] / steps, "sph_update_ms": 1e3 * b[1] / steps,
                       "transfer_ms": 1e3 * b[2] / steps,
                       "particle_updates_per_s": n * steps / (b[0] + b[1])}
    if REF_SO.exists():
        r = RefSim(n, **kw)
        for _ in range(steps):
            b, _ = r.step_timed()
        buckets_row(b, "reference linked lists (main branch, unmodified)")
        r.close()
    if RECON_SO.exists():
        for morton, label in ((False, "index_sort (RECONSTRUCTED from README, not reference source)"),
                              (True, "z_index_sort (RECONSTRUCTED from README, not reference source)")):
            r = ReconSim(n, morton, **kw)
            for _ in range(steps):
                b = r.step_timed()
            buckets_row(b, label)
            r.close()
    for mode, label in ((sph.SPH_KEY_FLAT, "this repo, flat keys"), (sph.SPH_KEY_MORTON, "this repo, Morton keys")):
        ctypes.CDLL("libc.so.6").srand(1)
        sim = sph.Simulator(sph.Settings(numParticles=n, randomInit=wl["randomInit"], boxDim=wl["boxDim"],
                                         numCellsPerDim=wl["numCellsPerDim"]), key_mode=mode)
        sim.setup()
        t = sph.Times()
        for _ in range(steps):
            sim.simulateAndTime(t)
        buckets_row((t.buildGrid, t.sphUpdate, t.memcpy), label)
        sim.close()
    print(json.dumps({"comparison": "neighbour-search variants, simulateAndTime() buckets (host wall clock, "
                                    "launch + sync inside each bucket, as the reference measures them)",
                      "workload": args.workload, "n": n, "steps": steps, "variants": rows}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--compare-variants", action="store_true",
                    help="config 2: time the three reference neighbour-search variants beside this repo")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS),
                    help="default: 16m_grid on one GPU (BASELINE configs[2]); 32m_grid per GPU on several "
                         "(BASELINE configs[4], the north_star weak-scaling point)")
    ap.add_argument("--key", default="flat", choices=["flat", "morton"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-morton", action="store_true", help="skip the Morton-key run of the same steps")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = the N=1 workload per GPU (default); strong = --total particles split over the GPUs")
    ap.add_argument("--total", type=int, default=64_000_000, help="global particle count for --scaling strong")
    ap.add_argument("--drift", type=float, default=1.0, help="N > 1: z-speed of the two counter-flowing halves of the fluid column")
    ap.add_argument("--sigma", type=float, default=0.0, help="N > 1: std of the per-particle z-velocity component")
    ap.add_argument("--no-n1", action="store_true", help="N > 1: skip the one-GPU run of the same per-GPU problem")
    ap.add_argument("--rebalance-every", type=int, default=16, help="N > 1: steps between slab boundary moves (0 = static)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the reduced-problem parity check")
    ap.add_argument("--timeline", action="store_true", help="N > 1: extra untimed pass reporting imbalance over the run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, local_rank, world = dist_env()
    if args.workload is None:
        args.workload = "16m_grid" if world == 1 else "32m_grid"   # both arms
    wl = WORKLOADS[args.workload]
    if args.compare_variants:
        if rank == 0:
            run_variants(args, wl)
    elif args.impl == "reference":
        run_reference(args, wl, rank, world)
    else:
        run_ours(args, wl, rank, local_rank, world)


if __name__ == "__main__":
    main()
