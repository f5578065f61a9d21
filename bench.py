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
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

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
            "config": {"workload": args.workload, **{k: wl[k] for k in ("n", "boxDim", "numCellsPerDim")},
                       "init": "random(glibc rand seed 1)" if wl["randomInit"] else "grid lattice"}}
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
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
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


def lattice_positions(n, h, box, z_shift=0.0):
    """The reference's grid init (ref: simulator.cu:438-453) vectorised: float32 multiply and
    add, separately rounded, x outer / z inner; optional shift along z for replicated sub-boxes."""
    h32, sp = np.float32(h), np.float32(0.9) * np.float32(h)
    nx = int(np.floor((np.float32(box) - np.float32(2) * h32) / sp) + 1)
    i = np.arange(n, dtype=np.int64)
    ix, iy, iz = i // (nx * nx), (i // nx) % nx, i % nx
    pos = np.empty((n, 3), np.float32)
    pos[:, 0] = h32 + sp * ix.astype(np.float32)
    pos[:, 1] = h32 + sp * iy.astype(np.float32)
    pos[:, 2] = h32 + sp * iz.astype(np.float32)
    if z_shift:
        pos[:, 2] += np.float32(z_shift)
    return pos


def run_slabs(args, wl, rank, local_rank, world):
    """N > 1: weak scaling, one slab per GPU.  The N=1 workload (sub-box) is replicated along
    z (the decomposition axis); the interfaces between sub-boxes are open, so every step does a
    real ghost halo exchange (pos/vel, then pressure terms) and particle migration over NCCL."""
    import torch
    import torch.distributed as dist
    import cudafluidsimulator_b200 as sph
    from cudafluidsimulator_b200.slab import SlabBackend, SlabDriver

    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from cudafluidsimulator_b200.slab import partition, slab_ranges
    nc = int(wl["numCellsPerDim"])
    h32, box = np.float32(0.1), wl["boxDim"]
    st = sph.Settings(numParticles=wl["n"], randomInit=wl["randomInit"], boxDim=box,
                      numCellsPerDim=wl["numCellsPerDim"])
    # Global problem = the N=1 workload stretched `world` times along z:
    #   grid  : the reference lattice (0.9h spacing from (h,h,h), x outer / z inner) with the same
    #           number of x-planes as the N=1 case and world x as many lattice points along z --
    #           one continuous fluid column across all slabs (real halos, real migration)
    #   random: uniform in the stretched box interior
    if wl["randomInit"]:
        nz = nc * world
        ranges = slab_ranges(nz, world)
        rng = np.random.default_rng(1234)
        n_glob = wl["n"] * world
        # every rank draws the same global set and keeps its slab (cheap at 1M x world)
        gpos = rng.uniform(1.0, box - 1.0, (n_glob, 3)).astype(np.float32)
        gpos[:, 2] = (np.float32(1.0) + rng.uniform(0, 1, n_glob).astype(np.float32) * np.float32(box * world - 2.0))
        mine = partition(gpos, 0.1, ranges)[rank]
        my_pos, my_ids = gpos[mine], mine.astype(np.uint32)
    else:
        sp = np.float32(0.9) * h32
        nl = int(np.floor((np.float32(box) - np.float32(2) * h32) / sp) + 1)     # 283 lattice points per box edge
        planes = int(np.ceil(wl["n"] / (nl * nl)))                                # x-planes of the N=1 case
        # two lattice points fewer per slab than the N=1 box: keeps (owned + 2 ghost) layers x nc^2
        # below 2^24 keys, i.e. the same three 8-bit sort passes as on one GPU
        nlz = (nl - 2) * world
        if args.scaling == "strong":
            # BASELINE.json configs[3]: a fixed global problem (default 64M particles, the dam-break
            # slab stretched 8 box lengths along z) split over however many GPUs there are
            nlz = (nl - 2) * 8
            planes = int(np.ceil(args.total / (nl * nlz)))
        zs = h32 + sp * np.arange(nlz, dtype=np.float32)
        nz = int(np.floor(zs[-1] / h32)) + 2                                       # + wall layer
        ranges = slab_ranges(nz, world)
        cz = (zs / h32).astype(np.int64)
        iz = np.nonzero((cz >= ranges[rank][0]) & (cz < ranges[rank][1]))[0]
        ix, iy, izz = np.meshgrid(np.arange(planes), np.arange(nl), iz, indexing="ij")
        my_pos = np.empty((ix.size, 3), np.float32)
        my_pos[:, 0] = (h32 + sp * ix.astype(np.float32)).ravel()
        my_pos[:, 1] = (h32 + sp * iy.astype(np.float32)).ravel()
        my_pos[:, 2] = zs[izz.ravel()]
        my_ids = ((ix.ravel().astype(np.int64) * nl + iy.ravel()) * nlz + izz.ravel()).astype(np.uint32)
        n_glob = planes * nl * nlz
    n = len(my_ids)
    zlo, zhi = ranges[rank]

    def make():
        b = SlabBackend(st, zlo, zhi, nz, capacity=int(n * 1.2) + 4096, device=local_rank,
                        ghost_capacity=int(n * 0.1) + 4096, emig_capacity=int(n * 0.05) + 4096)
        b.load(my_pos, np.zeros_like(my_pos), my_ids)
        return b, SlabDriver(b, rank, world, overlap=not args.no_overlap)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        local_ms = e0.elapsed_time(e1)
        timed.local_ms = local_ms
        t = torch.tensor([local_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # -- device-resident timed region ---------------------------------------------------
    b, drv = make()
    for _ in range(args.warmup):
        drv.step()
    l0 = b.launch_count
    with ClockSampler(local_rank) as clocks:
        ms = timed(drv.step, args.steps)
    launches = b.launch_count - l0
    if drv.host_ms is not None and rank == 0:   # SPH_SLAB_TRACE=1: host time per protocol phase
        tot = args.steps + args.warmup
        print("host ms/step by phase (rank 0): " +
              ", ".join(f"{k} {v / tot:.3f}" for k, v in drv.host_ms.items()), file=sys.stderr)
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {"rank": rank, "ms_per_step": timed.local_ms / args.steps, **drv.last,
                                      **{k: v for k, v in drv.stats.items()}})
    b.close()

    # -- e2e: every step also copies the owned particles' positions to pinned host memory --
    b, drv = make()
    host = torch.empty((b.capacity, 3), dtype=torch.float32).pin_memory()     # xyz, as the reference's float3
    stage = torch.empty((b.capacity, 3), dtype=torch.float32, device=b.device)
    copy_stream = torch.cuda.Stream()
    copy_done = torch.cuda.Event()
    copy_done.record()

    def step_e2e():
        # every step's owned positions reach pinned host memory; the PCIe copy of step k runs on a
        # side stream from a device staging copy while step k+1 computes
        drv.step()
        k = drv.last["n_owned"]
        copy_done.synchronize()                  # staging buffer free again
        stage[:k].copy_(b.cur_pos[:k, :3])       # drop the id word: 12 B per particle cross PCIe
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            host[:k].copy_(stage[:k], non_blocking=True)
            copy_done.record()
    for _ in range(args.warmup):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    b.close()

    if rank == 0:
        owned = [r["n_owned"] for r in per_rank]
        line = {
            "metric": "particle-updates/s", "value": n_glob * args.steps / (ms * 1e-3),
            "unit": "particle-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "n_per_gpu": n_glob // world, "n_total": n_glob, "boxDim": wl["boxDim"],
                       "numCellsPerDim": wl["numCellsPerDim"], "global_cells_z": nz,
                       "parallelism": f"{world} z-slabs, one per GPU: ghost halo exchange (pos/vel, then pressure "
                                      "terms) + particle migration per step over NCCL send/recv",
                       "init": "N=1 workload stretched along z: one continuous fluid body across all slabs",
                       "l2": f"state evolves step to step; working set {n * 124 / 1e6:.0f} MB per GPU vs 126 MB L2"},
            "clocks": clocks.summary(),
            "e2e": {"value": n_glob * args.steps / (e2e_ms * 1e-3), "unit": "particle-updates/s",
                    "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": n_glob * 12,
                    "api": "SlabDriver.step() + per-step D2H of every owned particle's position record into pinned host "
                           "memory (copy of step k overlaps the computation of step k+1)"},
            "gpu_launches": int(launches) * world,
            "load_balance": {"owned_per_rank": owned, "imbalance_max_over_mean": max(owned) / (sum(owned) / world),
                             "ghosts_last_step": [r["ghosts"] for r in per_rank],
                             "ms_per_step_per_rank": [round(r["ms_per_step"], 4) for r in per_rank],
                             "migrated_total": [r["migrated_particles"] for r in per_rank]},
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_ours(args, wl, rank, local_rank, world):
    if world > 1:
        return run_slabs(args, wl, rank, local_rank, world)

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
    prof_steps = max(3, min(10, args.steps))
    sim.profile_enable(True)
    sim.profile_read(reset=True)
    sim.advance(prof_steps)
    prof = sim.profile_read(reset=True)
    sim.profile_enable(False)
    stage_ms = {k: v["ms"] / prof_steps for k, v in prof.items() if v["launches"]}
    sim.close()

    # -- e2e: same workload through sph_step() with the per-step D2H of positions --------
    def e2e_run(pipeline):
        libc.srand(1)
        sim = sph.Simulator(st, key_mode=key_mode, device=local_rank, pipeline_readback=pipeline)
        sim.setup()
        for _ in range(args.warmup):
            sim.simulate()
        barrier()
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
    passes = 3
    bytes_stage = {  # algorithmic bytes per particle
        "hash": 20, "histogram": 4, "sort_passes": 4 + 8 + 16 * (passes - 1),
        "reorder_cellstart": 8 + 32 + 32 + 12 + 4 * (wl["numCellsPerDim"] ** 3) / n,
        "density": 16 + 12 + 8, "force_integrate": 16 + 16 + 8 + 4 + 8 + 16 + 16 + 4 + 12,
    }
    ncu_kernel = {"density": "k_density_flat<0", "force_integrate": "k_force_integrate_flat",
                  "reorder_cellstart": "k_reorder", "sort_passes": "k_onesweep<0>", "histogram": "k_histogram"}

    def kernel_traffic(stage):   # DRAM bytes per launch of that stage's kernel from the committed ncu capture
        if not traffic or stage not in ncu_kernel:
            return None
        for name, b in traffic["bytes_per_launch"].items():
            if name.startswith(ncu_kernel[stage]):
                return b
        return None
    traffic = None
    tfile = ROOT / "profiles" / "r01_traffic.json"
    if tfile.exists() and args.workload == "16m_grid":
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
        "config": {"workload": args.workload, "n_per_gpu": n, "boxDim": wl["boxDim"],
                   "numCellsPerDim": wl["numCellsPerDim"], "h": 0.1, "timestep": 0.01,
                   "init": "random(glibc rand seed 1)" if wl["randomInit"] else "grid lattice (dam-break column)",
                   "key": args.key, "parallelism": "single GPU" if world == 1 else f"{world} independent replicas",
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
    ap.add_argument("--workload", default="16m_grid", choices=list(WORKLOADS))
    ap.add_argument("--key", default="flat", choices=["flat", "morton"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = the N=1 workload per GPU (default); strong = --total particles split over the GPUs")
    ap.add_argument("--total", type=int, default=64_000_000, help="global particle count for --scaling strong")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N > 1: halo exchanges and count round trips after, not under, the interior CTAs (A/B switch)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, local_rank, world = dist_env()
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
