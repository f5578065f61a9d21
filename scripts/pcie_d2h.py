#!/usr/bin/env python
"""Device-to-host copy ceiling with every GPU of the box copying at once (torchrun, one process per
GPU): what the per-step position readback of the multi-GPU e2e number can reach at best.
Prints one line per rank: GB/s of back-to-back 256 MB cudaMemcpyAsync D2H into pinned memory."""
import os
import time

import torch
import torch.distributed as dist

rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
if world > 1:
    dist.init_process_group("gloo")
torch.cuda.set_device(local)
nbytes = 256 << 20
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
for _ in range(3):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 20
t0 = time.perf_counter()
for _ in range(reps):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = reps * nbytes / dt / 1e9
out = [None] * world
if world > 1:
    dist.all_gather_object(out, gbs)
else:
    out = [gbs]
if rank == 0:
    print({"d2h_GBps_per_gpu_all_copying": [round(g, 1) for g in out], "sum": round(sum(out), 1), "gpus": world}, flush=True)
if world > 1:
    dist.destroy_process_group()
