"""Development probe (torchrun, one rank per GPU): pinned host -> device bandwidth of all ranks at the same time,
with and without binding the rank to the CPUs next to its GPU.  Answers whether the end-to-end number at N GPUs is
limited by the host side.   python -m torch.distributed.run --nproc-per-node N tools/h2d_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tissue_image_processing_b200 import _native as nat   # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")


def measure(tag):
    host = torch.empty(512 << 20, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    dev = torch.empty_like(host, device="cuda")
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([10 * host.numel() / dt / 1e9], dtype=torch.float64)
    if world > 1:
        allg = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(allg, gbs)
    else:
        allg = [gbs]
    if rank == 0:
        vals = [float(g) for g in allg]
        print("%s: per-rank GB/s %s  total %.1f" % (tag, " ".join("%.1f" % v for v in vals), sum(vals)), flush=True)


measure("unbound (%d CPUs)" % len(os.sched_getaffinity(0)))
cpus = nat.bind_host_thread_to_gpu(local)
if world > 1:
    info = [None] * world
    dist.all_gather_object(info, (local, None if cpus is None else (cpus[0], cpus[-1], len(cpus))))
    if rank == 0:
        print("NVML affinity per rank (first, last, count):", info, flush=True)
measure("bound")
if world > 1:
    dist.destroy_process_group()
