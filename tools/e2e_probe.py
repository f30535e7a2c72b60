"""Development probe (torchrun, one rank per GPU): end-to-end frames of 2048x2048x64 through FramePipeline from pinned
host memory for several slot counts / output dtypes; prints total GB/s over the ranks.
    python -m torch.distributed.run --nproc-per-node N tools/e2e_probe.py [frames_per_rank]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402
from tissue_image_processing_b200 import _native as nat        # noqa: E402
from tissue_image_processing_b200 import movie as mv            # noqa: E402
from tissue_image_processing_b200 import topology               # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
index, topo = topology.choose_device(local, world)
torch.cuda.set_device(index)
if world > 1:
    dist.init_process_group("gloo")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Z, Y, X = 64, 2048, 2048
dev = torch.device("cuda", index)
host = []
for i in range(2):
    h = nat.pinned_empty((1, Z, Y, X), np.uint16)
    torch.from_numpy(h).copy_(bench.synth_frame_device(torch, 10 * rank + i, dev))
    host.append(h)
torch.cuda.synchronize()


def run(slots, out_dtype, total):
    pipe = mv.FramePipeline(devices=[index], slots=slots, mode="fast", out_dtype=out_dtype)
    counter = mv.SharedFrameCounter("probe")
    seen = [0]

    def sink(k, p, z, st):
        seen[0] += 1

    def gen():
        for i in counter.claims(total):
            yield i, host[i % 2]
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pipe.project_frames(gen(), sink, reference_channel=0, airyscan=False)
    if world > 1:
        dist.barrier()
    return time.perf_counter() - t0, seen[0]


run(2, "reference", 2 * world)
for slots, od in ((2, "reference"), (3, "reference"), (4, "reference"), (2, "uint16"), (3, "uint16")):
    dt, mine = run(slots, od, n * world)
    counts = [None] * world
    if world > 1:
        dist.all_gather_object(counts, mine)
    if rank == 0:
        gb = n * world * Z * Y * X * 2 / 1e9
        print("slots %d out %-9s: %.1f GB/s H2D total, %.2f ms/frame/rank, frames per rank %s" %
              (slots, od, gb / dt, dt * 1e3 / n, counts if world > 1 else mine), flush=True)
if world > 1:
    dist.destroy_process_group()
