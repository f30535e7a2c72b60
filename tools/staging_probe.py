"""Development probe: host copy rate of a pageable 1024x1024x48 uint16 frame into a pinned staging buffer
(movie._Staging.stage) for several thread counts, and plain numpy / torch copies for comparison.
    python tools/staging_probe.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tissue_image_processing_b200 import _native as nat        # noqa: E402
from tissue_image_processing_b200.movie import _Staging        # noqa: E402

nat.handle(0)
shape = (1, 48, 1024, 1024)
rng = np.random.default_rng(0)
frames = [rng.integers(0, 4000, size=shape).astype(np.uint16) for _ in range(4)]
nbytes = frames[0].nbytes
print("cpus", len(os.sched_getaffinity(0)), "frame MiB", nbytes >> 20, flush=True)
for threads in (1, 2, 4, 8, 12, 16, 24):
    st = _Staging(2, threads)
    for i in range(4):
        st.stage(i % 2, frames[i % 4])
    n = 24
    t0 = time.perf_counter()
    for i in range(n):
        st.stage(i % 2, frames[i % 4])
    dt = (time.perf_counter() - t0) / n
    st.close()
    print("threads %2d: %.2f ms / frame  %.1f GB/s" % (threads, dt * 1e3, nbytes / dt / 1e9), flush=True)
# torch's own multi-threaded copy into pinned memory
pin = torch.empty(shape, dtype=torch.uint16, pin_memory=True)
src = [torch.from_numpy(f) for f in frames]
for nt in (8, 16):
    torch.set_num_threads(nt)
    for i in range(4):
        pin.copy_(src[i % 4])
    t0 = time.perf_counter()
    for i in range(24):
        pin.copy_(src[i % 4])
    dt = (time.perf_counter() - t0) / 24
    print("torch copy_ %2d threads: %.2f ms / frame  %.1f GB/s" % (nt, dt * 1e3, nbytes / dt / 1e9), flush=True)
# cudaMemcpyAsync straight from pageable memory
dev = torch.empty(shape, dtype=torch.uint16, device="cuda")
for i in range(4):
    dev.copy_(src[i % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(24):
    dev.copy_(src[i % 4])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 24
print("pageable -> device (driver staging): %.2f ms / frame  %.1f GB/s" % (dt * 1e3, nbytes / dt / 1e9), flush=True)
t0 = time.perf_counter()
for i in range(24):
    dev.copy_(pin, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 24
print("pinned -> device: %.2f ms / frame  %.1f GB/s" % (dt * 1e3, nbytes / dt / 1e9), flush=True)
