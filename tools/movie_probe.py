"""Development probe: device-resident frames/s of a frame shape for several numbers of frames in flight.
    python tools/movie_probe.py Z Y X [streams ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402
from tissue_image_processing_b200 import _native as nat        # noqa: E402

Z, Y, X = (int(v) for v in sys.argv[1:4])
streams = [int(v) for v in sys.argv[4:]] or [1, 2, 3, 4, 6, 8]
dev = torch.device("cuda", 0)
frames = [bench.synth_frame_device(torch, 70 + i, dev, (Z, Y, X)) for i in range(4)]
p0 = nat.DeviceProjector(1, Z, Y, X, airyscan=False, mode="fast", device=0)
for i in range(4):
    p0.run(frames[i % 4])
torch.cuda.synchronize()
nat.set_profiling(True, 0)
nat.stage_times(reset=True, device=0)
for i in range(20):
    p0.run(frames[i % 4])
torch.cuda.synchronize()
st = nat.stage_times(reset=True, device=0)
nat.set_profiling(False, 0)
print("stage us (events between the stages):", {k: round(1e3 * v[0] / max(v[1], 1), 1) for k, v in st.items()}, flush=True)
for ns in streams:
    projs = [nat.DeviceProjector(1, Z, Y, X, airyscan=False, mode="fast", device=0, concurrent=ns > 1) for _ in range(ns)]
    strs = [torch.cuda.Stream(device=dev) for _ in range(ns)]

    def run(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in strs:
            st.wait_event(e0)
        for i in range(k):
            with torch.cuda.stream(strs[i % ns]):
                projs[i % ns].run(frames[i % 4])
        for st in strs:
            ev = torch.cuda.Event()
            ev.record(st)
            torch.cuda.current_stream().wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    torch.cuda.synchronize()
    for _ in range(2):                   # every (projector, frame) pair twice: graphs captured before timing
        for f in frames:
            for k in range(ns):
                with torch.cuda.stream(strs[k]):
                    projs[k].run(f)
    torch.cuda.synchronize()
    run(2 * ns)
    n = 96
    ms = min(run(n) for _ in range(3))
    print("%dx%dx%d  %d in flight: %.1f us / frame  %.0f frames/s  %.0f Gvoxel/s" % (
        Z, Y, X, ns, 1e3 * ms / n, n / ms * 1e3, n * Z * Y * X / ms / 1e6), flush=True)
