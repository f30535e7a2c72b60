"""Development probe: project N device-resident frames of one shape, one after the other (chained launches), and print
the per-frame time.  Run it under ncu for a launch list of a given shape:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/frame_probe.py C Z Y X [N] [mode]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402
from tissue_image_processing_b200 import _native as nat        # noqa: E402

C, Z, Y, X = (int(v) for v in sys.argv[1:5])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 4
mode = sys.argv[6] if len(sys.argv) > 6 else "fast"
dev = torch.device("cuda", 0)
frames = [bench.synth_frame_device(torch, 70 + i, dev, (Z, Y, X), C) for i in range(2)]
p = nat.DeviceProjector(C, Z, Y, X, airyscan=False, mode=mode, device=0)
for f in frames * 2:
    p.run(f)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    p.run(frames[i % 2])
e1.record()
torch.cuda.synchronize()
print("%dx%dx%dx%d %s: %.1f us / frame" % (C, Z, Y, X, mode, 1e3 * e0.elapsed_time(e1) / n), p.status())
