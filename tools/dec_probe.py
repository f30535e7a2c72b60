"""Per-stage device times of one config-2 frame (bench.py's synthetic frame unless --uniform)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tissue_image_processing_b200 import _native as nat
Z, Y, X = 64, 2048, 2048
if "--uniform" in sys.argv:
    stack = (torch.rand((1, Z, Y, X), device="cuda") * 3000).to(torch.uint16)
else:
    stack = bench.synth_frame_device(torch, 2, torch.device("cuda"))
p = nat.DeviceProjector(1, Z, Y, X, mode="fast")
nat.set_profiling(True)
for _ in range(3): p.run(stack)
nat.stage_times(reset=True)
for _ in range(10): p.run(stack)
st = nat.stage_times(reset=True)
print({k: round(v[0] / v[1], 4) for k, v in st.items()}, "total", round(sum(v[0] / v[1] for v in st.values()), 4))
