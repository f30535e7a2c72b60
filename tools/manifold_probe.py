"""Development probe: device time of tsp_build_manifold (SP:87-165) on a smooth synthetic score volume.
    python tools/manifold_probe.py [Z Y X]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tissue_image_processing_b200 import _native as nat        # noqa: E402

Z, Y, X = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 2048, 2048)
dev = torch.device("cuda", 0)
nat.handle(0)
g = torch.Generator(device=dev).manual_seed(3)
yy = torch.arange(Y, device=dev, dtype=torch.float32)[:, None]
xx = torch.arange(X, device=dev, dtype=torch.float32)[None, :]
h = Z / 2 + 0.15 * Z * torch.sin(2 * 3.14159265 * 1.5 * yy / Y) + 0.10 * Z * torch.cos(2 * 3.14159265 * xx / X)
zz = torch.arange(Z, device=dev, dtype=torch.float32)[:, None, None]
score = 1000 * torch.exp(-(zz - h[None]) ** 2 / 8) + 30 * torch.rand((Z, Y, X), device=dev, generator=g)
score[Z // 2, Y // 2, X // 2] = 5000.0
nat.build_manifold_device(score)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    out = nat.build_manifold_device(score)
e1.record()
torch.cuda.synchronize()
print("build_manifold %dx%dx%d: %.2f ms  (height range %d..%d)" % (Z, Y, X, e0.elapsed_time(e1) / n, int(out.min()), int(out.max())))
