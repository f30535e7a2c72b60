"""Development probe: the same stack projected several times through one DeviceProjector (first call plain launches,
second captured, later ones graph replays) must give the same frame every time, whatever the workspace held before.
    python tools/repeat_probe.py [ZxYxX ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tissue_image_processing_b200 import _native as nat        # noqa: E402

nat.handle(0)
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(30, 1024, 1024)]
bad = 0
for (Z, Y, X) in shapes:
    g = torch.Generator(device="cuda").manual_seed(1)
    zz = torch.arange(Z, device="cuda", dtype=torch.float32)[:, None, None]
    vol = (torch.rand((Z, Y, X), device="cuda", generator=g) * 3000
           + 1000 * torch.exp(-(zz - Z * 0.4) ** 2 / 8)).to(torch.int32).to(torch.uint16)[None].contiguous()
    for mode in ("fast", "exact"):
        p = nat.DeviceProjector(1, Z, Y, X, airyscan=False, mode=mode, device=0)
        outs = []
        for fill in (0, 255, None, None, None):
            if fill is not None:
                p.workspace.fill_(fill)
            dp, dz = p.run(vol)
            torch.cuda.synchronize()
            outs.append((int(dz.sum().item()), float(dp.double().sum().item())))
        ok = all(o == outs[0] for o in outs)
        bad += not ok
        print("%dx%dx%d %s:" % (Z, Y, X, mode), "consistent" if ok else "DIFFERS BETWEEN CALLS %s" % outs, flush=True)
print("inconsistent cases:", bad)
sys.exit(1 if bad else 0)
