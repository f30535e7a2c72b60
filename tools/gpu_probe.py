"""Quick per-stage timing on the GPU box (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tissue_image_processing_b200 import _native as nat


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    Z, Y, X = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 2048, 2048)))
    modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["exact"]
    g = torch.Generator(device="cuda").manual_seed(0)
    # structured-ish synthetic on device: bright sheet + noise
    zz = torch.arange(Z, device="cuda", dtype=torch.float32)[:, None, None]
    yy = torch.arange(Y, device="cuda", dtype=torch.float32)[None, :, None]
    xx = torch.arange(X, device="cuda", dtype=torch.float32)[None, None, :]
    h = Z / 2 + 0.15 * Z * torch.sin(2 * np.pi * 1.5 * yy / Y) + 0.1 * Z * torch.cos(2 * np.pi * xx / X)
    vol = 300 + 2500 * torch.exp(-(zz - h) ** 2 / 8) + 40 * torch.randn((Z, Y, X), device="cuda", generator=g)
    stack = vol.clamp_(0, 65535).to(torch.uint16)[None].contiguous()
    del vol
    nbytes = stack.numel() * 2
    print("stack", tuple(stack.shape), nbytes / 2**20, "MiB")
    t = timeit(lambda: nat.percentile95_nonzero(stack[0]))
    print("percentile (hist+finalize+sync) ms", t, "GB/s", nbytes / t[0] / 1e6)
    for mode in modes:
        p = nat.DeviceProjector(1, Z, Y, X, mode=mode)
        t = timeit(lambda: p.run(stack), n=3, warm=1)
        print("frame", mode, "ms", t, "Gvox/s", stack.numel() / t[0] / 1e6, "status", p.status())
    zmap = p.zmap.clone()
    t = timeit(lambda: nat.band_project(stack, zmap))
    print("band_project (+range+sync) ms", t)


if __name__ == "__main__":
    main()
