"""Development aid: device time of the general path (bin_size > 1 / build_manifold) at a realistic frame size.
python tools/general_probe.py [Z Y X]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tissue_image_processing_b200 import _native as nat


def timeit(fn, n=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def main():
    dev = torch.device("cuda", 0)
    f = bench.synth_frame_device(torch, 0, dev)
    Cn, Z, Y, X = f.shape
    nat.handle(0)
    nat.load_library().tsp_set_profiling(nat.handle(0), 1)
    for kw in (dict(mode="exact"), dict(mode="exact", bin_size=2, method="max_averages"),
               dict(mode="exact", bin_size=4, method="max_std"), dict(mode="exact", bin_size=2, build_manifold=True),
               dict(mode="exact", bin_size=4, build_manifold=True), dict(mode="exact", build_manifold=True)):
        p = nat.DeviceProjector(Cn, Z, Y, X, **kw)
        ms = timeit(lambda: p.run(f), n=2)
        print(kw, "ms/frame %.2f" % ms, p.status())
        del p
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
