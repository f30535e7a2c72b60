"""Development aid: statistics of the fast-mode height map of the bench frame that drive the band-stage design."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from tissue_image_processing_b200 import _native as nat

dev = torch.device("cuda", 0)
f = bench.synth_frame_device(torch, 0, dev)
p = nat.DeviceProjector(1, bench.Z, bench.Y, bench.X, mode="fast")
p.run(f)
torch.cuda.synchronize()
z = p.zmap.cpu().numpy().astype(np.int32)
np.save("gpurun_out/zmap_bench.npy", z.astype(np.uint8))
Y, X = z.shape
zp = np.pad(z, 8, mode="edge")


def tile_stats(ty, tx, name):
    P, R = [], []
    for y0 in range(0, Y, ty):
        for x0 in range(0, X, tx):
            t = zp[y0:y0 + ty + 16, x0:x0 + tx + 16]
            P.append(len(np.unique(t)))
            R.append(int(t.max() - t.min()))
    P, R = np.array(P), np.array(R)
    print("%s: present mean %.2f p50 %d p90 %d max %d | range mean %.2f max %d | walked planes mean %.2f" % (
        name, P.mean(), np.median(P), np.percentile(P, 90), P.max(), R.mean(), R.max(), (R + 9).mean()))


for ty, tx in [(32, 64), (4, 64), (16, 16), (8, 32), (16, 32), (8, 64), (32, 32)]:
    tile_stats(ty, tx, "tile %dx%d" % (ty, tx))
# fraction of pixels whose 17x17 neighbourhood is uniform
from scipy.ndimage import maximum_filter, minimum_filter
mx = maximum_filter(z, size=17, mode="nearest")
mn = minimum_filter(z, size=17, mode="nearest")
print("uniform 17x17 neighbourhood: %.3f ; <=2 distinct range: %.3f ; range<=2: %.3f" % (
    (mx == mn).mean(), (mx - mn <= 1).mean(), (mx - mn <= 2).mean()))
print("zmap min/max", z.min(), z.max(), "steps between x-neighbours !=0: %.4f" % (np.diff(z, axis=1) != 0).mean())
