"""Development probe: the movie driver from a real TIFF file on disk to real output files (tiff_io reader / writer,
no I/O hook), timed end to end.
    python tools/tiff_movie_probe.py [T Z Y X] [--dry]
Writes a (T,1,Z,Y,X) uint16 movie (default 40 x 48 x 1024 x 1024 = BASELINE configs[2] frames, 3.8 GB) to a
temporary directory, runs movie_surface_projection on it twice (the first run warms the GPU path and the page cache)
and prints the seconds of the second run: frames are read from the file (page cache) straight into the pipeline's
pinned staging buffers by its host threads (preadv), projected, returned as uint16 and written as position1.tif +
zmap_position1.npy by several threads; a third run takes the frames as views of the file mapping instead (A/B).
--dry replaces the GPU call by a toy numpy operator (host-logic dry run without a GPU, small sizes)."""
import contextlib
import io
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tissue_image_processing_b200 import basic_image_manipulations as bim        # noqa: E402
from tissue_image_processing_b200 import surface_projection as sp                # noqa: E402
from tissue_image_processing_b200 import tiff_io                                 # noqa: E402
from tissue_image_processing_b200.movie import FramePipeline                     # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
dry = "--dry" in sys.argv
T, Z, Y, X = (int(v) for v in args[:4]) if len(args) >= 4 else (40, 48, 1024, 1024)

if dry:
    rng = np.random.default_rng(3)
    distinct = [rng.integers(1, 4000, (1, Z, Y, X), dtype=np.uint16) for _ in range(min(T, 4))]

    def toy(chunk, **kw):                                # stands in for the GPU call: the host path is what runs
        stack = chunk[0].astype(np.float64)
        return stack.max(axis=1), stack[0].argmax(axis=0)
    pipe = FramePipeline(operator=toy, out_dtype="uint16")
else:
    import torch
    import bench
    dev = torch.device("cuda", 0)
    distinct = [bench.synth_frame_device(torch, 70 + i, dev, (Z, Y, X)).cpu().numpy()[None] for i in range(min(T, 4))]
    pipe = FramePipeline(out_dtype="uint16")
movie = np.empty((T, 1, Z, Y, X), dtype=np.uint16)
for t in range(T):
    movie[t] = distinct[t % len(distinct)]
work = tempfile.mkdtemp(prefix="tsp_tiff_probe_")
try:
    path = os.path.join(work, "m1.tif")
    t0 = time.perf_counter()
    tiff_io.write_tiff(path, movie, "TCZYX")
    print("wrote %s: %.2f GB in %.2f s" % (path, os.path.getsize(path) / 1e9, time.perf_counter() - t0), flush=True)
    bim.open_image = tiff_io.open_tiff
    sp.tiff_writer = tiff_io.hook_writer
    for run in ("warm", "timed", "timed_mmap_views"):
        out = os.path.join(work, run)
        if run == "timed_mmap_views":            # A/B: frames as views of the file mapping, copied by the staging threads
            os.environ["TSP_TIFF_MMAP"] = "1"
        os.mkdir(out)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            sp.movie_surface_projection([path], 0, [1], 1, out, "max_averages", 1, False, 0, 0, 0, False,
                                        frame_pipeline=pipe)
        s = time.perf_counter() - t0
        print("%s run: %d frames of %dx%dx%d in %.3f s = %.0f frames/s  (%s), outputs %s" % (
            run, T, Z, Y, X, s, T / s, {k: round(v, 3) for k, v in sp.last_job_timings.items()},
            {f: os.path.getsize(os.path.join(out, f)) for f in sorted(os.listdir(out))}), flush=True)
    got = tiff_io.TiffImage(os.path.join(work, "timed", "position1.tif"))
    assert got.shape5 == (T, 1, 1, Y, X) and got.dtype == np.uint16, got.shape5
    first = got.get_image_dask_data()[0, 0, 0].compute()
    again = got.get_image_dask_data()[len(distinct), 0, 0].compute() if T > len(distinct) else first
    assert first.any() and np.array_equal(first, again), "frames t and t + %d hold the same stack" % len(distinct)
    if not dry:
        want = sp.time_point_surface_projection(movie[0:1], "TCZYX", 0, airyscan=False)
        assert np.array_equal(first, want[0].astype(np.uint16)), "file differs from the blocking operator call"
    print("output file checked", flush=True)
    # A/B of the output writes: threads per file, and the two files of a position one after the other / side by side
    from concurrent.futures import ThreadPoolExecutor
    proj = np.ascontiguousarray(got.get_image_data()[:, :, 0])
    zmap = np.load(os.path.join(work, "timed", "zmap_position1.npy"))
    tif, npy = os.path.join(work, "ab.tif"), os.path.join(work, "ab.npy")

    def both(th, side_by_side):
        for f in (tif, npy):
            if os.path.exists(f):
                os.remove(f)
        t0 = time.perf_counter()
        jobs = [lambda: tiff_io.write_tiff(tif, proj, "TCYX", threads=th), lambda: tiff_io.save_npy(npy, zmap, threads=th)]
        if side_by_side:
            with ThreadPoolExecutor(2) as pool:
                list(pool.map(lambda j: j(), jobs))
        else:
            for j in jobs:
                j()
        return time.perf_counter() - t0

    for side in (False, True):
        for th in (1, 2, 4, 8):
            ts = sorted(both(th, side) for _ in range(3))
            print("write %.0f MB tif + %.0f MB npy, %d thread(s) per file, %s: min %.3f s median %.3f s" % (
                proj.nbytes / 1e6, zmap.nbytes / 1e6, th, "side by side" if side else "one after the other", ts[0], ts[1]),
                flush=True)
finally:
    shutil.rmtree(work, ignore_errors=True)
