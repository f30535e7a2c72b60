#!/usr/bin/env python
"""Benchmark of the surface-projection hot path (contract in the task statement / SURVEY.md section 8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fast|exact|bitexact] [--impl reference]

A step = the whole operator (reference surface_projection.py:17-85) on one synthetic
2048x2048x64 uint16 single-channel frame (BASELINE.json configs[1]) per GPU; frames are independent, so
N GPUs = a movie partitioned by frame, no collective on the data path (scaling "weak").
  value         device-resident voxels/s (CUDA events around exactly K steps, max over ranks)
  e2e           the same through the public host-buffer API (movie.FramePipeline over the C ABI): pinned host frame
                in, float64 projection + int64 height map out, H2D and D2H inside the timed region
  roofline      dominant kernel: algorithmic bytes / CUDA-event duration vs the measured HBM copy peak
  configs       device-resident time, voxels/s and roofline fraction of ALL FIVE BASELINE configs
  modes         the exact (direct FIR, fp32) and bitexact (scipy's float64 order) score modes on the headline frame
  movie         BASELINE configs[2]: the FIXED 200-frame 1024x1024x48 movie through movie_surface_projection
                (strong scaling over the ranks, uint16 outputs converted on the device) next to the weak-scaled legs
  cpu_baseline  the CPU oracle (port of the reference path, same scipy calls) on a bounded crop, host cores
`--impl reference` times only that CPU path (the reference is pure numpy/scipy; /root/reference is not on
the GPU box, the oracle restates it bit for bit - tests/test_oracle_golden.py).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Z, Y, X = 64, 2048, 2048
WORKLOAD = "single 2048x2048x64 uint16 stack, one channel (BASELINE configs[1]); one frame per step per GPU"
METRIC = "projected voxels/s"
UNIT = "voxels/s"

# BASELINE.json configs: (key, C, Z, Y, X, note)
CONFIGS = [
    ("configs[0] 512x512x32", 1, 32, 512, 512, "the reference's own CPU-runnable case"),
    ("configs[1] 2048x2048x64", 1, 64, 2048, 2048, "headline"),
    ("configs[2] 1024x1024x48 movie frame", 1, 48, 1024, 1024, "one frame of the 200-frame time-lapse"),
    ("configs[3] 2ch 2048x2048x64", 2, 64, 2048, 2048, "second channel projected along the first channel's height map"),
    ("configs[4] 4096x4096x128", 1, 128, 4096, 4096, "whole 4 GiB frame in one call"),
    ("configs[4] tile 2048x2048x128", 1, 128, 2048, 2048, "one chunk_size=2048 tile of it (SP:294-301)"),
]


def algorithmic_bytes(C, z, y, x):
    """reference channel read 3x as uint16 (percentile, score, projection), other channels once,
    int32 height map + float32 projection written."""
    v = z * y * x
    return 2 * v * (C + 2) + y * x * (4 + 4 * C)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_digest():
    """sha256 over the CUDA sources of the fast path (+ the shared header and the ABI header): ncu captures under
    profiles/ are keyed by it, so a traffic figure is only quoted for the very kernels that were profiled."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "tissue_image_processing_b200", "csrc")
    for name in ("common.cuh", "percentile.cu", "fast.cu", "band.cu", "../../include/tsp_b200.h"):
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_traffic(stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per frame of the stage's kernels from the committed ncu --set full
    capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py) - or None when the kernels changed since."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, "no capture committed"
    with open(path) as f:
        cap = json.load(f)
    if cap.get("source_digest") != kernel_source_digest():
        return None, "capture %s is older than the kernels (source digest differs)" % cap.get("capture", "?")
    return cap.get("stage_bytes", {}).get(stage), "ncu --set full, %s (per frame, all kernels of the stage)" % cap.get("capture")


def synth_frame_device(torch, seed, device, shape=None, channels=1):
    """Same construction as oracle/synth.py (bright sheet on a smooth surface, sparse texture, noise),
    generated on the device because the numpy generator needs minutes at this size."""
    Zs, Ys, Xs = shape or (Z, Y, X)
    g = torch.Generator(device=device).manual_seed(1000 + seed)
    yy = torch.arange(Ys, device=device, dtype=torch.float32)[None, :, None]
    xx = torch.arange(Xs, device=device, dtype=torch.float32)[None, None, :]
    h = Zs / 2 + 0.15 * Zs * torch.sin(2 * np.pi * 1.5 * yy / Ys + 0.1 * seed) + 0.10 * Zs * torch.cos(2 * np.pi * xx / Xs)
    out = torch.empty((channels, Zs, Ys, Xs), dtype=torch.uint16, device=device)
    for c in range(channels):
        tex = 0.5 + 0.5 * (torch.rand((1, Ys, Xs), device=device, generator=g) < 0.15)
        for z0 in range(0, Zs, 16):                                   # plane blocks keep the temporaries small
            zz = torch.arange(z0, min(z0 + 16, Zs), device=device, dtype=torch.float32)[:, None, None]
            sig = 300.0 + (2500.0 - 1000.0 * c) * torch.exp(-(zz - h - c) ** 2 / 8.0) * tex
            sig += torch.sqrt(8.0 * sig) * torch.randn(sig.shape, device=device, generator=g)     # ~ 8*Poisson(sig/8)
            sig += 20.0 * torch.randn(sig.shape, device=device, generator=g)
            out[c, z0:z0 + 16] = sig.clamp_(0, 65535).round_().to(torch.uint16)
    return out                                                        # (C, Z, Y, X)


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.tmp.close()
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------------
# CPU arm
# --------------------------------------------------------------------------------------------------
def _cpu_crop(seed, shape):
    from oracle import synth
    return synth.synth_stack(shape[0], shape[1], shape[2], C=1, seed=seed)[None]


def _cpu_one(args):
    seed, shape = args
    from oracle import surface_projection_oracle as orc
    img = _cpu_crop(seed, shape)
    t0 = time.perf_counter()
    orc.time_point_surface_projection(img, "TCZYX", 0, airyscan=False, z_map=True)
    return time.perf_counter() - t0


def cpu_baseline(shape=(48, 640, 640), procs=1, rounds=1):
    """Oracle (bit-exact port of the reference operator) on crops of the workload frame.  Returns voxels/s
    over all processes (the reference path is single-threaded; parallelism = independent frames)."""
    vox = shape[0] * shape[1] * shape[2]
    jobs = [(100 + i, shape) for i in range(procs * rounds)]
    t0 = time.perf_counter()
    if procs == 1:
        for j in jobs:
            _cpu_one(j)
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            pool.map(_cpu_one, jobs)
    wall = time.perf_counter() - t0
    return vox * len(jobs) / wall, wall


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    procs = max(1, min(cores, 64))          # every host core the box gives us (bounded: each worker holds ~1 GB)
    # bounded sample: the whole run (W + K steps) stays within ~2.5 minutes at ~3 Mvoxel/s per core
    budget_s = min(6.0, 150.0 / max(1, args.warmup + args.steps))
    side = int(max(128, min(640, (budget_s * 3.0e6 / 48) ** 0.5 // 32 * 32)))
    shape = (48, side, side)
    per_step = []
    for i in range(args.warmup + args.steps):
        v, wall = cpu_baseline(shape, procs=procs)
        if i >= args.warmup:
            per_step.append((v, wall))
    value = float(np.mean([v for v, _ in per_step]))
    ms = float(np.mean([w for _, w in per_step]) * 1e3)
    sample = ("%d independent %dx%dx%d crops of the workload frame per step, one process each; voxels/s of the CPU "
              "path is independent of the crop shape to ~10 %%, so this stands for the 2048x2048x64 frame "
              "(extrapolated, not run at full size)" % (procs, shape[1], shape[2], shape[0]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 data / f64 accumulate (scipy)",
            "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class Timer:
    """Device-resident frame loops of one frame shape: one frame at a time on one stream (chained launches), and
    `streams` frames in flight on as many CUDA streams (plain launches)."""

    def __init__(self, torch, nat, device, index, C, Zs, Ys, Xs, mode, streams, frames, params=None):
        self.torch, self.device, self.frames, self.ns = torch, device, frames, streams
        kw = dict(reference_channel=0, airyscan=False, mode=mode, device=index, params=params)
        self.one = nat.DeviceProjector(C, Zs, Ys, Xs, **kw)
        self.many = [nat.DeviceProjector(C, Zs, Ys, Xs, concurrent=True, **kw) for _ in range(streams)] if streams > 1 else []
        self.strs = [torch.cuda.Stream(device=device) for _ in range(streams)] if streams > 1 else []

    def prime(self):
        """Every (projector, frame) pair twice: the library replays a frame's launches as a CUDA graph from the third
        call with the same buffers on (first call plain, second captured) - none of that inside a timed region."""
        self.torch.cuda.synchronize()          # the frames were generated on the current stream: side streams must see them
        for _ in range(2):
            for f in self.frames:
                self.one.run(f)
                for k, p in enumerate(self.many):
                    with self.torch.cuda.stream(self.strs[k]):
                        p.run(f)
        self.torch.cuda.synchronize()

    def serial(self, n):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            self.one.run(self.frames[i % len(self.frames)])
        e1.record()
        return e0, e1

    def inflight(self, n):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in self.strs:
            st.wait_event(e0)
        for i in range(n):
            with torch.cuda.stream(self.strs[i % self.ns]):
                self.many[i % self.ns].run(self.frames[i % len(self.frames)])
        for st in self.strs:
            ev = torch.cuda.Event()
            ev.record(st)
            torch.cuda.current_stream().wait_event(ev)
        e1.record()
        return e0, e1


def run_configs(args, torch, nat, device, index, peak, barrier, max_over_ranks):
    """Device-resident figures of all five BASELINE configs (fast mode): one frame at a time and `--streams` frames
    in flight; inputs of one shape cycle through a pool larger than the 126 MB L2."""
    out = {}
    for key, C, Zs, Ys, Xs, note in CONFIGS:
        frame_bytes = 2 * C * Zs * Ys * Xs
        pool = int(max(2, min(16, -(-300_000_000 // frame_bytes)))) if frame_bytes < (3 << 30) else 1
        frames = [synth_frame_device(torch, 200 + i, device, (Zs, Ys, Xs), C) for i in range(pool)]
        steps = max(4, min(args.steps, int(0.25e12 // (C * Zs * Ys * Xs * 64)) + 4))
        streams = args.movie_streams if frame_bytes < (200 << 20) else args.streams
        if frame_bytes < (40 << 20):
            streams = 2 * args.movie_streams         # a 16 MiB stack is launch-latency bound: more frames in flight
        if frame_bytes >= (3 << 30):
            streams = 2
        print("bench: %s (%d frames in the input pool, %d in flight)" % (key, pool, streams), file=sys.stderr, flush=True)
        t = Timer(torch, nat, device, index, C, Zs, Ys, Xs, "fast", streams, frames)
        t.prime()
        t.serial(3)
        t.inflight(2 * streams)
        barrier()
        e0, e1 = t.serial(steps)
        barrier()
        ms1 = max_over_ranks(e0.elapsed_time(e1)) / steps
        e0, e1 = t.inflight(steps * streams)
        barrier()
        msn = max_over_ranks(e0.elapsed_time(e1)) / (steps * streams)
        algo = algorithmic_bytes(C, Zs, Ys, Xs)
        vox = C * Zs * Ys * Xs
        out[key] = {"note": note, "channels": C, "shape_zyx": [Zs, Ys, Xs], "algorithmic_bytes": algo,
                    "ms_single_stream": ms1, "gvox_s_single_stream": vox / ms1 / 1e6,
                    "frac_single_stream": algo / (ms1 * 1e-3) / 1e9 / peak,
                    "frames_in_flight": streams, "ms_in_flight": msn, "gvox_s_in_flight": vox / msn / 1e6,
                    "frac_in_flight": algo / (msn * 1e-3) / 1e9 / peak,
                    "input_pool": "%d frames, %d MiB (> L2)" % (pool, pool * frame_bytes >> 20),
                    "near_tie_pixels": t.one.status()["near_tie_pixels"]}
        del t, frames
        torch.cuda.empty_cache()
    # configs[4] "wide z-band projection": the same tile with sigma_mask = (3, 2, 2) (tsp_params; band of +-12 planes
    # through the materialised mask path)
    key, C, Zs, Ys, Xs = "configs[4] tile 2048x2048x128, wide band sigma_mask=(3,2,2)", 1, 128, 2048, 2048
    frames = [synth_frame_device(torch, 300, device, (Zs, Ys, Xs), C)]
    t = Timer(torch, nat, device, index, C, Zs, Ys, Xs, "fast", 1, frames, params=dict(sigma_mask=(3.0, 2.0, 2.0)))
    t.prime()
    barrier()
    e0, e1 = t.serial(3)
    barrier()
    ms1 = max_over_ranks(e0.elapsed_time(e1)) / 3
    out[key] = {"note": "tsp_params.sigma_mask widened: band radius 12 planes, general (materialised) band stage",
                "ms_single_stream": ms1, "gvox_s_single_stream": C * Zs * Ys * Xs / ms1 / 1e6}
    del t, frames
    torch.cuda.empty_cache()
    return out


def run_modes(args, torch, nat, device, index, frames, barrier, max_over_ranks):
    """"report both, never only fast" - the exact and bit-exact score modes on the headline frame."""
    out = {}
    vox = Z * Y * X
    fp32_peak = 148 * 128 * 1.965e9            # FMA lanes x clock: the issue bound of the direct FIR (~510 MAC / voxel)
    for mode, steps in (("exact", 5), ("bitexact", 2)):
        t = Timer(torch, nat, device, index, 1, Z, Y, X, mode, 1, frames)
        t.prime()
        barrier()
        e0, e1 = t.serial(steps)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / steps
        out[mode] = {"ms_per_frame": ms, "gvox_s": vox / ms / 1e6,
                     "algorithmic_frac_of_hbm_peak": algorithmic_bytes(1, Z, Y, X) / (ms * 1e-3) / 1e9 / measured_peaks()[0]}
        if mode == "exact":
            out[mode]["fp32_issue_frac"] = 510.0 * vox / (ms * 1e-3) / fp32_peak
            out[mode]["note"] = "direct separable FIR, fp32 FMA; bound = 510 MAC/voxel at 148 x 128 lanes x 1.965 GHz = 3.7 ms"
        else:
            out[mode]["note"] = "fp64 accumulation in scipy's summation order: bit-identical to the reference"
        del t
        torch.cuda.empty_cache()
    return out


class _CyclicMovie:
    """In-memory stand-in for an aicsimageio image (the surface the drivers use): a T-frame movie whose time points
    cycle through a few distinct frames held in ordinary (pageable) host memory - what a file reader hands back."""

    class _Lazy:
        def __init__(self, owner, t0=None):
            self.owner, self.t0 = owner, t0

        def __getitem__(self, idx):
            t = idx[0] if isinstance(idx, tuple) else idx
            return _CyclicMovie._Lazy(self.owner, t.start)

        def compute(self):
            return self.owner.frames[self.t0 % len(self.owner.frames)][None]

    def __init__(self, frames, T):
        import types
        self.frames, self.T = frames, T
        C, Zs, Ys, Xs = frames[0].shape
        self.dims = types.SimpleNamespace(T=T, C=C, Z=Zs, Y=Ys, X=Xs)

    @property
    def metadata(self):                                    # a fresh object per read, like re-opening the file
        import types
        d = self.dims
        stage = types.SimpleNamespace(x=0.0, y=0.0, z=0.0, x_unit="um", y_unit="um", z_unit="um")
        pixels = types.SimpleNamespace(size_t=d.T, size_c=d.C, size_z=d.Z, physical_size_x=0.1, physical_size_y=0.1,
                                       physical_size_z=0.5, dimension_order="XYZCT", type="uint16", planes=list(range(d.C)))
        return types.SimpleNamespace(images=[types.SimpleNamespace(name="p", stage_label=stage, pixels=pixels)])

    def set_scene(self, i):
        pass

    def get_image_dask_data(self):
        return _CyclicMovie._Lazy(self)


MOVIE_SHAPE = (48, 1024, 1024)          # BASELINE configs[2]: 200-frame 1024x1024x48 time-lapse, frame-parallel
MOVIE_FRAMES = 200


def run_movie_leg(args, torch, nat, mv, device, index, rank, world, barrier, max_over_ranks):
    """BASELINE configs[2].  (1) the fixed 200-frame job end to end through movie_surface_projection (strong scaling:
    the ranks share the 200 time points, rank 0 assembles and writes); (2) device-resident frames/s with
    `--movie-streams` frames in flight, `--movie-frames` per GPU (weak); (3) the same count per GPU end to end from
    pinned host frames through FramePipeline (weak)."""
    if args.movie_frames <= 0:
        return None
    from tissue_image_processing_b200 import basic_image_manipulations as bim
    from tissue_image_processing_b200 import surface_projection as sp
    FramePipeline, SharedFrameCounter = mv.FramePipeline, mv.SharedFrameCounter
    Zm, Ym, Xm = MOVIE_SHAPE
    n = args.movie_frames
    dframes = [synth_frame_device(torch, 50 + 10 * rank + i, device, MOVIE_SHAPE) for i in range(4)]
    ns = max(1, args.movie_streams)
    t = Timer(torch, nat, device, index, 1, Zm, Ym, Xm, args.mode, ns, dframes)
    t.prime()
    t.inflight(2 * ns)
    barrier()
    e0, e1 = t.inflight(n)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    # --- weak, end to end from pinned frames
    hframes = []
    for f in dframes:
        hbuf = nat.pinned_empty((1, Zm, Ym, Xm), np.uint16)
        torch.from_numpy(hbuf).copy_(f)
        hframes.append(hbuf)
    torch.cuda.synchronize()
    pipe16 = FramePipeline(devices=[index], slots=args.slots, mode=args.mode, out_dtype="uint16")
    seen = [0]

    def sink(k, proj, zmap, status):
        seen[0] += int(zmap[0, 0] >= 0)

    def gen(counter, total):
        for i in counter.claims(total):
            yield i, hframes[i % 4]

    warm_counter, movie_counter = SharedFrameCounter("movie_warm"), SharedFrameCounter("movie_weak")
    pipe16.project_frames(gen(warm_counter, 4 * world), sink, reference_channel=0, airyscan=False)
    barrier()
    t0 = time.perf_counter()
    pipe16.project_frames(gen(movie_counter, n * world), sink, reference_channel=0, airyscan=False)
    barrier()
    weak_s = max_over_ranks(time.perf_counter() - t0)
    # --- the fixed 200-frame job through the driver (in-memory image source, TIFF writer hook = a no-op sink)
    pageable = [np.array(h) for h in hframes]                       # ordinary host memory, like a reader's output
    source = _CyclicMovie(pageable, MOVIE_FRAMES)
    old_open, old_writer = bim.open_image, sp.tiff_writer
    bim.open_image = lambda path: source
    written = []
    sp.tiff_writer = lambda path, image, axes, metadata: written.append((path, image.shape, str(image.dtype)))
    out_dir = tempfile.mkdtemp(prefix="tsp_bench_") if rank == 0 else tempfile.gettempdir()
    if world > 1:
        import torch.distributed as dist
        box = [out_dir]
        dist.broadcast_object_list(box, src=0, group=mv.host_group())
        out_dir = box[0]
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            sp.movie_surface_projection(["bench_movie.czi"], 0, [1], 1, out_dir, "max_averages", 1, False, 0, 0, 0, False,
                                        output_name="warm_", mode=args.mode, frame_pipeline=pipe16)      # warm-up job
        barrier()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            sp.movie_surface_projection(["bench_movie.czi"], 0, [1], 1, out_dir, "max_averages", 1, False, 0, 0, 0, False,
                                        mode=args.mode, frame_pipeline=pipe16)
        barrier()
        job_s = max_over_ranks(time.perf_counter() - t0)
        job_phases = {k: round(v, 4) for k, v in sp.last_job_timings.items()}
        # the projection part alone (project_movie: read, stage, project, gather on rank 0), without the driver's
        # np.save / np.load / concatenate of the resume files
        outs = mv.allocate_outputs([((MOVIE_FRAMES, 1, 1, Ym, Xm), np.uint16), ((MOVIE_FRAMES, 1, 1, Ym, Xm), np.uint16)])
        proj, zmap = outs.arrays
        barrier()
        t0 = time.perf_counter()
        pipe16.project_movie("bench_movie.czi", 0, proj, zmap, gather="root", reference_channel=0, airyscan=False,
                             atoh_shift=0, min_z=0, max_z=0)
        barrier()
        proj_s = max_over_ranks(time.perf_counter() - t0)
        shared_outputs = outs.shared
        del proj, zmap
        outs.close()
    finally:
        bim.open_image, sp.tiff_writer = old_open, old_writer
        if rank == 0:
            import shutil
            shutil.rmtree(out_dir, ignore_errors=True)
    return {"workload": "1024x1024x48 uint16 time-lapse frames (BASELINE configs[2])",
            "fixed_job": {"frames": MOVIE_FRAMES, "scaling": "strong",
                          "api": "movie_surface_projection(files, ..., frame_pipeline=FramePipeline(out_dtype='uint16')) on an "
                                 "in-memory image source whose frames are ordinary (pageable) arrays: staged into pinned "
                                 "buffers by host threads, projected, uint16 on the device, scattered by every rank into "
                                 "the job's output arrays (one shared-memory mapping on a single host; gloo assembly "
                                 "otherwise), resume .npy files + TIFF hook + zmap written by rank 0",
                          "shared_output_arrays": bool(shared_outputs),
                          "rank0_phase_seconds": job_phases,
                          "seconds": job_s, "frames_per_s": MOVIE_FRAMES / job_s,
                          "projection_seconds": proj_s, "projection_frames_per_s": MOVIE_FRAMES / proj_s,
                          "outputs": written[-1:] if rank == 0 else None},
            "weak": {"frames_per_gpu": n, "frames": n * world, "frames_in_flight": ns,
                     "frames_per_s": n * world / (dev_ms * 1e-3), "ms_per_frame": dev_ms / n,
                     "voxels_per_s": n * world * Zm * Ym * Xm / (dev_ms * 1e-3),
                     "frames_per_s_e2e": n * world / weak_s, "ms_per_frame_e2e": weak_s * 1e3 / n,
                     "note": "frames_per_s: device-resident; frames_per_s_e2e: pinned host frame in, uint16 projection + "
                             "uint16 height map out (converted on the device)"},
            # round-1 keys, kept for continuity
            "frames": n * world, "frames_per_s": n * world / (dev_ms * 1e-3), "frames_per_s_e2e": n * world / weak_s,
            "ms_per_frame": dev_ms / n, "ms_per_frame_e2e": weak_s * 1e3 / n, "frames_in_flight": ns}


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU arm")
    from tissue_image_processing_b200 import topology
    index, topo = topology.choose_device(local_rank, world)
    torch.cuda.set_device(index)
    device = torch.device("cuda", index)
    if world > 1:
        # stdout carries exactly one JSON line: whatever NCCL has to say (its version banner) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    from tissue_image_processing_b200 import _native as nat
    from tissue_image_processing_b200 import movie as mv
    import tissue_image_processing_b200 as tsp
    host_cpus = nat.bind_host_thread_to_gpu(index)      # pinned frames on the GPU's own NUMA node
    mv.host_group()                                     # the gloo group for output assembly (collective: all ranks)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nframes = 2
    frames = [synth_frame_device(torch, 10 * rank + i, device) for i in range(nframes)]
    vox = Z * Y * X
    timer = Timer(torch, nat, device, index, 1, Z, Y, X, args.mode, args.streams, frames)

    # ---- device-resident ---------------------------------------------------------------------------
    timer.prime()
    timer.serial(args.warmup)
    barrier()
    sampler = ClockSampler(index) if rank == 0 else None
    launches0 = nat.launch_count(index)
    ev0, ev1 = timer.serial(args.steps)
    barrier()
    serial_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = nat.launch_count(index) - launches0
    status = timer.one.status()
    # the same K steps with `--streams` independent frames in flight (one DeviceProjector and CUDA stream each, the way
    # movie.FramePipeline's frame slots run): the issue-bound stages of one frame overlap the HBM-bound ones of another
    dev_ms, in_flight = serial_ms, 1
    if args.streams > 1:
        timer.inflight(max(args.warmup, args.streams))
        barrier()
        launches0 = nat.launch_count(index)
        e0, e1 = timer.inflight(args.steps)
        barrier()
        dev_ms, in_flight = max_over_ranks(e0.elapsed_time(e1)), args.streams
        launches = nat.launch_count(index) - launches0
    # per-stage times: a second, untimed-for-the-headline pass of the same K steps with the library's own CUDA
    # events between the stages (they sit on the launching stream, so they are kept out of the timed region)
    nat.set_profiling(True, index)
    nat.stage_times(reset=True, device=index)
    timer.serial(args.steps)
    barrier()
    stages = nat.stage_times(reset=True, device=index)
    nat.set_profiling(False, index)

    if args.device_only:       # development aid: kernels only, one short line
        if sampler:
            sampler.stop()
        if rank == 0:
            sm = {k: round(v[0] / max(v[1], 1), 5) for k, v in stages.items() if v[1]}
            emit({"ms_per_step": dev_ms / args.steps, "frames_in_flight": in_flight,
                  "single_stream_ms_per_step": serial_ms / args.steps, "stage_ms": sm, "status": status})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public API ----------------------------------------------------------
    print("bench: end to end", file=sys.stderr, flush=True)
    host_frames = []
    for f in frames:
        h = nat.pinned_empty((1, 1, Z, Y, X), np.uint16)
        torch.from_numpy(h).copy_(f.view(1, 1, Z, Y, X))
        host_frames.append(h)
    torch.cuda.synchronize()
    kw = dict(reference_channel=0, airyscan=False, z_map=True, mode=args.mode, device=index)
    for i in range(min(args.warmup, 3)):
        tsp.time_point_surface_projection(host_frames[i % nframes], "TCZYX", **kw)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        p64, z64 = tsp.time_point_surface_projection(host_frames[i % nframes], "TCZYX", **kw)
    barrier()
    single_s = max_over_ranks(time.perf_counter() - t0)
    # the same call on an ordinary (pageable) numpy array - what a reader hands to the reference's operator: the
    # library stages it through pinned chunks with a pool of host threads (api.cu: staged_copy_in)
    pageable_frame = np.array(host_frames[0])
    for i in range(2):
        tsp.time_point_surface_projection(pageable_frame, "TCZYX", **kw)
    barrier()
    t0 = time.perf_counter()
    nsingle = max(3, args.steps // 4)
    for i in range(nsingle):
        p64, z64 = tsp.time_point_surface_projection(pageable_frame, "TCZYX", **kw)
    barrier()
    single_pageable_s = max_over_ranks(time.perf_counter() - t0) / nsingle
    del pageable_frame
    # the movie API: same frames through the slot pipeline (copy-in of frame t+1 overlaps the kernels of frame t)
    pipe = mv.FramePipeline(devices=[index], slots=args.slots, mode=args.mode)
    checksum = [0.0]

    def sink(t, proj, zmap, status):
        checksum[0] += float(proj[0, 0, 0]) + float(zmap[0, 0])        # the result is read on the host

    # the job's world * K frames are claimed from a counter shared by the ranks (movie.SharedFrameCounter): the host
    # links of a multi-GPU box are not equally fast, a rank on a faster link takes more frames; one rank = plain count
    def claimed(counter, total):
        for i in counter.claims(total):
            yield i, host_frames[i % nframes][0]

    warm_counter, e2e_counter = mv.SharedFrameCounter("e2e_warm"), mv.SharedFrameCounter("e2e")
    pipe.project_frames(claimed(warm_counter, world * min(args.warmup, 3)), sink, reference_channel=0, airyscan=False)
    barrier()
    bytes0 = pipe.h2d_bytes
    t0 = time.perf_counter()
    pipe.project_frames(claimed(e2e_counter, world * args.steps), sink, reference_channel=0, airyscan=False)
    my_s = time.perf_counter() - t0
    my_gbs = (pipe.h2d_bytes - bytes0) / my_s / 1e9          # this rank's host-link rate while all ranks copy
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    link_rates = None
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, round(my_gbs, 1), group=mv.host_group())
        link_rates = allr
    peak, peak_src = measured_peaks()
    configs = run_configs(args, torch, nat, device, index, peak, barrier, max_over_ranks) if not args.skip_configs else None
    modes = run_modes(args, torch, nat, device, index, frames, barrier, max_over_ranks) if not args.skip_configs else None
    movie = run_movie_leg(args, torch, nat, mv, device, index, rank, world, barrier, max_over_ranks)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * args.steps * vox / (dev_ms * 1e-3)
    frame_bytes = algorithmic_bytes(1, Z, Y, X)
    frame_gbs = frame_bytes * args.steps / (dev_ms * 1e-3) / 1e9
    # dominant kernel = the stage with the largest share of device time
    stage_ms = {k: v[0] / max(v[1], 1) for k, v in stages.items()}
    dom = max(stage_ms, key=stage_ms.get)
    # algorithmic bytes of each stage of the fast path (per frame): every score-path stage is charged the one
    # uint16 read of the reference channel it exists for; band = one uint16 read + both outputs
    stage_bytes = {"percentile": 2 * vox, "percentile_count": 2 * vox, "percentile_sample": 2 * vox // 32,
                   "decimate": 2 * vox, "blur_score": 2 * vox, "blur_pre": 2 * vox,
                   "prepare": 2 * vox, "band": 2 * vox + Y * X * 8, "interp_argmax": 2 * vox, "coarse": 2 * vox,
                   "argmax": 2 * vox}
    stage_kernel = {"decimate": "decimate_ring_kernel", "percentile_count": "window_count_kernel",
                    "percentile_sample": "sample_window_kernel", "band": "band_project4_kernel",
                    "interp_argmax": "interp_argmax_kernel", "coarse": "coarse_xy_kernel + coarse_zmix_kernel"}
    achieved = stage_bytes.get(dom, 2 * vox) / (stage_ms[dom] * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(dom) if args.mode == "fast" else (None, "captured for the fast mode only")
    # the CPU leg runs on rank 0 at N=1 only (at N>1 it would only add minutes next to seven idle ranks)
    cpu_v, cpu_wall = cpu_baseline((48, 1024, 1024), procs=1) if world == 1 else (None, 0.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (u16 in)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "mode": args.mode, "frames_per_step_per_gpu": 1,
                   "frames_in_flight": in_flight,
                   "concurrency": "K independent frames, %d in flight on %d CUDA streams per GPU (single_stream = "
                                  "the same K frames one after the other)" % (in_flight, in_flight),
                   "l2": "inputs (512 MiB per frame, 2 alternating) larger than the 126 MB L2",
                   "algorithmic_bytes_per_frame": frame_bytes,
                   "deterministic": "fixed-point coarse accumulation: bit-identical results run to run"},
        "e2e": {"value": world * args.steps * vox / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": int(vox * 2), "d2h_bytes_per_step": int(Y * X * 16 + 256),
                "ms_per_step": e2e_s * 1e3 / args.steps,
                "api": "movie.FramePipeline.project_frames(pinned host uint16 frames) -> float64 projection, int64 "
                       "height map per frame on the host (%d frame slots: copy-in overlaps kernels)" % args.slots,
                "frame_assignment": "the job's N*K frames are claimed by the ranks from a shared counter (the host "
                                    "links of a multi-GPU box differ in speed); bytes are per frame",
                "h2d_gbs_rank0": round(my_gbs, 1), "h2d_gbs_per_rank": link_rates,
                "h2d_gbs_total": round(world * args.steps * vox * 2 / e2e_s / 1e9, 1),
                "host_link_probe_gbs": topo.get("probe_gbs"),
                "single_call_ms": single_s * 1e3 / args.steps,
                "single_call_api": "time_point_surface_projection(frame, 'TCZYX', ...) one blocking call per frame "
                                   "(frame in pinned memory); single_call_pageable_ms: the same call on an ordinary "
                                   "numpy array",
                "single_call_pageable_ms": single_pageable_s * 1e3},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": stage_kernel.get(dom, dom), "stage": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel_ms": stage_ms[dom]},
        "frame_roofline": {"achieved": frame_gbs, "peak": peak, "unit": "GB/s", "frac": frame_gbs / peak,
                           "frac_of_nominal_8TBs": frame_gbs / 8000.0,
                           "note": "whole operator, SURVEY 8(d) algorithmic bytes / device time"},
        "single_stream": {"ms_per_step": serial_ms / args.steps,
                          "value": world * args.steps * vox / (serial_ms * 1e-3), "unit": UNIT,
                          "frac_of_measured_peak": frame_bytes * args.steps / (serial_ms * 1e-3) / 1e9 / peak,
                          "note": "one frame at a time on one stream = the latency of a single stack"},
        "stage_ms": stage_ms,
        "configs": configs,
        "modes": modes,
        "movie": movie,
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": ("one 1024x1024x48 crop of the workload frame, %.1f s (extrapolated to the full "
                                    "frame: the CPU path's voxels/s does not depend on the shape)" % cpu_wall)
                         if world == 1 else "not measured at N>1 (see the N=1 line)"},
        "clocks": clocks,
        "device_mapping": topo,
        "host_cpus": ("%d CPUs next to the GPU (NVML affinity)" % len(host_cpus)) if host_cpus else "not bound",
        "frame_status": status,
        "kernel_source_digest": kernel_source_digest(),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else any library prints to fd 1 (NCCL's version
    banner, for instance) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default="fast", choices=["fast", "exact", "bitexact"])
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--device-only", action="store_true", help="development aid: skip the end-to-end and CPU legs")
    ap.add_argument("--skip-configs", action="store_true", help="development aid: skip the five-config and mode tables")
    ap.add_argument("--movie-frames", type=int, default=48,
                    help="frames per GPU of the weak-scaled 1024x1024x48 movie legs (BASELINE configs[2]); 0 skips the movie")
    ap.add_argument("--slots", type=int, default=2, help="frame slots of the end-to-end pipeline (frames a rank holds)")
    ap.add_argument("--movie-streams", type=int, default=4, help="frames in flight per GPU in the movie leg")
    ap.add_argument("--streams", type=int, default=3, help="independent frames in flight per GPU (CUDA streams)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
