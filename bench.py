#!/usr/bin/env python
"""Benchmark of the surface-projection hot path (contract in the task statement / SURVEY.md section 8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fast|exact|bitexact] [--impl reference]

A step = the whole operator (reference surface_projection.py:17-85) on one synthetic
2048x2048x64 uint16 single-channel frame (BASELINE.json configs[1]) per GPU; frames are independent, so
N GPUs = a movie partitioned by frame, no collective on the data path (scaling "weak").
  value  device-resident voxels/s (CUDA events around exactly K steps, max over ranks)
  e2e    the same through the public host-buffer API time_point_surface_projection(): pinned host frame in,
         float64 projection + int64 height map out, H2D and D2H inside the timed region
  roofline      dominant kernel: algorithmic bytes / CUDA-event duration vs the measured HBM copy peak
  cpu_baseline  the CPU oracle (port of the reference path, same scipy calls) on a bounded crop, host cores
`--impl reference` times only that CPU path (the reference is pure numpy/scipy; /root/reference is not on
the GPU box, the oracle restates it bit for bit - tests/test_oracle_golden.py).
"""
import argparse
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


def synth_frame_device(torch, seed, device, shape=None):
    """Same construction as oracle/synth.py (bright sheet on a smooth surface, sparse texture, noise),
    generated on the device because the numpy generator needs minutes at this size."""
    Zs, Ys, Xs = shape or (Z, Y, X)
    g = torch.Generator(device=device).manual_seed(1000 + seed)
    zz = torch.arange(Zs, device=device, dtype=torch.float32)[:, None, None]
    yy = torch.arange(Ys, device=device, dtype=torch.float32)[None, :, None]
    xx = torch.arange(Xs, device=device, dtype=torch.float32)[None, None, :]
    h = Zs / 2 + 0.15 * Zs * torch.sin(2 * np.pi * 1.5 * yy / Ys + 0.1 * seed) + 0.10 * Zs * torch.cos(2 * np.pi * xx / Xs)
    tex = 0.5 + 0.5 * (torch.rand((1, Ys, Xs), device=device, generator=g) < 0.15)
    sig = 300.0 + 2500.0 * torch.exp(-(zz - h) ** 2 / 8.0) * tex
    sig += torch.sqrt(8.0 * sig) * torch.randn(sig.shape, device=device, generator=g)      # ~ 8*Poisson(sig/8)
    sig += 20.0 * torch.randn(sig.shape, device=device, generator=g)
    return sig.clamp_(0, 65535).round_().to(torch.uint16)[None].contiguous()               # (1, Z, Y, X)


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
    sample = "%d independent %dx%dx%d crops of the workload frame per step, one process each" % (
        procs, shape[1], shape[2], shape[0])
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
MOVIE_SHAPE = (48, 1024, 1024)          # BASELINE configs[2]: 200-frame 1024x1024x48 time-lapse, frame-parallel


def run_movie_leg(args, torch, nat, pipe, device, local_rank, rank, barrier, max_over_ranks, world):
    """BASELINE configs[2] next to the headline: `--movie-frames` frames of 1024x1024x48 per GPU (frames are dealt
    round-robin to the ranks, so N GPUs project N times as many in the same time), device-resident with 3 frames
    in flight and end to end through movie.FramePipeline from pinned host frames."""
    n = args.movie_frames
    if n <= 0:
        return None
    Zm, Ym, Xm = MOVIE_SHAPE
    dframes = [synth_frame_device(torch, 50 + 10 * rank + i, device, MOVIE_SHAPE) for i in range(4)]
    ns = max(1, args.movie_streams)
    projs = [nat.DeviceProjector(1, Zm, Ym, Xm, reference_channel=0, airyscan=False, mode=args.mode,
                                 device=local_rank, concurrent=ns > 1) for _ in range(ns)]
    strs = [torch.cuda.Stream(device=device) for _ in range(ns)]

    def resident(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in strs:
            st.wait_event(e0)
        for i in range(k):
            with torch.cuda.stream(strs[i % ns]):
                projs[i % ns].run(dframes[i % 4])
        for st in strs:
            ev = torch.cuda.Event()
            ev.record(st)
            torch.cuda.current_stream().wait_event(ev)
        e1.record()
        return e0, e1

    resident(2 * ns)
    barrier()
    e0, e1 = resident(n)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    hframes = []
    for f in dframes:
        hbuf = nat.pinned_empty((1, Zm, Ym, Xm), np.uint16)
        torch.from_numpy(hbuf).copy_(f)
        hframes.append(hbuf)
    torch.cuda.synchronize()
    seen = [0]

    def sink(t, proj, zmap, status):
        seen[0] += int(zmap[0, 0] >= 0)

    from tissue_image_processing_b200.movie import SharedFrameCounter

    def gen(counter, total):
        for i in counter.claims(total):
            yield i, hframes[i % 4]

    warm_counter, movie_counter = SharedFrameCounter("movie_warm"), SharedFrameCounter("movie")
    pipe.project_frames(gen(warm_counter, 4 * world), sink, reference_channel=0, airyscan=False)
    barrier()
    t0 = time.perf_counter()
    pipe.project_frames(gen(movie_counter, n * world), sink, reference_channel=0, airyscan=False)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    return {"workload": "1024x1024x48 uint16 time-lapse frames (BASELINE configs[2]), %d frames per GPU, frames "
                        "partitioned over the ranks, no collective" % n,
            "frames": n * world, "frames_per_s": n * world / (dev_ms * 1e-3),
            "frames_per_s_e2e": n * world / e2e_s, "ms_per_frame": dev_ms / n, "ms_per_frame_e2e": e2e_s * 1e3 / n,
            "voxels_per_s": n * world * Zm * Ym * Xm / (dev_ms * 1e-3),
            "frames_in_flight": ns,
            "note": "frames_per_s: device-resident, frames_in_flight frames in flight per GPU; frames_per_s_e2e: pinned host frame in, "
                    "float64 projection + int64 height map out (PCIe-bound)"}


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: whatever NCCL has to say (its version banner) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    import tissue_image_processing_b200 as tsp
    from tissue_image_processing_b200 import _native as nat
    host_cpus = nat.bind_host_thread_to_gpu(local_rank)      # pinned frames on the GPU's own NUMA node

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

    nframes = 2
    frames = [synth_frame_device(torch, 10 * rank + i, device) for i in range(nframes)]
    vox = Z * Y * X
    proj = nat.DeviceProjector(1, Z, Y, X, reference_channel=0, airyscan=False, mode=args.mode, device=local_rank)

    # ---- device-resident ---------------------------------------------------------------------------
    for i in range(args.warmup):
        proj.run(frames[i % nframes])
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = nat.launch_count(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        proj.run(frames[i % nframes])
    ev1.record()
    barrier()
    serial_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = nat.launch_count(local_rank) - launches0
    status = proj.status()
    # the same K steps with `--streams` independent frames in flight (one DeviceProjector and CUDA stream each, the way
    # movie.FramePipeline's frame slots run): the issue-bound stages of one frame overlap the HBM-bound ones of another
    dev_ms, in_flight = serial_ms, 1
    if args.streams > 1:
        projs = [nat.DeviceProjector(1, Z, Y, X, reference_channel=0, airyscan=False, mode=args.mode,
                                     device=local_rank, concurrent=True) for _ in range(args.streams)]
        strs = [torch.cuda.Stream(device=device) for _ in range(args.streams)]

        def pipelined(nsteps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for st in strs:
                st.wait_event(e0)
            for i in range(nsteps):
                with torch.cuda.stream(strs[i % args.streams]):
                    projs[i % args.streams].run(frames[i % nframes])
            for st in strs:
                ev = torch.cuda.Event()
                ev.record(st)
                torch.cuda.current_stream().wait_event(ev)
            e1.record()
            return e0, e1

        pipelined(max(args.warmup, args.streams))
        barrier()
        launches0 = nat.launch_count(local_rank)
        e0, e1 = pipelined(args.steps)
        barrier()
        dev_ms, in_flight = max_over_ranks(e0.elapsed_time(e1)), args.streams
        launches = nat.launch_count(local_rank) - launches0
    # per-stage times: a second, untimed-for-the-headline pass of the same K steps with the library's own CUDA
    # events between the stages (they sit on the launching stream, so they are kept out of the timed region)
    nat.set_profiling(True, local_rank)
    nat.stage_times(reset=True, device=local_rank)
    for i in range(args.steps):
        proj.run(frames[i % nframes])
    barrier()
    stages = nat.stage_times(reset=True, device=local_rank)
    nat.set_profiling(False, local_rank)

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
    host_frames = []
    for f in frames:
        h = nat.pinned_empty((1, 1, Z, Y, X), np.uint16)
        torch.from_numpy(h).copy_(f.view(1, 1, Z, Y, X))
        host_frames.append(h)
    torch.cuda.synchronize()
    kw = dict(reference_channel=0, airyscan=False, z_map=True, mode=args.mode, device=local_rank)
    for i in range(min(args.warmup, 3)):
        tsp.time_point_surface_projection(host_frames[i % nframes], "TCZYX", **kw)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        p64, z64 = tsp.time_point_surface_projection(host_frames[i % nframes], "TCZYX", **kw)
    barrier()
    single_s = max_over_ranks(time.perf_counter() - t0)
    # the movie API: same frames through the slot pipeline (copy-in of frame t+1 overlaps the kernels of frame t)
    from tissue_image_processing_b200.movie import FramePipeline, SharedFrameCounter
    pipe = FramePipeline(devices=[local_rank], slots=args.slots, mode=args.mode)
    checksum = [0.0]

    def sink(t, proj, zmap, status):
        checksum[0] += float(proj[0, 0, 0]) + float(zmap[0, 0])        # the result is read on the host

    # the job's world * K frames are claimed from a counter shared by the ranks (movie.SharedFrameCounter): the host
    # links of a multi-GPU box are not equally fast, a rank on a faster link takes more frames; one rank = plain count
    def frames(counter, total):
        for i in counter.claims(total):
            yield i, host_frames[i % nframes][0]

    warm_counter, e2e_counter = SharedFrameCounter("e2e_warm"), SharedFrameCounter("e2e")
    pipe.project_frames(frames(warm_counter, world * min(args.warmup, 3)), sink, reference_channel=0, airyscan=False)
    barrier()
    t0 = time.perf_counter()
    pipe.project_frames(frames(e2e_counter, world * args.steps), sink, reference_channel=0, airyscan=False)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    movie = run_movie_leg(args, torch, nat, pipe, device, local_rank, rank, barrier, max_over_ranks, world)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
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
    # dram__bytes_read.sum + dram__bytes_write.sum per frame of each stage's kernels, from the ncu --set full capture
    # of this very workload (profiles/r1f_ncu_full_summary.txt); only meaningful for the default fast mode
    ncu_traffic = {"percentile_sample": 16.83e6, "percentile_count": 536.9e6 + 6.41e6, "percentile": 0.05e6,
                   "decimate": 555.8e6 + 19.64e6, "coarse": 2 * 17.45e6, "interp_argmax": 5.57e6,
                   "band": 111.2e6 + 7.89e6 + 0.07e6}
    traffic = ncu_traffic.get(dom) if args.mode == "fast" else None
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
                   "algorithmic_bytes_per_frame": frame_bytes},
        "e2e": {"value": world * args.steps * vox / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": int(vox * 2), "d2h_bytes_per_step": int(Y * X * 16 + 256),
                "ms_per_step": e2e_s * 1e3 / args.steps,
                "api": "movie.FramePipeline.project_frames(pinned host uint16 frames) -> float64 projection, int64 "
                       "height map per frame on the host (%d frame slots: copy-in overlaps kernels)" % args.slots,
                "frame_assignment": "the job's N*K frames are claimed by the ranks from a shared counter (the host "
                                    "links of a multi-GPU box differ in speed); bytes are per frame",
                "single_call_ms": single_s * 1e3 / args.steps,
                "single_call_api": "time_point_surface_projection(frame, 'TCZYX', ...) one blocking call per frame"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": stage_kernel.get(dom, dom), "stage": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": "ncu --set full, profiles/r1f_ncu_full_summary.txt (per frame, all kernels of "
                                       "the stage)", "peak_source": peak_src,
                     "kernel_ms": stage_ms[dom]},
        "frame_roofline": {"achieved": frame_gbs, "peak": peak, "unit": "GB/s", "frac": frame_gbs / peak,
                           "frac_of_nominal_8TBs": frame_gbs / 8000.0,
                           "note": "whole operator, SURVEY 8(d) algorithmic bytes / device time"},
        "single_stream": {"ms_per_step": serial_ms / args.steps,
                          "value": world * args.steps * vox / (serial_ms * 1e-3), "unit": UNIT,
                          "frac_of_measured_peak": frame_bytes * args.steps / (serial_ms * 1e-3) / 1e9 / peak,
                          "note": "one frame at a time on one stream = the latency of a single stack"},
        "stage_ms": stage_ms,
        "movie": movie,
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": ("one 1024x1024x48 crop of the workload frame, %.1f s" % cpu_wall) if world == 1
                         else "not measured at N>1 (see the N=1 line)"},
        "clocks": clocks,
        "host_cpus": ("%d CPUs next to the GPU (NVML affinity)" % len(host_cpus)) if host_cpus else "not bound",
        "frame_status": status,
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
    ap.add_argument("--movie-frames", type=int, default=48,
                    help="frames per GPU of the 1024x1024x48 movie leg (BASELINE configs[2]); 0 skips it")
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
