"""Frame-parallel movie projection (the loop of reference surface_projection.py:205-212).

The reference projects the time points of a movie one after the other on one CPU core.  Time points
are independent, so here they are

  * pipelined on each GPU through the frame slots of the C ABI (``tsp_frame_submit`` /
    ``tsp_frame_wait``): the pinned host->device copy of frame t+1 overlaps the kernels of frame t and
    the device->host copy of frame t-1;
  * partitioned over GPUs by time frame: ``devices=[0, 1, ...]`` runs one host thread per GPU in this
    process (ctypes releases the GIL), and under ``torchrun`` (one process per GPU) each rank takes the
    next unclaimed frame from a counter shared through the job's store (``SharedFrameCounter``) - the host
    links of an 8-GPU box are not equally fast (measured: 23 and 36 GB/s per GPU with all eight copying), a
    static split would wait for the slowest.  No collective runs on the data path; ranks only meet when the
    driver assembles the output arrays (``gather_movie``).
"""
from __future__ import annotations

import os
import threading

import numpy as np

from . import _native


def frame_owner(t, world_size):
    """Rank (or device slot) that projects time point ``t``: round-robin."""
    return t % world_size


def rank_world():
    """(rank, world_size) of this process: torch.distributed if initialised, else the torchrun environment."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:                                       # pragma: no cover
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class SharedFrameCounter:
    """Hands out 0, 1, 2, ... exactly once across the ranks of a torchrun job: an atomic add on the job's
    key-value store (the rendezvous TCPStore - control plane, ~0.1 ms per claim, no tensor traffic).  A single
    process counts locally.  Every rank must create its counters in the same program order (the key is a
    sequence number)."""

    _created = 0

    def __init__(self, tag="frames"):
        SharedFrameCounter._created += 1
        self.key = "tsp_b200/%s/%d" % (tag, SharedFrameCounter._created)
        self.local = 0
        self.store = None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self.store = dist.distributed_c10d._get_default_store()
        except (ImportError, AttributeError, RuntimeError):   # pragma: no cover
            self.store = None

    def next(self):
        if self.store is None:
            self.local += 1
            return self.local - 1
        return int(self.store.add(self.key, 1)) - 1

    def claims(self, total):
        """Iterate over the indices this rank manages to claim, in increasing order, until `total` is reached."""
        while True:
            i = self.next()
            if i >= total:
                return
            yield i


def gather_movie(arrays, owner_of_frame=None, world_size=None):
    """Combine per-rank partial movie arrays (axis 0 = time, frames a rank does not own are zero) onto every
    rank.  Output assembly only - uses the CPU (gloo) group when one exists; a single process is a no-op."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return arrays
    for a in arrays:
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return arrays


class FramePipeline:
    """Projects a sequence of frames on one or several GPUs of this process.

    ``operator`` is only a seam for host-side tests (it replaces the GPU call by a Python callable with the
    signature of ``time_point_surface_projection``); the product path is the C ABI.
    """

    def __init__(self, devices=None, slots=2, mode=None, operator=None):
        self.operator = operator
        self.mode = mode
        if operator is None:
            import torch
            if not torch.cuda.is_available():
                raise RuntimeError("FramePipeline needs a B200 GPU (there is no CPU fallback)")
            if devices is None:
                devices = [torch.cuda.current_device()]
        self.devices = list(devices) if devices is not None else [0]
        self.slots = max(1, min(int(slots), _native.MAX_SLOTS))

    # ---- one GPU: slot pipeline -------------------------------------------------------------------
    def _run_device(self, device, frames, params, sink):
        """frames: iterable of (t, stack) with stack a C-contiguous uint16 (C,Z,Y,X) host array (pinned memory
        makes the copies asynchronous).  sink(t, proj float64 (C,Y,X), zmap int64 (Y,X), status)."""
        mode = self.mode or params.get("mode") or os.environ.get("TSP_MODE", "fast")
        kw = dict(reference_channel=int(params.get("reference_channel", 0)), min_z=int(params.get("min_z", 0)),
                  max_z=int(params.get("max_z", 0)), airyscan=bool(params.get("airyscan", True)),
                  atoh_shift=int(params.get("atoh_shift", 0)), mode=mode, device=device)
        bin_size = int(params.get("bin_size", 1) or 1)
        if bin_size > 1:
            if params.get("method") not in _native.METHODS:
                raise TypeError("exceptions must derive from BaseException")      # SP:53 raises a str
            kw.update(bin_size=bin_size, method=params["method"])
        if params.get("build_manifold", False):
            kw.update(build_manifold=True)
        inflight = [None] * self.slots          # (t, stack, proj, zmap) per slot
        bufs = [None] * self.slots              # pinned output buffers per slot, reused while the shape holds

        def drain(slot):
            if inflight[slot] is None:
                return
            t, stack, proj, zmap = inflight[slot]
            status = _native.frame_wait(slot, device)
            inflight[slot] = None
            sink(t, proj, zmap, status)

        i = 0
        try:
            for t, stack in frames:
                slot = i % self.slots
                drain(slot)
                Cn, _, Y, X = stack.shape
                if bufs[slot] is None or bufs[slot][0].shape != (Cn, Y, X):
                    bufs[slot] = (_native.pinned_empty((Cn, Y, X), np.float64), _native.pinned_empty((Y, X), np.int64))
                proj, zmap = bufs[slot]
                _native.frame_submit(slot, stack, proj, zmap, **kw)
                inflight[slot] = (t, stack, proj, zmap)
                i += 1
            for k in range(self.slots):
                drain((i + k) % self.slots)
        finally:
            for slot in range(self.slots):           # never leave a slot busy behind an exception
                if inflight[slot] is not None:
                    try:
                        _native.frame_wait(slot, device)
                    except Exception:                # noqa: BLE001
                        pass
                    inflight[slot] = None

    # ---- public ------------------------------------------------------------------------------------
    def project_frames(self, frames, sink, **params):
        """Project ``frames`` (iterable of (t, stack)) and call ``sink(t, proj, zmap, status)`` for each.  The arrays
        handed to ``sink`` are reused for later frames: copy what must be kept.  With several devices the
        frames are dealt round-robin to one worker thread per GPU; ``sink`` is then called under a lock."""
        if self.operator is not None:
            for t, stack in frames:
                proj, zmap = self.operator(stack[None], axes="TCZYX", z_map=True,
                                           **{k: v for k, v in params.items() if k not in ("mode", "axes", "z_map")})
                sink(t, np.asarray(proj, dtype=np.float64), np.asarray(zmap, dtype=np.int64), {})
            return
        if len(self.devices) == 1:
            self._run_device(self.devices[0], frames, params, sink)
            return
        import queue
        lock = threading.Lock()
        queues = [queue.Queue(maxsize=2 * self.slots) for _ in self.devices]
        errors = []

        def locked_sink(*a):
            with lock:
                sink(*a)

        def worker(k):
            _native.bind_host_thread_to_gpu(self.devices[k])      # staging buffers next to the worker's GPU

            def gen():
                while True:
                    item = queues[k].get()
                    if item is None:
                        return
                    yield item
            try:
                self._run_device(self.devices[k], gen(), params, locked_sink)
            except Exception as exc:                 # noqa: BLE001
                errors.append(exc)
                while queues[k].get() is not None:   # keep the feeder from blocking
                    pass

        threads = [threading.Thread(target=worker, args=(k,), daemon=True) for k in range(len(self.devices))]
        for th in threads:
            th.start()
        for i, item in enumerate(frames):
            queues[frame_owner(i, len(self.devices))].put(item)
        for q in queues:
            q.put(None)
        for th in threads:
            th.join()
        if errors:
            raise errors[0]

    def project_movie(self, path, series, out_projection, out_zmap, mode=None, **params):
        """Driver hook used by ``movie_surface_projection``: read the time points of ``path`` (through the
        ``basic_image_manipulations.open_image`` hook), project the ones this rank claims (shared counter:
        every time point exactly once across the ranks, faster ranks take more) and scatter them into the
        (T,C,1,Y,X) / (T,1,1,Y,X) arrays of SP:201-202."""
        from . import basic_image_manipulations as bim
        img = bim.open_image(path)
        img.set_scene(series)
        data = img.get_image_dask_data()
        T = img.dims.T
        counter = SharedFrameCounter("movie")
        if mode is not None:
            params = dict(params, mode=mode)
        for k in ("axes", "z_map"):
            params.pop(k, None)

        def frames():
            for t in counter.claims(T):
                chunk = np.asarray(data[t:t + 1].compute())[0]          # (C, Z, Y, X)
                if chunk.dtype != np.uint16:
                    chunk = chunk.astype(np.uint16)
                yield t, np.ascontiguousarray(chunk)

        def sink(t, proj, zmap, status):
            out_projection[t, :, 0] = proj
            out_zmap[t, 0, 0] = zmap

        self.project_frames(frames(), sink, **params)
        gather_movie([out_projection, out_zmap])
