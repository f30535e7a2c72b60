"""Frame-parallel projection of movies and tiled images (the loops of reference surface_projection.py:205-212 and
:294-301).

The reference projects the time points of a movie (or the XY tiles of a large image) one after the other on one CPU
core.  They are independent, so here they are

  * pipelined on each GPU through the frame slots of the C ABI (``tsp_frame_submit`` / ``tsp_frame_wait``): the
    host->device copy of frame t+1 overlaps the kernels of frame t and the device->host copy of frame t-1.  Frames
    that do not already live in pinned memory are first copied into a ring of pinned staging buffers (one per slot,
    several host threads per copy), so every DMA is asynchronous whatever the reader returned;
  * returned in the dtype the caller stores: float64 / int64 like the reference operator, or - ``out_dtype="uint16"``,
    what the movie driver writes to disk (BIM:481, SP:229-231) - converted on the device, 4 instead of 16 bytes per
    pixel over the host link;
  * partitioned over GPUs by frame, dynamically: ``devices=[0, 1, ...]`` runs one host thread per GPU in this process
    and every thread takes the next frame from one shared queue; under ``torchrun`` (one process per GPU) each rank
    takes the next unclaimed frame from a counter shared through the job's store (``SharedFrameCounter``) - the host
    links of an 8-GPU box are not equally fast (measured: 23 and 36 GB/s per GPU with all eight copying), a static
    split would wait for the slowest.  No collective runs on the data path; ranks only meet when the driver assembles
    the output arrays on rank 0 (``gather_frames``: point-to-point sends of the owned frames over a gloo group).
"""
from __future__ import annotations

import atexit
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _native


def frame_owner(t, world_size):
    """Static round-robin owner of frame ``t`` (only used where no shared counter is available)."""
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


def init_job():
    """Join the job a launcher started (``torchrun``: WORLD_SIZE > 1 in the environment) unless the caller already
    did: a gloo process group - the ranks only meet on the control plane (frame claims, barriers) and when the output
    arrays are assembled - and this rank's GPU, chosen from the measured host links (``topology.choose_device``).
    Returns (rank, world_size).  A single process: nothing to do."""
    world = int(os.environ.get("WORLD_SIZE", "1") or 1)
    if world <= 1:
        return rank_world()
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    if torch.cuda.is_available():
        from . import topology
        local_rank = int(os.environ.get("LOCAL_RANK", dist.get_rank()))
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        device, _ = topology.choose_device(local_rank, local_world)
        torch.cuda.set_device(device)
        _native.bind_host_thread_to_gpu(device)
    return dist.get_rank(), dist.get_world_size()


def as_uint16_stack(stack):
    """The dtypes the B200 path takes: uint16 as is, uint8 widened; anything else is refused (the reference converts
    every dtype with ``astype('float32')``; silently wrapping int32 / float stacks into uint16 would corrupt them)."""
    stack = np.asarray(stack)
    if stack.dtype == np.uint16:
        return stack
    if stack.dtype == np.uint8:
        return stack.astype(np.uint16)
    raise TypeError("the B200 projection path takes uint8/uint16 stacks (got %s)" % stack.dtype)


# ----------------------------------------------------------------------------------------------------------------
# control plane of a multi-rank job
# ----------------------------------------------------------------------------------------------------------------
_store_lock = threading.Lock()
_job_store = None


def _job_kv_store():
    """A key-value store shared by the ranks of the job: the rendezvous store torch.distributed already runs (reached
    through c10d's accessor when this torch has it), else a TCPStore of our own next to the rendezvous port."""
    global _job_store
    import torch.distributed as dist
    with _store_lock:
        if _job_store is not None:
            return _job_store
        getter = getattr(dist.distributed_c10d, "_get_default_store", None)
        store = None
        if getter is not None:
            try:
                store = getter()
            except Exception:                                 # noqa: BLE001
                store = None
        if store is None:
            from datetime import timedelta
            port = int(os.environ.get("TSP_STORE_PORT", int(os.environ.get("MASTER_PORT", "29500")) + 17))
            store = dist.TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), port, dist.get_world_size(),
                                  is_master=dist.get_rank() == 0, timeout=timedelta(seconds=300))
        _job_store = dist.PrefixStore("tsp_b200", store)
        return _job_store


class SharedFrameCounter:
    """Hands out 0, 1, 2, ... exactly once across the ranks of a torchrun job: an atomic add on the job's key-value
    store (control plane, ~0.1 ms per claim, no tensor traffic).  A single process counts locally.  Every rank must
    create its counters in the same program order (the key is a sequence number)."""

    _created = 0

    def __init__(self, tag="frames"):
        SharedFrameCounter._created += 1
        self.key = "%s/%d" % (tag, SharedFrameCounter._created)
        self.local = 0
        self.store = None
        self._lock = threading.Lock()
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self.store = _job_kv_store()
        except ImportError:                                   # pragma: no cover
            self.store = None

    def next(self):
        if self.store is None:
            with self._lock:
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


_gloo_group = None


def host_group():
    """A process group that can move CPU tensors: the default group when it has a CPU backend (gloo), otherwise a
    gloo group created next to it (collective: every rank must reach the first call).  None for a single process."""
    global _gloo_group
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    if _gloo_group is None:
        backend = str(dist.get_backend()).lower()
        _gloo_group = dist.group.WORLD if "gloo" in backend else dist.new_group(backend="gloo")
    return _gloo_group


def gather_frames(arrays, owned, dst=0, all_ranks=False):
    """Assemble the frames each rank projected.  ``arrays``: per-rank arrays with axis 0 = frame; ``owned``: the
    frame indices this rank filled in.  Only owned frames travel: every rank sends them to ``dst`` point to point
    (gloo); with ``all_ranks`` the assembled arrays are broadcast back.  Returns True on the ranks that hold the
    complete arrays afterwards.  Output assembly only - nothing here is on the projection data path."""
    import torch
    import torch.distributed as dist
    group = host_group()
    if group is None:
        return True
    rank, world = dist.get_rank(), dist.get_world_size()
    idx = torch.tensor(sorted(int(t) for t in owned), dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([idx.numel()], dtype=torch.int64), group=group)
    def wire(a):                      # gloo moves bytes; uint16 (the movie dtype) is not one of its scalar types
        return torch.from_numpy(a).view(torch.uint8)

    if rank == dst:
        for src in range(world):
            n = int(counts[src].item())
            if src == dst or n == 0:
                continue
            their = torch.empty(n, dtype=torch.int64)
            dist.recv(their, src=src, group=group)
            for a in arrays:
                buf = np.empty((n,) + tuple(a.shape[1:]), dtype=a.dtype)
                dist.recv(wire(buf), src=src, group=group)
                a[their.numpy()] = buf
    elif idx.numel() > 0:
        dist.send(idx, dst=dst, group=group)
        for a in arrays:
            dist.send(wire(np.ascontiguousarray(a[idx.numpy()])), dst=dst, group=group)
    if all_ranks:
        for a in arrays:
            if not a.flags.c_contiguous:
                raise ValueError("gather_frames(all_ranks=True) needs C-contiguous arrays")
            dist.broadcast(wire(a), src=dst, group=group)
        return True
    return rank == dst


# ----------------------------------------------------------------------------------------------------------------
# output arrays of a job
# ----------------------------------------------------------------------------------------------------------------
_SHM_DIR = "/dev/shm"
_SHM_PREFIX = os.path.join(_SHM_DIR, "tsp_b200_out_")


_live_shm_paths = set()                      # backing files this process created and has not removed yet


def _unlink_shm(paths):
    for path in list(paths):
        try:
            os.unlink(path)
        except OSError:
            pass
        _live_shm_paths.discard(path)


atexit.register(lambda: _unlink_shm(_live_shm_paths))       # a job that dies half way leaves nothing in /dev/shm


def is_shared_output(a):
    """True for arrays (and views of arrays) handed out by ``allocate_outputs`` in its shared form."""
    while a is not None:
        if isinstance(a, np.memmap) and str(getattr(a, "filename", "") or "").startswith(_SHM_PREFIX):
            return True
        a = getattr(a, "base", None)
    return False


class JobOutputs:
    """The output arrays of one job (``arrays``), zero-filled.  ``shared`` says that every rank maps the same memory:
    whatever a rank writes is what rank 0 saves - no assembly step.  ``close()`` drops the mapping (rank 0 removes
    the backing files); call it on every rank once the arrays are not needed any more."""

    def __init__(self, arrays, shared, paths=()):
        self.arrays, self.shared, self._paths = list(arrays), bool(shared), list(paths)
        _live_shm_paths.update(self._paths)

    def close(self):
        self.arrays = []
        _unlink_shm(self._paths)
        self._paths = []


def allocate_outputs(specs):
    """``specs``: [(shape, dtype), ...] -> JobOutputs.  A single process gets ordinary arrays.  In a torchrun job whose
    ranks all run on one host (the layout of ``bench.py --gpus N`` and of the drivers: GPUs of ONE node) the arrays
    live in POSIX shared memory (files under /dev/shm, created by rank 0, mapped by everyone): each rank scatters the
    frames it projected straight into the job's arrays, which replaces shipping them to rank 0 through gloo (TCP
    loopback, ~1 GB/s: the 800 MB of a 200-frame 1024 x 1024 movie took longer than projecting it on eight GPUs).
    Ranks on several hosts, no /dev/shm, or TSP_NO_SHARED_OUTPUTS=1: ordinary per-rank arrays, assembled by
    ``gather_frames``.  Collective: every rank must call it with the same specs."""
    specs = [(tuple(int(v) for v in shape), np.dtype(dtype)) for shape, dtype in specs]
    group = host_group()
    if group is None:
        return JobOutputs([np.zeros(shape, dtype=dtype) for shape, dtype in specs], False)
    import socket
    import uuid
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    hosts = [None] * world
    dist.all_gather_object(hosts, socket.gethostname(), group=group)
    want = len(set(hosts)) == 1 and not os.environ.get("TSP_NO_SHARED_OUTPUTS") and os.path.isdir(_SHM_DIR)
    box = [None]
    arrays = []
    if rank == 0 and want:
        paths = ["%s%s_%d" % (_SHM_PREFIX, uuid.uuid4().hex, i) for i in range(len(specs))]
        try:
            for path, (shape, dtype) in zip(paths, specs):
                arrays.append(np.memmap(path, dtype=dtype, mode="w+", shape=shape))       # sparse file: reads as zeros
            box = [paths]
        except (OSError, ValueError):
            arrays = []
            for path in paths:
                if os.path.exists(path):
                    os.unlink(path)
    dist.broadcast_object_list(box, src=0, group=group)
    paths = box[0]
    ok = 1
    if paths is None:
        ok = 0
    elif rank != 0:
        try:
            arrays = [np.memmap(path, dtype=dtype, mode="r+", shape=shape) for path, (shape, dtype) in zip(paths, specs)]
        except (OSError, ValueError):
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int64)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 1:
        return JobOutputs(arrays, True, paths if rank == 0 else ())
    if rank == 0 and paths:
        for path in paths:
            if os.path.exists(path):
                os.unlink(path)
    return JobOutputs([np.zeros(shape, dtype=dtype) for shape, dtype in specs], False)


def finish_outputs(arrays, owned):
    """After every rank has scattered its frames into ``arrays``: make rank 0 hold the complete arrays.  Shared
    arrays only need everyone to be done (a barrier); ordinary ones are assembled by ``gather_frames``.  Returns True
    on the ranks that hold the complete arrays."""
    group = host_group()
    if group is None:
        return True
    import torch.distributed as dist
    if all(is_shared_output(a) for a in arrays):
        dist.barrier(group=group)
        if dist.get_rank() == 0:
            # rank 0 is about to read pages other ranks wrote: walking its own mapping once (one word per page) is
            # ~10 x cheaper than letting np.save take the page faults inside write() (0.03 s against 0.19 s per 400 MB)
            for a in arrays:
                flat = a.reshape(-1)
                step = max(1, 4096 // max(1, a.dtype.itemsize))
                int(flat[::step].view(np.ndarray).astype(np.int64, copy=False).sum())
        return True
    return gather_frames(arrays, owned, all_ranks=False)


def gather_movie(arrays, owned=None, all_ranks=True):
    """Backwards-compatible name: assemble per-rank movie arrays on every rank (``owned`` = the time points this rank
    projected; None = every frame that is non-zero here)."""
    if owned is None:
        owned = [t for t in range(arrays[0].shape[0]) if np.any(arrays[0][t])]
    gather_frames(arrays, owned, all_ranks=all_ranks)
    return arrays


# ----------------------------------------------------------------------------------------------------------------
# the pipeline
# ----------------------------------------------------------------------------------------------------------------
class _Staging:
    """Pinned staging buffers of one GPU worker (one per frame slot, reallocated when the frame shape changes) and
    the host threads that fill them."""

    def __init__(self, slots, copy_threads):
        self.bufs = [None] * slots
        self.pool = ThreadPoolExecutor(max_workers=copy_threads) if copy_threads > 1 else None
        self.copy_threads = copy_threads

    def close(self):
        if self.pool is not None:
            self.pool.shutdown(wait=True)

    @staticmethod
    def is_pinned(arr):
        import torch
        if not arr.flags.writeable:            # a read-only view of a file mapping (tiff_io) is never pinned memory
            return False
        try:
            return arr.flags.c_contiguous and torch.from_numpy(arr).is_pinned()
        except (TypeError, ValueError, RuntimeError):
            return False

    def owns(self, slot, arr):
        """True when ``arr`` is the staging buffer of ``slot`` (False: the frame was pinned already and passed through)."""
        return self.bufs[slot] is not None and arr is self.bufs[slot]

    def stage(self, slot, stack):
        """-> C-contiguous uint16 (C,Z,Y,X) array in pinned memory holding ``stack`` (itself when already pinned).
        A lazy block of a file (``tiff_io``: has ``read_into``) is read straight into the pinned buffer by the copy
        threads - the only host copy of that frame."""
        if hasattr(stack, "read_into"):
            if stack.ndim != 4:
                raise RuntimeError("sequence argument must have length equal to input rank")
            dt = np.dtype(stack.dtype)
            if dt.kind == "u" and dt.itemsize == 2:
                buf = self.bufs[slot]
                if buf is None or buf.shape != tuple(stack.shape):
                    buf = self.bufs[slot] = _native.pinned_empty(tuple(stack.shape), np.uint16)
                stack.read_into(buf, threads=self.copy_threads, pool=self.pool)
                return buf
            stack = stack.compute()
        stack = as_uint16_stack(stack)
        if stack.ndim != 4:
            raise RuntimeError("sequence argument must have length equal to input rank")
        if self.is_pinned(stack):
            return stack
        buf = self.bufs[slot]
        if buf is None or buf.shape != stack.shape:
            buf = self.bufs[slot] = _native.pinned_empty(stack.shape, np.uint16)
        n = stack.shape[0] * stack.shape[1]
        if self.pool is None or stack.nbytes < (8 << 20) or n < 2:
            np.copyto(buf, stack)
        else:                                      # plane ranges in parallel: numpy releases the GIL while it copies
            src = stack.reshape((n,) + stack.shape[2:])
            dst = buf.reshape((n,) + stack.shape[2:])
            parts = min(self.copy_threads, n)
            cuts = [n * k // parts for k in range(parts + 1)]
            list(self.pool.map(lambda k: np.copyto(dst[cuts[k]:cuts[k + 1]], src[cuts[k]:cuts[k + 1]]), range(parts)))
        return buf


class FramePipeline:
    """Projects a sequence of frames on one or several GPUs of this process.

    ``out_dtype``: "reference" (float64 projection, int64 height map - what the operator returns) or "uint16" (both
    uint16, converted on the device - what the movie driver stores).
    ``operator`` is only a seam for host-side tests (it replaces the GPU call by a Python callable with the
    signature of ``time_point_surface_projection``); the product path is the C ABI.
    """

    def __init__(self, devices=None, slots=2, mode=None, operator=None, out_dtype="reference", copy_threads=None):
        if out_dtype not in ("reference", "uint16"):
            raise ValueError("out_dtype must be 'reference' or 'uint16'")
        self.operator = operator
        self.mode = mode
        self.out_dtype = out_dtype
        if operator is None:
            import torch
            if not torch.cuda.is_available():
                raise RuntimeError("FramePipeline needs a B200 GPU (there is no CPU fallback)")
            if devices is None:
                devices = [torch.cuda.current_device()]
        self.devices = list(devices) if devices is not None else [0]
        self.slots = max(1, min(int(slots), _native.MAX_SLOTS))
        if copy_threads is None:
            # measured on the 16-CPU B200 box (tools/staging_probe.py, one 96 MiB frame into pinned memory): 1 thread
            # 13.8 GB/s, 8 threads 33.8, 16 threads 48.2 - the host link takes 47-54: the CPUs this process can use,
            # shared with the other GPUs it feeds and with the other ranks of the node, up to 16
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
            share = max(1, len(self.devices)) * max(1, local_world)
            copy_threads = max(1, min(16, len(os.sched_getaffinity(0)) // share))
        self.copy_threads = int(copy_threads)
        self.h2d_bytes = 0                       # bytes handed to the GPUs so far (bench.py reports GB/s from it)

    # ---- one GPU: slot pipeline -------------------------------------------------------------------
    def _operator_kwargs(self, params, device):
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
        extra = {k: params[k] for k in _native.PARAM_KEYS if params.get(k) is not None}
        if extra:
            kw.update(params=extra)
        return kw

    def _run_device(self, device, frames, params, sink):
        """frames: iterable of (key, stack) with stack a uint8/uint16 (C,Z,Y,X) host array.
        sink(key, proj (C,Y,X), zmap (Y,X), status).

        Two host threads: a feeder reads the next frame and copies it into a free pinned staging buffer (slots + 2 of
        them: one per frame in flight, one in this thread's hands, one being filled) while this thread submits staged
        frames to the GPU's frame slots, waits for finished ones and hands them to ``sink`` - reading / staging frame
        t+1 overlaps everything else of frames t, t-1.  The feeder inherits the CPU affinity of the thread that runs
        this method (its staging buffers are first touched next to the GPU)."""
        kw = self._operator_kwargs(params, device)
        pdt, zdt = (np.uint16, np.uint16) if self.out_dtype == "uint16" else (np.float64, np.int64)
        inflight = [None] * self.slots          # (key, pinned stack, staging id, proj, zmap) per slot
        bufs = [None] * self.slots              # pinned output buffers per slot, reused while the shape holds
        nstage = self.slots + 2
        staging = _Staging(nstage, self.copy_threads)
        free_ids = queue.Queue()
        for b in range(nstage):
            free_ids.put(b)
        staged = queue.Queue(maxsize=nstage)
        stop = threading.Event()

        def feeder():
            try:
                for key, stack in frames:
                    b = None
                    while b is None:
                        if stop.is_set():
                            return
                        try:
                            b = free_ids.get(timeout=0.1)
                        except queue.Empty:
                            pass
                    staged.put((key, staging.stage(b, stack), b, None))
                staged.put(None)
            except BaseException as exc:                 # noqa: BLE001 - re-raised by the consumer
                staged.put((None, None, None, exc))

        def drain(slot):
            if inflight[slot] is None:
                return
            key, pinned, b, proj, zmap = inflight[slot]
            status = _native.frame_wait(slot, device)
            inflight[slot] = None
            if b is not None:
                free_ids.put(b)
            sink(key, proj, zmap, status)

        th = threading.Thread(target=feeder, daemon=True, name="tsp-feeder-gpu%s" % device)
        th.start()
        i = 0
        try:
            while True:
                item = staged.get()
                if item is None:
                    break
                key, pinned, b, exc = item
                if exc is not None:
                    raise exc
                slot = i % self.slots
                drain(slot)
                Cn, _, Y, X = pinned.shape
                if bufs[slot] is None or bufs[slot][0].shape != (Cn, Y, X):
                    bufs[slot] = (_native.pinned_empty((Cn, Y, X), pdt), _native.pinned_empty((Y, X), zdt))
                proj, zmap = bufs[slot]
                _native.frame_submit(slot, pinned, proj, zmap, **kw)
                self.h2d_bytes += pinned.nbytes
                inflight[slot] = (key, pinned, b if staging.owns(b, pinned) else None, proj, zmap)
                if inflight[slot][2] is None:
                    free_ids.put(b)                  # the frame was already pinned: its staging buffer was not used
                i += 1
            for k in range(self.slots):
                drain((i + k) % self.slots)
        finally:
            stop.set()
            for slot in range(self.slots):           # never leave a slot busy behind an exception
                if inflight[slot] is not None:
                    try:
                        _native.frame_wait(slot, device)
                    except Exception:                # noqa: BLE001
                        pass
                    inflight[slot] = None
            while th.is_alive():                     # unblock a feeder waiting on a full queue
                try:
                    staged.get_nowait()
                except queue.Empty:
                    pass
                th.join(timeout=0.05)
            staging.close()

    # ---- public ------------------------------------------------------------------------------------
    def frame_of(self, chunk):
        """The (C,Z,Y,X) frame of a lazily sliced (1,C,Z,Y,X) chunk of an image: an ndarray (``.compute()``), or -
        for a block of a TIFF file (``tiff_io``: has ``read_into``) on the GPU path - the lazy block itself, which
        the staging threads read straight into pinned memory (TSP_TIFF_MMAP=1: a view of the file mapping instead)."""
        if hasattr(chunk, "read_into") and self.operator is None and not os.environ.get("TSP_TIFF_MMAP"):
            return chunk[0]
        return np.asarray(chunk.compute())[0]

    def project_frames(self, frames, sink, **params):
        """Project ``frames`` (iterable of (key, stack)) and call ``sink(key, proj, zmap, status)`` for each.  The
        arrays handed to ``sink`` are reused for later frames: copy what must be kept.  With several devices one
        worker thread per GPU takes the next frame from a shared queue (a GPU on a faster host link takes more);
        ``sink`` is then called under a lock."""
        if self.operator is not None:
            skip = ("mode", "axes", "z_map") + _native.PARAM_KEYS
            for key, stack in frames:
                stack = as_uint16_stack(stack)
                proj, zmap = self.operator(stack[None], axes="TCZYX", z_map=True,
                                           **{k: v for k, v in params.items() if k not in skip})
                if self.out_dtype == "uint16":
                    sink(key, np.asarray(proj).astype(np.uint16), np.asarray(zmap).astype(np.uint16), {})
                else:
                    sink(key, np.asarray(proj, dtype=np.float64), np.asarray(zmap, dtype=np.int64), {})
            return
        if len(self.devices) == 1:
            self._run_device(self.devices[0], frames, params, sink)
            return
        lock = threading.Lock()
        work = queue.Queue(maxsize=2 * self.slots * len(self.devices))
        errors = []

        def locked_sink(*a):
            with lock:
                sink(*a)

        def worker(k):
            _native.bind_host_thread_to_gpu(self.devices[k])      # staging buffers next to the worker's GPU

            def gen():
                while not errors:
                    item = work.get()
                    if item is None:
                        return
                    yield item
            try:
                self._run_device(self.devices[k], gen(), params, locked_sink)
            except Exception as exc:                 # noqa: BLE001
                errors.append(exc)

        threads = [threading.Thread(target=worker, args=(k,), daemon=True) for k in range(len(self.devices))]
        for th in threads:
            th.start()

        def put(item):                               # never blocks on a queue nobody reads any more
            while any(th.is_alive() for th in threads):
                try:
                    work.put(item, timeout=0.2)
                    return True
                except queue.Full:
                    pass
            return False

        for item in frames:
            if errors or not put(item):
                break
        for _ in threads:
            put(None)
        for th in threads:
            th.join()
        if errors:
            raise errors[0]

    def project_movie(self, path, series, out_projection, out_zmap, mode=None, gather="all", **params):
        """Project the time points of ``path`` / ``series`` (read through the ``basic_image_manipulations.open_image``
        hook) that this rank claims (shared counter: every time point exactly once across the ranks, faster ranks
        take more) and scatter them into ``out_projection`` (T,C,1,Y,X) / ``out_zmap`` (T,1,1,Y,X) - the arrays of
        SP:201-202, in whatever dtype the caller allocated.  ``gather``: "all" assembles the arrays on every rank,
        "root" on rank 0 only, None leaves each rank with its own frames; arrays from ``allocate_outputs`` that all
        ranks map need no assembly (only a barrier).  Returns the time points projected here."""
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
        owned = []

        def frames():
            for t in counter.claims(T):
                yield t, self.frame_of(data[t:t + 1])                    # (C, Z, Y, X)

        def sink(t, proj, zmap, status):
            out_projection[t, :, 0] = proj
            out_zmap[t, 0, 0] = zmap
            owned.append(t)

        self.project_frames(frames(), sink, **params)
        if gather:
            if is_shared_output(out_projection) and is_shared_output(out_zmap):
                finish_outputs([out_projection, out_zmap], owned)         # arrays of allocate_outputs: a barrier
            else:
                gather_frames([out_projection, out_zmap], owned, all_ranks=(gather == "all"))
        return owned
