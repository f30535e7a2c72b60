"""ctypes binding of libtsp_b200.so (C ABI in include/tsp_b200.h).

There is no CPU fallback: if the shared library is missing or no B200 is visible the calls raise.
PyTorch is used only for device buffers, pinned host buffers and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtsp_b200.so")

MODE_FAST, MODE_EXACT, MODE_BITEXACT = 0, 1, 2
MODES = {"fast": MODE_FAST, "exact": MODE_EXACT, "bitexact": MODE_BITEXACT}

ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_BAND_INDEX, ERR_CHOOSE_LIMIT = -1, -2, -3, -4, -5


ABI_VERSION = 3


class Params(C.Structure):
    """tsp_params: the constants the reference hard-codes (SP:28, SP:35, SP:37, SP:55, SP:70-71)."""
    _fields_ = [("percentile", C.c_float), ("pedestal", C.c_int32), ("sigma_pre", C.c_float * 3),
                ("sigma_score", C.c_float * 3), ("sigma_mask", C.c_float * 3), ("reserved", C.c_int32 * 5)]


class FrameDesc(C.Structure):
    _fields_ = [("channels", C.c_int32), ("planes", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32),
                ("reference_channel", C.c_int32), ("min_z", C.c_int32), ("max_z", C.c_int32),
                ("airyscan", C.c_int32), ("atoh_shift", C.c_int32), ("mode", C.c_int32),
                ("bin_size", C.c_int32), ("method", C.c_int32), ("build_manifold", C.c_int32),
                ("flags", C.c_int32), ("has_params", C.c_int32), ("reserved", C.c_int32), ("params", Params)]


class FrameStatus(C.Structure):
    _fields_ = [("band_index_error", C.c_int32), ("has_nonzero", C.c_int32), ("percentile95", C.c_float),
                ("zmap_min", C.c_int32), ("zmap_max", C.c_int32), ("nonzero_count", C.c_int64),
                ("near_tie_pixels", C.c_int32), ("reserved", C.c_int32 * 5)]


EXPORTS = {
    # name: (restype, argtypes)
    "tsp_abi_version": (C.c_int, []),
    "tsp_last_error": (C.c_char_p, []),
    "tsp_default_params": (None, [C.POINTER(Params)]),
    "tsp_debug_set": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "tsp_percentile_nonzero_u16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_void_p,
                                             C.POINTER(FrameStatus)]),
    "tsp_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "tsp_destroy": (C.c_int, [C.c_void_p]),
    "tsp_project_workspace_bytes": (C.c_size_t, [C.POINTER(FrameDesc)]),
    "tsp_project_frame": (C.c_int, [C.c_void_p, C.POINTER(FrameDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "tsp_get_frame_status": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FrameStatus)]),
    "tsp_project_frame_host": (C.c_int, [C.c_void_p, C.POINTER(FrameDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(FrameStatus)]),
    "tsp_frame_submit": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(FrameDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "tsp_frame_wait": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(FrameStatus)]),
    "tsp_gaussian_blur_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p]),
    "tsp_gaussian_blur_u16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    "tsp_percentile95_nonzero_u16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                               C.POINTER(FrameStatus)]),
    "tsp_focus_score_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p]),
    "tsp_argmax_z_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p]),
    "tsp_band_project": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tsp_band_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "tsp_manifold_workspace_bytes": (C.c_size_t, []),
    "tsp_build_manifold": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_size_t, C.c_void_p]),
    "tsp_block_reduce_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p]),
    "tsp_resize_argmax_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tsp_resize_round_i32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p]),
    "tsp_project_m_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "tsp_project_m": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p, C.c_size_t, C.c_void_p]),
    "tsp_debug_coarse_taps": (C.c_int, [C.POINTER(C.c_double), C.c_int]),
    "tsp_launch_count": (C.c_int64, [C.c_void_p]),
    "tsp_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "tsp_stage_count": (C.c_int, []),
    "tsp_stage_name": (C.c_char_p, [C.c_int]),
    "tsp_get_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
}

_lib = None
_lib_lock = threading.Lock()
_handles = {}


class NativeLibraryMissing(ImportError):
    pass


def load_library():
    """dlopen libtsp_b200.so and declare every symbol of include/tsp_b200.h.  Loading needs no GPU."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise NativeLibraryMissing(
                "%s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in EXPORTS.items():
            fn = getattr(lib, name)          # AttributeError here = header and library disagree
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.tsp_abi_version() != ABI_VERSION:
            raise NativeLibraryMissing("%s has ABI %d, this binding needs %d - rebuild it"
                                       % (LIB_PATH, lib.tsp_abi_version(), ABI_VERSION))
        _lib = lib
        return lib


def last_error():
    return load_library().tsp_last_error().decode("utf-8", "replace")


class TspError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        super().__init__("%s failed (%d): %s" % (where, code, last_error()))


def check(code, where):
    """Map C status codes onto the exceptions the reference raises."""
    if code == 0:
        return
    if code == ERR_BAND_INDEX:
        raise IndexError("index out of bounds for the cropped z axis (reference surface_projection.py:68-69): "
                         + last_error())
    if code == ERR_CHOOSE_LIMIT:
        raise ValueError("Need at least 0 and at most 64 array objects.")
    raise TspError(code, where)


def handle(device=None):
    """One tsp_handle per (process, GPU)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("tissue_image_processing_b200 needs a B200 GPU (there is no CPU fallback)")
    if device is None:
        device = torch.cuda.current_device()
    device = int(device)
    lib = load_library()
    with _lib_lock:
        h = _handles.get(device)
        if h is None:
            out = C.c_void_p()
            rc = lib.tsp_create(device, C.byref(out))
            if rc != 0:
                raise TspError(rc, "tsp_create")
            h = _handles[device] = out
        return h


def launch_count(device=None):
    return int(load_library().tsp_launch_count(handle(device)))


def set_profiling(enable, device=None):
    check(load_library().tsp_set_profiling(handle(device), 1 if enable else 0), "tsp_set_profiling")


def stage_times(reset=True, device=None):
    """{stage name: (total ms, intervals)} recorded since the last reset (synchronises the device)."""
    lib = load_library()
    n = lib.tsp_stage_count()
    ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
    check(lib.tsp_get_stage_times(handle(device), ms, cnt, 1 if reset else 0), "tsp_get_stage_times")
    return {lib.tsp_stage_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}


METHODS = {"max_averages": 0, "max_std": 1, "multi_channel": 2}


FRAME_CONCURRENT = 1        # desc.flags: frames of other streams are in flight on this GPU (no chained launches)
FRAME_OUT_U16 = 2           # host-buffer calls: uint16 projection / height map (the movie driver's on-disk dtype)

PARAM_KEYS = ("percentile", "pedestal", "sigma_pre", "sigma_score", "sigma_mask")


def default_params():
    """The reference's constants as a dict (from the library: tsp_default_params)."""
    p = Params()
    load_library().tsp_default_params(C.byref(p))
    return {"percentile": float(p.percentile), "pedestal": int(p.pedestal), "sigma_pre": tuple(p.sigma_pre),
            "sigma_score": tuple(p.sigma_score), "sigma_mask": tuple(p.sigma_mask)}


def debug_set(key, value, device=None):
    """Kernel-variant switch for tests / A-B measurements (tsp_debug_set)."""
    check(load_library().tsp_debug_set(handle(device), key.encode(), int(value)), "tsp_debug_set")


def make_desc(C_, Z, Y, X, reference_channel=0, min_z=0, max_z=0, airyscan=False, atoh_shift=0, mode="fast",
              bin_size=1, method="max_averages", build_manifold=False, concurrent=False, out_u16=False,
              params=None):
    """params: dict with any of PARAM_KEYS (None values = the reference's constant)."""
    d = FrameDesc()
    d.flags = (FRAME_CONCURRENT if concurrent else 0) | (FRAME_OUT_U16 if out_u16 else 0)
    if params and set(params) - set(PARAM_KEYS):
        raise TypeError("unknown projection parameters: %s" % sorted(set(params) - set(PARAM_KEYS)))
    if params and any(params.get(k) is not None for k in PARAM_KEYS):
        d.has_params = 1
        load_library().tsp_default_params(C.byref(d.params))
        if params.get("percentile") is not None:
            d.params.percentile = float(params["percentile"])
        if params.get("pedestal") is not None:
            d.params.pedestal = int(params["pedestal"])
        for key in ("sigma_pre", "sigma_score", "sigma_mask"):
            if params.get(key) is not None:
                vals = tuple(float(v) for v in params[key])
                if len(vals) != 3:
                    raise RuntimeError("sequence argument must have length equal to input rank")
                setattr(d.params, key, (C.c_float * 3)(*vals))
    d.bin_size = int(bin_size)
    d.method = METHODS[method] if isinstance(method, str) else int(method)
    d.build_manifold = 1 if build_manifold else 0
    d.channels, d.planes, d.rows, d.cols = int(C_), int(Z), int(Y), int(X)
    d.reference_channel = int(reference_channel)
    d.min_z, d.max_z = int(min_z), int(max_z)
    d.airyscan = 1 if airyscan else 0
    d.atoh_shift = int(atoh_shift)
    d.mode = MODES[mode] if isinstance(mode, str) else int(mode)
    return d


def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def status_dict(st):
    return {"band_index_error": bool(st.band_index_error), "has_nonzero": bool(st.has_nonzero),
            "percentile95": float(st.percentile95), "zmap_min": int(st.zmap_min), "zmap_max": int(st.zmap_max),
            "nonzero_count": int(st.nonzero_count), "near_tie_pixels": int(st.near_tie_pixels)}


# --------------------------------------------------------------------------------------------------
# host-buffer operator (the plugin boundary)
# --------------------------------------------------------------------------------------------------
def pinned_empty(shape, dtype):
    """numpy array backed by pinned host memory (DMA-able), kept alive by its torch tensor."""
    import torch
    tdtype = {np.dtype("float64"): torch.float64, np.dtype("int64"): torch.int64,
              np.dtype("uint16"): torch.uint16, np.dtype("float32"): torch.float32,
              np.dtype("int32"): torch.int32}[np.dtype(dtype)]
    return torch.empty(tuple(int(s) for s in shape), dtype=tdtype, pin_memory=True).numpy()


def bind_host_thread_to_gpu(device=None):
    """Pin the calling host thread to the CPUs next to ``device`` (NVML's ideal affinity) so that the pinned staging
    buffers it allocates afterwards are first-touched on the GPU's own NUMA node - with eight GPUs feeding from one
    host the copies otherwise cross the socket interconnect.  Best effort: returns the CPU list, or None when NVML
    or the topology information is not available (nothing changes then).  TSP_NO_AFFINITY=1 disables it."""
    if os.environ.get("TSP_NO_AFFINITY"):
        return None
    try:
        import pynvml
        import torch
        dev = torch.cuda.current_device() if device is None else int(device)
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = os.sched_getaffinity(0)
        if not after:                                     # never leave the thread without CPUs
            os.sched_setaffinity(0, before)
            return None
        return sorted(after)
    except Exception:                                     # noqa: BLE001 - topology hints are optional
        return None


def project_frame_host(stack, reference_channel, min_z=0, max_z=0, airyscan=False, atoh_shift=0, mode="fast",
                       device=None, out_proj=None, out_zmap=None, bin_size=1, method="max_averages",
                       build_manifold=False, params=None, out_u16=False):
    """stack: C-contiguous uint16 ndarray (C,Z,Y,X) in host memory.  Returns (projection float64 (C,Y,X),
    zmap int64 (Y,X), status dict); uint16 both with ``out_u16`` (the movie driver's on-disk dtype)."""
    lib = load_library()
    h = handle(device)
    assert stack.dtype == np.uint16 and stack.ndim == 4 and stack.flags.c_contiguous
    Cn, Z, Y, X = stack.shape
    desc = make_desc(Cn, Z, Y, X, reference_channel, min_z, max_z, airyscan, atoh_shift, mode, bin_size, method,
                     build_manifold, out_u16=out_u16, params=params)
    pdt, zdt = (np.uint16, np.uint16) if out_u16 else (np.float64, np.int64)
    proj = out_proj if out_proj is not None else pinned_empty((Cn, Y, X), pdt)
    zmap = out_zmap if out_zmap is not None else pinned_empty((Y, X), zdt)
    assert proj.dtype == pdt and zmap.dtype == zdt and proj.flags.c_contiguous and zmap.flags.c_contiguous
    st = FrameStatus()
    rc = lib.tsp_project_frame_host(h, C.byref(desc), C.c_void_p(stack.ctypes.data), C.c_void_p(proj.ctypes.data),
                                    C.c_void_p(zmap.ctypes.data), C.byref(st))
    check(rc, "tsp_project_frame_host")
    return proj, zmap, status_dict(st)


MAX_SLOTS = 4


def frame_submit(slot, stack, proj, zmap, reference_channel, min_z=0, max_z=0, airyscan=False, atoh_shift=0,
                 mode="fast", device=None, bin_size=1, method="max_averages", build_manifold=False, params=None):
    """Enqueue one frame on a slot (asynchronous).  stack (C,Z,Y,X) uint16, proj (C,Y,X) and zmap (Y,X) are host
    arrays that must stay alive (and untouched) until frame_wait(slot); float64 / int64 outputs are the reference's
    dtypes, uint16 / uint16 selects the on-device conversion (TSP_FRAME_OUT_U16)."""
    lib = load_library()
    assert stack.dtype == np.uint16 and stack.ndim == 4 and stack.flags.c_contiguous
    Cn, Z, Y, X = stack.shape
    out_u16 = proj.dtype == np.uint16
    assert (proj.dtype, zmap.dtype) in ((np.float64, np.int64), (np.uint16, np.uint16))
    assert proj.shape == (Cn, Y, X) and proj.flags.c_contiguous
    assert zmap.shape == (Y, X) and zmap.flags.c_contiguous
    desc = make_desc(Cn, Z, Y, X, reference_channel, min_z, max_z, airyscan, atoh_shift, mode, bin_size, method,
                     build_manifold, out_u16=out_u16, params=params)
    rc = lib.tsp_frame_submit(handle(device), int(slot), C.byref(desc), C.c_void_p(stack.ctypes.data),
                              C.c_void_p(proj.ctypes.data), C.c_void_p(zmap.ctypes.data))
    check(rc, "tsp_frame_submit")


def frame_wait(slot, device=None):
    st = FrameStatus()
    rc = load_library().tsp_frame_wait(handle(device), int(slot), C.byref(st))
    check(rc, "tsp_frame_wait")
    return status_dict(st)


# --------------------------------------------------------------------------------------------------
# device-resident operator
# --------------------------------------------------------------------------------------------------
class DeviceProjector:
    """Reusable device-side buffers for one frame shape; everything stays on the current stream."""

    def __init__(self, Cn, Z, Y, X, reference_channel=0, min_z=0, max_z=0, airyscan=False, atoh_shift=0,
                 mode="fast", device=None, bin_size=1, method="max_averages", build_manifold=False,
                 concurrent=False, params=None):
        """concurrent=True: several projectors of this GPU run frames on different streams at the same time
        (throughput mode, plain launches); False chains the kernels for the lowest single-frame latency."""
        import torch
        self.lib = load_library()
        self.h = handle(device)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.desc = make_desc(Cn, Z, Y, X, reference_channel, min_z, max_z, airyscan, atoh_shift, mode, bin_size,
                              method, build_manifold, concurrent, params=params)
        self.ws_bytes = int(self.lib.tsp_project_workspace_bytes(C.byref(self.desc)))
        if self.ws_bytes == 0:
            raise TspError(ERR_INVALID, "tsp_project_workspace_bytes")
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.proj = torch.empty((Cn, Y, X), dtype=torch.float32, device=self.device)
        self.zmap = torch.empty((Y, X), dtype=torch.int32, device=self.device)
        # torch's caching allocator may have handed out blocks that earlier work of the allocating stream is still
        # writing (temporaries freed a moment ago): a frame launched on ANOTHER stream must be ordered behind that
        with torch.cuda.device(self.device):
            self._ready = torch.cuda.Event()
            self._ready.record()
        self._ordered = set()

    def run(self, d_stack, stream=None):
        """d_stack: CUDA uint16 tensor (C,Z,Y,X), contiguous.  Asynchronous."""
        assert d_stack.is_cuda and d_stack.is_contiguous() and d_stack.element_size() == 2
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device)
        if stream.cuda_stream not in self._ordered:
            stream.wait_event(self._ready)
            self._ordered.add(stream.cuda_stream)
        rc = self.lib.tsp_project_frame(self.h, C.byref(self.desc), C.c_void_p(d_stack.data_ptr()),
                                        C.c_void_p(self.proj.data_ptr()), C.c_void_p(self.zmap.data_ptr()),
                                        C.c_void_p(self.workspace.data_ptr()), self.ws_bytes, _stream_ptr(stream))
        check(rc, "tsp_project_frame")
        return self.proj, self.zmap

    def status(self, stream=None):
        st = FrameStatus()
        rc = self.lib.tsp_get_frame_status(self.h, C.c_void_p(self.workspace.data_ptr()), _stream_ptr(stream),
                                           C.byref(st))
        check(rc, "tsp_get_frame_status")
        return status_dict(st)


# --------------------------------------------------------------------------------------------------
# building blocks (device tensors in, device tensors out)
# --------------------------------------------------------------------------------------------------
def _sig3(sigma):
    return (C.c_double * 3)(*[float(s) for s in sigma])


def gaussian_blur(d_vol, sigma, fp64_accumulate=True, stream=None):
    """scipy.ndimage.gaussian_filter(mode='nearest') of a 3-D CUDA tensor (float32 or uint16)."""
    import torch
    lib, h = load_library(), handle(d_vol.device.index)
    assert d_vol.is_cuda and d_vol.dim() == 3 and d_vol.is_contiguous()
    Z, Y, X = d_vol.shape
    out, tmp = torch.empty_like(d_vol), torch.empty_like(d_vol)
    if d_vol.dtype == torch.float32:
        rc = lib.tsp_gaussian_blur_f32(h, C.c_void_p(d_vol.data_ptr()), C.c_void_p(out.data_ptr()),
                                       C.c_void_p(tmp.data_ptr()), Z, Y, X, _sig3(sigma),
                                       1 if fp64_accumulate else 0, _stream_ptr(stream))
    elif d_vol.dtype == torch.uint16:
        rc = lib.tsp_gaussian_blur_u16(h, C.c_void_p(d_vol.data_ptr()), C.c_void_p(out.data_ptr()),
                                       C.c_void_p(tmp.data_ptr()), Z, Y, X, _sig3(sigma), _stream_ptr(stream))
    else:
        raise TypeError("gaussian_blur supports float32 and uint16 volumes, got %s" % d_vol.dtype)
    check(rc, "tsp_gaussian_blur")
    return out


def percentile95_nonzero(d_vol, airyscan=False, stream=None):
    lib, h = load_library(), handle(d_vol.device.index)
    assert d_vol.is_cuda and d_vol.is_contiguous() and d_vol.element_size() == 2
    st = FrameStatus()
    rc = lib.tsp_percentile95_nonzero_u16(h, C.c_void_p(d_vol.data_ptr()), d_vol.numel(), 1 if airyscan else 0,
                                          _stream_ptr(stream), C.byref(st))
    check(rc, "tsp_percentile95_nonzero_u16")
    return status_dict(st)


def percentile_nonzero(d_vol, percentile, pedestal=0, stream=None):
    """np.percentile(v[v > pedestal] - pedestal, percentile) of a uint16 CUDA tensor, numpy's float32 arithmetic."""
    lib, h = load_library(), handle(d_vol.device.index)
    assert d_vol.is_cuda and d_vol.is_contiguous() and d_vol.element_size() == 2
    st = FrameStatus()
    rc = lib.tsp_percentile_nonzero_u16(h, C.c_void_p(d_vol.data_ptr()), d_vol.numel(), float(percentile),
                                        int(pedestal), _stream_ptr(stream), C.byref(st))
    check(rc, "tsp_percentile_nonzero_u16")
    return status_dict(st)


def focus_score(d_channel, airyscan=False, fp64_accumulate=False, stream=None):
    import torch
    lib, h = load_library(), handle(d_channel.device.index)
    assert d_channel.is_cuda and d_channel.dim() == 3 and d_channel.is_contiguous()
    Z, Y, X = d_channel.shape
    score = torch.empty((Z, Y, X), dtype=torch.float32, device=d_channel.device)
    tmp = torch.empty_like(score)
    rc = lib.tsp_focus_score_f32(h, C.c_void_p(d_channel.data_ptr()), C.c_void_p(score.data_ptr()),
                                 C.c_void_p(tmp.data_ptr()), Z, Y, X, 1 if airyscan else 0,
                                 1 if fp64_accumulate else 0, _stream_ptr(stream))
    check(rc, "tsp_focus_score_f32")
    return score


def argmax_z(d_score, z_offset=0, stream=None):
    import torch
    lib, h = load_library(), handle(d_score.device.index)
    Z, Y, X = d_score.shape
    zmap = torch.empty((Y, X), dtype=torch.int32, device=d_score.device)
    rc = lib.tsp_argmax_z_f32(h, C.c_void_p(d_score.data_ptr()), C.c_void_p(zmap.data_ptr()), Z, Y, X,
                              int(z_offset), _stream_ptr(stream))
    check(rc, "tsp_argmax_z_f32")
    return zmap


def build_manifold_device(d_score, stream=None):
    """SP:87-128 on a device score volume (Z,Y,X) float32 -> (Y,X) int32."""
    import torch
    lib, h = load_library(), handle(d_score.device.index)
    Z, Y, X = d_score.shape
    chosen = torch.empty((Y, X), dtype=torch.int32, device=d_score.device)
    nws = int(lib.tsp_manifold_workspace_bytes())
    ws = torch.empty(nws, dtype=torch.uint8, device=d_score.device)
    rc = lib.tsp_build_manifold(h, C.c_void_p(d_score.data_ptr()), C.c_void_p(chosen.data_ptr()), Z, Y, X,
                                C.c_void_p(ws.data_ptr()), nws, _stream_ptr(stream))
    check(rc, "tsp_build_manifold")
    return chosen


def build_manifold(score, device=None):
    """Host convenience: numpy float32 (Z,Y,X) -> int64 (Y,X) like the reference's build_continues_manifold."""
    import torch
    handle(device)
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    d = torch.from_numpy(np.ascontiguousarray(score, dtype=np.float32)).to(dev)
    return build_manifold_device(d).cpu().numpy().astype(np.int64)


def block_reduce(d_volume, bin_size, variance=False, stream=None):
    import torch
    lib, h = load_library(), handle(d_volume.device.index)
    Z, Y, X = d_volume.shape
    out = torch.empty((Z, -(-Y // bin_size), -(-X // bin_size)), dtype=torch.float32, device=d_volume.device)
    rc = lib.tsp_block_reduce_f32(h, C.c_void_p(d_volume.data_ptr()), C.c_void_p(out.data_ptr()), Z, Y, X,
                                  int(bin_size), 1 if variance else 0, _stream_ptr(stream))
    check(rc, "tsp_block_reduce_f32")
    return out


def resize_argmax(d_score, rows, cols, z_offset=0, stream=None):
    import torch
    lib, h = load_library(), handle(d_score.device.index)
    Z, cy, cx = d_score.shape
    zmap = torch.empty((rows, cols), dtype=torch.int32, device=d_score.device)
    ws = torch.empty(256, dtype=torch.uint8, device=d_score.device)
    rc = lib.tsp_resize_argmax_f32(h, C.c_void_p(d_score.data_ptr()), C.c_void_p(zmap.data_ptr()), Z, int(rows),
                                   int(cols), cy, cx, int(z_offset), C.c_void_p(ws.data_ptr()), 256,
                                   _stream_ptr(stream))
    check(rc, "tsp_resize_argmax_f32")
    return zmap


def resize_round(d_coarse, rows, cols, stream=None):
    import torch
    lib, h = load_library(), handle(d_coarse.device.index)
    cy, cx = d_coarse.shape
    zmap = torch.empty((rows, cols), dtype=torch.int32, device=d_coarse.device)
    rc = lib.tsp_resize_round_i32(h, C.c_void_p(d_coarse.data_ptr()), C.c_void_p(zmap.data_ptr()), int(rows),
                                  int(cols), cy, cx, _stream_ptr(stream))
    check(rc, "tsp_resize_round_i32")
    return zmap


def band_project(d_stack, d_zmap, reference_channel=0, atoh_shift=0, airyscan=False, stream=None):
    """SP:62-81 from a given height map.  d_stack (C,Z,Y,X) uint16, d_zmap (Y,X) int32."""
    import torch
    lib, h = load_library(), handle(d_stack.device.index)
    Cn, Z, Y, X = d_stack.shape
    proj = torch.empty((Cn, Y, X), dtype=torch.float32, device=d_stack.device)
    nws = int(lib.tsp_band_workspace_bytes(Cn, Z, Y, X))
    ws = torch.empty(nws, dtype=torch.uint8, device=d_stack.device)
    rc = lib.tsp_band_project(h, C.c_void_p(d_stack.data_ptr()), C.c_void_p(d_zmap.data_ptr()),
                              C.c_void_p(proj.data_ptr()), Cn, Z, Y, X, int(reference_channel), int(atoh_shift),
                              1 if airyscan else 0, C.c_void_p(ws.data_ptr()), nws, _stream_ptr(stream))
    check(rc, "tsp_band_project")
    st = FrameStatus()
    rc = lib.tsp_get_frame_status(h, C.c_void_p(ws.data_ptr()), _stream_ptr(stream), C.byref(st))
    check(rc, "tsp_band_project")
    return proj


def project_m(d_channel, method, bin_size, stream=None):
    """SPM:14-35 on a (Z,Y,X) uint16 CUDA tensor (already channel-selected and z-cropped)."""
    import torch
    lib, h = load_library(), handle(d_channel.device.index)
    Z, Y, X = d_channel.shape
    out = torch.empty((Y, X), dtype=torch.uint16, device=d_channel.device)
    nws = int(lib.tsp_project_m_workspace_bytes(Z, Y, X, int(bin_size)))
    ws = torch.empty(max(nws, 256), dtype=torch.uint8, device=d_channel.device)
    rc = lib.tsp_project_m(h, C.c_void_p(d_channel.data_ptr()), C.c_void_p(out.data_ptr()), Z, Y, X,
                           int(method), int(bin_size), C.c_void_p(ws.data_ptr()), nws, _stream_ptr(stream))
    check(rc, "tsp_project_m")
    return out
