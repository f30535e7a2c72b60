"""CPU-side checks of the C ABI: the library loads without a GPU and exports every symbol that
include/tsp_b200.h declares; struct layouts agree with the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as entry
    entry.build()
    from tissue_image_processing_b200 import _native
    return _native


def _header_functions():
    text = open(os.path.join(ROOT, "include", "tsp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsp_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(native):
    lib = native.load_library()
    declared = _header_functions()
    assert "tsp_project_frame_host" in declared and len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libtsp_b200.so does not export " + name
    assert sorted(native.EXPORTS) == declared, "python binding table and header disagree"


def test_abi_version_and_struct_sizes(native):
    lib = native.load_library()
    assert lib.tsp_abi_version() == 3 == native.ABI_VERSION
    assert ctypes.sizeof(native.Params) == 16 * 4 and native.Params.sigma_mask.offset == 32
    assert ctypes.sizeof(native.FrameDesc) == 32 * 4 and native.FrameDesc.params.offset == 64
    assert ctypes.sizeof(native.FrameStatus) == 5 * 4 + 4 + 8 + 4 + 5 * 4   # with alignment padding
    assert native.FrameStatus.nonzero_count.offset == 24
    assert native.FrameDesc.bin_size.offset == 40 and native.FrameDesc.build_manifold.offset == 48


def test_binned_workspace_query(native):
    lib = native.load_library()
    plain = native.make_desc(2, 16, 128, 128, mode="fast")
    binned = native.make_desc(2, 16, 128, 128, mode="fast", bin_size=2, method="multi_channel")
    manifold = native.make_desc(2, 16, 128, 128, mode="fast", build_manifold=True)
    n0 = lib.tsp_project_workspace_bytes(ctypes.byref(plain))
    n1 = lib.tsp_project_workspace_bytes(ctypes.byref(binned))
    n2 = lib.tsp_project_workspace_bytes(ctypes.byref(manifold))
    assert n1 >= 2 * 16 * 128 * 128 * 4 + 16 * 64 * 64 * 4 > n0 and n2 >= 2 * 16 * 128 * 128 * 4
    bad = native.make_desc(1, 16, 128, 128, bin_size=2, method=7)
    assert lib.tsp_project_workspace_bytes(ctypes.byref(bad)) == 0 and b"method" in lib.tsp_last_error()


def test_workspace_query_needs_no_gpu(native):
    lib = native.load_library()
    d = native.make_desc(1, 32, 512, 512, mode="exact")
    n = lib.tsp_project_workspace_bytes(ctypes.byref(d))
    assert n >= 2 * 32 * 512 * 512 * 4
    bad = native.make_desc(1, 32, 512, 512, reference_channel=3, mode="exact")
    assert lib.tsp_project_workspace_bytes(ctypes.byref(bad)) == 0
    assert b"reference_channel" in lib.tsp_last_error()


def test_no_cpu_fallback_without_gpu(native):
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tissue_image_processing_b200 as pkg
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.time_point_surface_projection(np.zeros((1, 1, 4, 8, 8), np.uint16), "TCZYX", 0)


def test_descriptor_flags(native):
    """desc.flags: 0, TSP_FRAME_CONCURRENT and TSP_FRAME_OUT_U16 are valid, anything else (or a non-zero reserved
    word) is refused, and the flags do not change the workspace."""
    lib = native.load_library()
    plain = native.make_desc(1, 16, 128, 128, mode="fast")
    conc = native.make_desc(1, 16, 128, 128, mode="fast", concurrent=True)
    u16 = native.make_desc(1, 16, 128, 128, mode="fast", out_u16=True)
    assert plain.flags == 0 and conc.flags == native.FRAME_CONCURRENT == 1 and u16.flags == native.FRAME_OUT_U16 == 2
    n0 = lib.tsp_project_workspace_bytes(ctypes.byref(plain))
    assert n0 > 0 and lib.tsp_project_workspace_bytes(ctypes.byref(conc)) == n0
    assert lib.tsp_project_workspace_bytes(ctypes.byref(u16)) == n0
    bad = native.make_desc(1, 16, 128, 128, mode="fast")
    bad.flags = 4
    assert lib.tsp_project_workspace_bytes(ctypes.byref(bad)) == 0 and b"flags" in lib.tsp_last_error()
    bad.flags = 0
    bad.reserved = 1
    assert lib.tsp_project_workspace_bytes(ctypes.byref(bad)) == 0


def test_params_block(native):
    """tsp_params: defaults are the reference's constants; default values given explicitly keep the fast-mode
    workspace; non-default sigmas route to the direct FIR (two float volumes); out-of-range values are refused."""
    lib = native.load_library()
    assert native.default_params() == {"percentile": 95.0, "pedestal": 10000, "sigma_pre": (0.5, 1.0, 1.0),
                                       "sigma_score": (0.5, 30.0, 30.0), "sigma_mask": (1.0, 2.0, 2.0)}
    plain = native.make_desc(1, 16, 256, 256, mode="fast")
    assert plain.has_params == 0
    n0 = lib.tsp_project_workspace_bytes(ctypes.byref(plain))
    same = native.make_desc(1, 16, 256, 256, mode="fast", params=dict(percentile=95, sigma_mask=(1, 2, 2)))
    assert same.has_params == 1 and lib.tsp_project_workspace_bytes(ctypes.byref(same)) == n0
    pct = native.make_desc(1, 16, 256, 256, mode="fast", params=dict(percentile=99.5, pedestal=500))
    assert lib.tsp_project_workspace_bytes(ctypes.byref(pct)) == n0          # still the fast score stage
    vols = 2 * 16 * 256 * 256 * 4
    wide = native.make_desc(1, 16, 256, 256, mode="fast", params=dict(sigma_mask=(3, 2, 2)))
    assert lib.tsp_project_workspace_bytes(ctypes.byref(wide)) >= n0 + vols   # fast score + materialised band mask
    sig = native.make_desc(1, 16, 256, 256, mode="fast", params=dict(sigma_score=(0.5, 20, 20)))
    assert vols <= lib.tsp_project_workspace_bytes(ctypes.byref(sig)) < n0 + vols
    for bad in (dict(percentile=101), dict(pedestal=-1), dict(sigma_pre=(-1, 1, 1)), dict(sigma_mask=(1, 2, 1e4))):
        d = native.make_desc(1, 16, 256, 256, params=bad)
        assert lib.tsp_project_workspace_bytes(ctypes.byref(d)) == 0, bad
    with pytest.raises(TypeError):
        native.make_desc(1, 16, 256, 256, params=dict(sigma=3))
