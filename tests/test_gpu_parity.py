"""GPU parity tests proper: the CUDA path (through the Python mirror and the C ABI) against the oracle /
the golden outputs of the real reference.  Run on the B200 box: pytest -m gpu."""
import contextlib
import io

import numpy as np
import pytest

from oracle import golden_cases, synth
from oracle import surface_projection_oracle as orc
from tests.parity import compare_frame

pytestmark = pytest.mark.gpu

MODES = ["bitexact", "exact", "fast"]


@pytest.fixture(scope="module")
def tsp():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import tissue_image_processing_b200 as pkg
    from tissue_image_processing_b200 import _native
    _native.load_library()          # fails loudly if libtsp_b200.so is missing
    return pkg


def _oracle_gap(build, axes, kw):
    kw = dict(kw)
    kw["z_map"] = True
    (_, _), score = orc.time_point_surface_projection(build(), axes, return_score=True, **kw)
    return orc.top2_relative_gap(score)


SUPPORTED = [c for c in golden_cases.CASES if not c[3].get("build_manifold")]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", SUPPORTED, ids=lambda c: c[0])
def test_golden_cases(case, mode, golden, tsp):
    name, build, axes, kw = case
    _, arrays = golden
    want_proj = arrays[name + "/projection"].astype(np.float64)
    res = tsp.time_point_surface_projection(build(), axes, mode=mode, **kw)
    if not kw.get("z_map"):
        assert isinstance(res, np.ndarray) and res.dtype == np.float64
        kw2 = dict(kw, z_map=True)
        got_proj, got_zmap = tsp.time_point_surface_projection(build(), axes, mode=mode, **kw2)
        assert np.array_equal(res, got_proj)
        (_, want_zmap) = orc.time_point_surface_projection(build(), axes, **kw2)
    else:
        got_proj, got_zmap = res
        want_zmap = arrays[name + "/zmap"]
    assert got_proj.dtype == np.float64 and got_zmap.dtype == np.int64
    gap = _oracle_gap(build, axes, kw)
    stats = compare_frame(got_proj, got_zmap, want_proj, want_zmap, gap, exact_zmap=(mode == "bitexact"))
    if mode == "bitexact":
        assert np.array_equal(got_proj, want_proj), "bit-exact mode must reproduce the projection exactly"
    print(name, mode, stats)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", golden_cases.ERROR_CASES, ids=lambda c: c[0])
def test_error_cases(case, mode, golden, tsp):
    name, build, axes, kw = case
    manifest, _ = golden
    with pytest.raises(Exception) as info:
        tsp.time_point_surface_projection(build(), axes, mode=mode, **kw)
    assert type(info.value).__name__ == manifest["cases"][name]["raises"]


@pytest.mark.parametrize("case", golden_cases.SPM_CASES, ids=lambda c: c[0])
def test_surface_projection_m_golden(case, golden, tsp):
    name, build, axes, kw = case
    _, arrays = golden
    out = tsp.surface_projection_m(build(), axes, **kw)
    assert out.dtype == np.uint16
    assert np.array_equal(out, arrays[name + "/projection"])


def test_surface_projection_m_choose_limit(tsp):
    a = np.zeros((1, 64, 8, 8), dtype=np.uint16)
    with pytest.raises(ValueError):
        tsp.surface_projection_m(a, "CZYX", 0, 0, 64, "max_averages", 2)


@pytest.mark.parametrize("mode", MODES)
def test_config1_full_size(mode, tsp):
    """BASELINE config 1: 512x512x32 single channel, checked against the oracle run on this box."""
    img = synth.synth_stack(32, 512, 512, seed=1)[None]
    kw = dict(reference_channel=0, airyscan=False, z_map=True)
    (want_proj, want_zmap), score = orc.time_point_surface_projection(img, "TCZYX", return_score=True, **kw)
    got_proj, got_zmap = tsp.time_point_surface_projection(img, "TCZYX", mode=mode, **kw)
    stats = compare_frame(got_proj, got_zmap, want_proj, want_zmap, orc.top2_relative_gap(score),
                          exact_zmap=(mode == "bitexact"))
    print("config1", mode, stats)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_white_noise_tolerance_rule(mode, tsp):
    """Near-ties everywhere: differences are only allowed where the oracle gap is below 1e-4."""
    img = synth.white_noise_stack(24, 320, 352, seed=5)[None]
    kw = dict(reference_channel=0, airyscan=False, z_map=True)
    (want_proj, want_zmap), score = orc.time_point_surface_projection(img, "TCZYX", return_score=True, **kw)
    got_proj, got_zmap = tsp.time_point_surface_projection(img, "TCZYX", mode=mode, **kw)
    stats = compare_frame(got_proj, got_zmap, want_proj, want_zmap, orc.top2_relative_gap(score))
    print("white", mode, stats)


@pytest.mark.parametrize("shape,airy", [((12, 264, 520), True), ((9, 300, 1000), False), ((6, 131, 256), False),
                                        ((5, 77, 2056), True)])
def test_fast_mode_tma_ring_shapes(shape, airy, tsp):
    """The TMA-ring decimation (X >= 256, X % 8 == 0) at ragged chunk widths, short images and with the
    airyscan pedestal, against the oracle; the strip kernel (tsp_debug_set "no_ring") must satisfy the same rules."""
    from tissue_image_processing_b200 import _native
    Z, Y, X = shape
    img = synth.synth_stack(Z, Y, X, seed=sum(shape), airyscan=airy)[None]
    kw = dict(reference_channel=0, airyscan=airy, z_map=True)
    (want_proj, want_zmap), score = orc.time_point_surface_projection(img, "TCZYX", return_score=True, **kw)
    gap = orc.top2_relative_gap(score)
    got_proj, got_zmap = tsp.time_point_surface_projection(img, "TCZYX", mode="fast", **kw)
    print("ring", shape, compare_frame(got_proj, got_zmap, want_proj, want_zmap, gap))
    _native.debug_set("no_ring", 1)
    try:
        alt_proj, alt_zmap = tsp.time_point_surface_projection(img, "TCZYX", mode="fast", **kw)
    finally:
        _native.debug_set("no_ring", 0)
    print("strip", shape, compare_frame(alt_proj, alt_zmap, want_proj, want_zmap, gap))
    assert (alt_zmap != got_zmap).mean() < 1e-3
