"""Stage-level GPU tests: every kernel of the path against the numpy/scipy statement of the same stage."""
import numpy as np
import pytest

from oracle import synth
from oracle import surface_projection_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    import torch
    assert torch.cuda.is_available()
    from tissue_image_processing_b200 import _native
    _native.load_library()
    return _native


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("sigma", [orc.SIGMA_PRE, orc.SIGMA_SCORE, orc.SIGMA_MASK, (0.5, 3.0, 7.5)])
@pytest.mark.parametrize("shape", [(9, 70, 65), (5, 300, 257), (1, 33, 1030), (3, 1, 17)])
def test_gaussian_blur_f32_bit_exact(nat, sigma, shape):
    rng = np.random.default_rng(sum(shape))
    vol = rng.integers(0, 4096, size=shape).astype(np.float32) + rng.random(shape).astype(np.float32)
    got = nat.gaussian_blur(_cuda(vol), sigma, fp64_accumulate=True).cpu().numpy()
    want = orc.blur_image(vol, sigma)
    assert np.array_equal(got, want)
    got32 = nat.gaussian_blur(_cuda(vol), sigma, fp64_accumulate=False).cpu().numpy()
    np.testing.assert_allclose(got32, want, rtol=3e-6, atol=0)


@pytest.mark.parametrize("sigma,shape", [((0.5, 30.0, 30.0), (3, 384, 448)), ((0.5, 30.0, 30.0), (2, 200, 192)),
                                         ((1.0, 5.0, 11.0), (2, 300, 320)), ((0.5, 2.5, 29.9), (1, 129, 1024))])
def test_gaussian_blur_f32_streaming_line_filters(nat, sigma, shape):
    """The fp32 (exact-mode) in-plane passes of radius 9 .. 120 at sizes where the streaming kernels apply (X % 64 == 0,
    Y >= 128): the 384-position shared-memory ring (full at radius 120, with slack below), lines longer and shorter
    than one ring, a last step that is only partly inside the image - against scipy's float64 accumulation."""
    rng = np.random.default_rng(sum(shape))
    vol = rng.integers(0, 4096, size=shape).astype(np.float32) + rng.random(shape).astype(np.float32)
    got32 = nat.gaussian_blur(_cuda(vol), sigma, fp64_accumulate=False).cpu().numpy()
    np.testing.assert_allclose(got32, orc.blur_image(vol, sigma), rtol=3e-6, atol=0)


@pytest.mark.parametrize("shape", [(12, 96, 112), (7, 53, 67)])
def test_gaussian_blur_u16_truncates_like_scipy(nat, shape):
    rng = np.random.default_rng(3)
    vol = rng.integers(0, 65535, size=shape).astype(np.uint16)
    got = nat.gaussian_blur(_cuda(vol), orc.SIGMA_M).cpu().numpy()
    assert np.array_equal(got, orc.blur_image(vol, orc.SIGMA_M))


def _np_p95(vol, airyscan):
    img = vol.astype(np.float32)
    if airyscan:
        img -= 10000
        img[img < 0] = 0
    nz = img[img > 0]
    return (None if nz.size == 0 else np.percentile(nz, 95)), nz.size


@pytest.mark.parametrize("kind", ["structured", "white", "sparse", "two_valued", "single", "zeros", "airyscan",
                                  "odd_offset"])
def test_percentile_matches_numpy(nat, kind):
    rng = np.random.default_rng(11)
    airy = kind == "airyscan"
    if kind == "structured":
        vol = synth.synth_stack(16, 200, 210, seed=3)[0]
    elif kind == "white":
        vol = rng.integers(0, 65536, size=(8, 128, 130)).astype(np.uint16)
    elif kind == "sparse":
        vol = synth.sparse_spike_stack(8, 100, 100, seed=2)[0]
    elif kind == "two_valued":
        vol = np.where(rng.random((4, 64, 64)) < 0.949, 100, 60000).astype(np.uint16)
    elif kind == "single":
        vol = np.zeros((2, 8, 8), dtype=np.uint16)
        vol[1, 3, 3] = 777
    elif kind == "zeros":
        vol = np.zeros((3, 16, 16), dtype=np.uint16)
    elif kind == "airyscan":
        vol = synth.synth_stack(8, 96, 96, seed=4, airyscan=True)[0]
    else:
        vol = rng.integers(0, 5000, size=(3, 37, 41)).astype(np.uint16)
    if kind == "odd_offset":
        flat = _cuda(np.concatenate([[0], vol.ravel()]).astype(np.uint16))[1:]      # 2-byte aligned start
        st = nat.percentile95_nonzero(flat.contiguous() if False else flat, airyscan=airy)
    else:
        st = nat.percentile95_nonzero(_cuda(vol), airyscan=airy)
    want, n = _np_p95(vol, airy)
    assert st["nonzero_count"] == n
    assert st["has_nonzero"] == (n > 0)
    if n:
        assert np.float32(st["percentile95"]) == np.float32(want)


def test_percentile_float32_rank_beyond_2_24(nat):
    """SURVEY trap T1 at n > 2**24: numpy quantises the rank in float32."""
    n = 20_000_003
    vol = np.full(n, 7, dtype=np.uint16)
    vol[19_000_002:] = 9            # float64 rank 19000001.9 would interpolate, float32 rank does not
    rng = np.random.default_rng(0)
    rng.shuffle(vol)
    st = nat.percentile95_nonzero(_cuda(vol))
    want = np.percentile(vol.astype(np.float32), 95)
    assert np.float32(st["percentile95"]) == np.float32(want) == np.float32(9.0)


@pytest.mark.parametrize("kind", ["white16", "white12", "top", "constant", "airyscan", "gamma", "sparse",
                                  "odd_offset", "window_edge"])
def test_percentile_sampled_window_path(nat, kind):
    """Volumes large enough (> 8 M voxels) for the sample -> value window -> streaming count pass; every
    distribution must give numpy's float32 percentile exactly (or fall back to the full histogram by itself)."""
    rng = np.random.default_rng(23)
    n = 12_000_011
    airy = kind == "airyscan"
    if kind == "white16":
        vol = rng.integers(0, 65536, size=n)
    elif kind == "white12":
        vol = rng.integers(0, 4096, size=n)
    elif kind == "top":
        vol = rng.integers(65000, 65536, size=n)            # the window saturates at 65535
    elif kind == "constant":
        vol = np.full(n, 1234)                              # every voxel lies inside the window
    elif kind == "airyscan":
        vol = rng.integers(9000, 14000, size=n)
    elif kind == "gamma":
        vol = np.minimum(rng.gamma(2.0, 300.0, size=n), 65535)
    elif kind == "sparse":
        vol = np.where(rng.random(n) < 0.02, rng.integers(1, 3000, size=n), 0)
    elif kind == "window_edge":
        vol = np.where(rng.random(n) < 0.9497, rng.integers(1000, 1024, size=n), rng.integers(1024, 1100, size=n))
    else:
        vol = rng.integers(0, 5000, size=n)
    vol = vol.astype(np.uint16)
    if kind == "odd_offset":
        d = _cuda(np.concatenate([[0], vol]).astype(np.uint16))[1:]                 # 2-byte aligned start
    else:
        d = _cuda(vol)
    st = nat.percentile95_nonzero(d, airyscan=airy)
    want, cnt = _np_p95(vol, airy)
    assert st["nonzero_count"] == cnt
    assert np.float32(st["percentile95"]) == np.float32(want)


def test_argmax_first_maximum_wins(nat):
    rng = np.random.default_rng(2)
    score = rng.integers(0, 4, size=(9, 40, 50)).astype(np.float32)       # many exact ties
    got = nat.argmax_z(_cuda(score), z_offset=3).cpu().numpy()
    assert np.array_equal(got, 3 + np.argmax(score, axis=0))


@pytest.mark.parametrize("shift,ref", [(0, 0), (2, 1), (-3, 0)])
def test_band_projection_from_oracle_height_map(nat, shift, ref):
    Z, Y, X, C = 14, 75, 99, 2
    img = synth.synth_stack(Z, Y, X, C=C, seed=8)
    rng = np.random.default_rng(4)
    zmap = np.clip((synth.height_field(Z, Y, X) + rng.integers(-1, 2, size=(Y, X))).astype(np.int64), 0, Z - 1)
    zmap[:6, :] = 0
    zmap[-6:, :] = Z - 1 - max(shift, 0)
    zmap = np.clip(zmap, 0, Z - 1 - max(shift, 0))
    image = img.astype(np.float32)
    z_other = zmap if shift == 0 else np.clip(zmap + shift, 0, Z)
    want = orc.project_channels(image, ref, orc.band_mask(zmap, Z), orc.band_mask(z_other, Z))
    got = nat.band_project(_cuda(img), _cuda(zmap.astype(np.int32)), reference_channel=ref, atoh_shift=shift)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=0)


@pytest.mark.parametrize("shape,C,shift,noisy", [((40, 96, 128), 1, 0, True), ((37, 70, 200), 2, 0, True),
                                                 ((21, 64, 64), 3, 2, False), ((30, 33, 72), 2, -2, True),
                                                 ((44, 160, 192), 1, 0, "gappy"), ((35, 96, 136), 2, 0, "gappy"),
                                                 ((26, 160, 256), 2, 0, "mixed"), ((19, 96, 192), 3, -1, "mixed"),
                                                 ((6, 64, 128), 1, 0, "mixed"), ((64, 256, 320), 1, 0, False),
                                                 ((40, 160, 256), 1, 0, "steep"), ((48, 192, 320), 2, 1, "steep")])
def test_band_projection_tma_ring(nat, shape, C, shift, noisy):
    """TMA path of the band stage (X % 8 == 0, tile inside the image): shallow tiles in the register kernel, plane
    ranges deeper than the ring (a noisy height map walks every plane, refilling the ring) through the worklist, one /
    two / three channels, shifted masks; against the oracle and bit for bit against the older kernels
    (tsp_debug_set "band_variant" 3 and 2), which do the same arithmetic in the same order."""
    Z, Y, X = shape
    img = synth.synth_stack(Z, Y, X, C=C, seed=sum(shape))
    if img.ndim == 3:
        img = img[None]
    rng = np.random.default_rng(7)
    hi = Z - 1 - max(shift, 0)
    if noisy == "gappy":           # few distinct heights: most planes of a tile's range are absent (skipped sweeps)
        levels = np.array([1, 2, 9, 10, 11, 23, hi - 3, hi])
        zmap = levels[rng.integers(0, levels.size, size=(Y, X))].astype(np.int64)
        zmap[: Y // 2, : X // 2] = 12
    elif noisy == "mixed":         # shallow tiles (register kernel) next to deep ones (worklist), surfaces at both stack ends
        zmap = np.clip(synth.height_field(Z, Y, X).astype(np.int64), 0, hi)
        zmap[:40, :70] = 0
        zmap[:40, 70:140] = hi
        zmap[40:56, :64] = 1
        zmap[-32:, -64:] = rng.integers(0, hi + 1, size=(32, 64))
        zmap[Y // 2, X // 2] = min(hi, zmap[Y // 2, X // 2] + 3)
    elif noisy == "steep":         # slope growing with x: tiles spanning 1 .. 5 planes (register kernel) next to
        yy, xx = np.mgrid[0:Y, 0:X]  # tiles of 6 .. 16 planes (worklist)
        zmap = np.clip(np.floor(yy * (0.02 + 0.2 * xx / X)).astype(np.int64), 0, hi)
    elif noisy:
        zmap = rng.integers(0, hi + 1, size=(Y, X)).astype(np.int64)
    else:
        zmap = np.clip(synth.height_field(Z, Y, X).astype(np.int64), 0, hi)
    image = img.astype(np.float32)
    z_other = zmap if shift == 0 else np.clip(zmap + shift, 0, Z)
    want = orc.project_channels(image, 0, orc.band_mask(zmap, Z), orc.band_mask(z_other, Z))
    got = nat.band_project(_cuda(img), _cuda(zmap.astype(np.int32)), reference_channel=0, atoh_shift=shift).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)
    try:
        for variant in (3, 2):                          # generic TMA kernel alone; register-prefetch kernel
            nat.debug_set("band_variant", variant)
            alt = nat.band_project(_cuda(img), _cuda(zmap.astype(np.int32)), reference_channel=0,
                                   atoh_shift=shift).cpu().numpy()
            assert np.array_equal(got, alt), variant
    finally:
        nat.debug_set("band_variant", 0)


def test_band_projection_index_error(nat):
    img = synth.synth_stack(6, 20, 20, seed=1)
    zmap = np.full((20, 20), 6, dtype=np.int32)
    with pytest.raises(IndexError):
        nat.band_project(_cuda(img), _cuda(zmap))


def test_host_c_abi_call_matches_device_call(nat):
    import torch
    img = synth.synth_stack(10, 64, 96, C=2, seed=6)
    proj, zmap, st = nat.project_frame_host(img, 0, mode="exact")
    p = nat.DeviceProjector(2, 10, 64, 96, mode="exact")
    dproj, dzmap = p.run(_cuda(img))
    torch.cuda.synchronize()
    assert np.array_equal(proj, dproj.cpu().numpy().astype(np.float64))
    assert np.array_equal(zmap, dzmap.cpu().numpy().astype(np.int64))
    assert st["has_nonzero"] and not st["band_index_error"]
    assert p.status()["percentile95"] == st["percentile95"]


@pytest.mark.parametrize("shape", [(24, 520, 776), (9, 131, 264), (70, 300, 1032), (5, 77, 2056)])
def test_interpolation_stage_variants_agree(nat, shape):
    """The interpolation + argmax stage staged in shared memory (default) and reading its control points from L2
    (round-1 kernel, tsp_debug_set "interp_global") run the same arithmetic: identical height maps and projections,
    also with plain instead of graph-replayed launches."""
    import torch
    Z, Y, X = shape
    stack = torch.from_numpy(synth.synth_stack(Z, Y, X, seed=sum(shape))).cuda()
    p = nat.DeviceProjector(1, Z, Y, X, mode="fast")
    outs = []
    try:
        for key, val in ((None, 0), ("interp_global", 1), ("graphs", 0)):
            if key:
                nat.debug_set(key, val)
            for _ in range(3):                       # first call plain, second captured, third replayed
                proj, zmap = p.run(stack)
            torch.cuda.synchronize()
            outs.append((proj.clone(), zmap.clone()))
            if key:
                nat.debug_set(key, 1 - val)
    finally:
        nat.debug_set("interp_global", 0)
        nat.debug_set("graphs", 1)
    for proj, zmap in outs[1:]:
        assert torch.equal(zmap, outs[0][1]) and torch.equal(proj, outs[0][0])
