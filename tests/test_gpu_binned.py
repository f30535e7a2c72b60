"""GPU tests of the binned score methods (SP:39-53, 59-65) and the continuous manifold (SP:87-165): every building
block against the numpy/scipy statement of the same step, then the operator against the oracle and against the
golden output of the real reference (the `manifold` case of tests/golden)."""
import numpy as np
import pytest

from oracle import golden_cases, synth
from oracle import surface_projection_oracle as orc
from tests.parity import compare_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    import torch
    assert torch.cuda.is_available()
    from tissue_image_processing_b200 import _native
    _native.load_library()
    return _native


@pytest.fixture(scope="module")
def tsp(nat):
    import tissue_image_processing_b200 as pkg
    return pkg


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape,b", [((3, 40, 44), 2), ((2, 41, 47), 3), ((4, 17, 9), 5), ((1, 64, 64), 8), ((2, 5, 7), 16)])
@pytest.mark.parametrize("variance", [False, True])
def test_block_reduce_matches_numpy(nat, shape, b, variance):
    rng = np.random.default_rng(sum(shape) + b)
    vol = (rng.random(shape) * 3000).astype(np.float32)
    want = orc.block_reduce(vol, (1, b, b), np.var if variance else np.mean)
    got = nat.block_reduce(_cuda(vol), b, variance).cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, want, rtol=3e-6, atol=0)       # float32 pairwise sums vs float64 sums


@pytest.mark.parametrize("shape,out", [((3, 5, 7), (10, 14)), ((2, 4, 6), (11, 17)), ((4, 33, 21), (65, 41)),
                                       ((1, 1, 9), (2, 18)), ((2, 7, 7), (7, 13)), ((5, 20, 30), (77, 119)),
                                       ((6, 64, 64), (128, 128)), ((3, 43, 86), (128, 257))])
def test_resize_argmax_matches_scipy_zoom(nat, shape, out):
    rng = np.random.default_rng(sum(shape))
    score = (rng.random(shape) * 1000).astype(np.float32)
    full = orc.resize_order1(score, (shape[0],) + out)
    want = 3 + np.argmax(full, axis=0)
    got = nat.resize_argmax(_cuda(score), out[0], out[1], z_offset=3).cpu().numpy()
    diff = got != want
    if diff.any():                       # only where the resampled top two are within float32 rounding
        assert (orc.top2_relative_gap(full)[diff] < 1e-6).all()
    assert diff.mean() < 1e-3


@pytest.mark.parametrize("shape,out", [((5, 7), (10, 14)), ((33, 21), (65, 41)), ((64, 64), (128, 128)), ((1, 9), (3, 27))])
def test_resize_round_matches_scipy_zoom(nat, shape, out):
    rng = np.random.default_rng(sum(shape))
    coarse = rng.integers(0, 40, size=shape).astype(np.int32)
    want = np.round(orc.resize_order1(coarse.astype("float32"), out)).astype("int")
    got = nat.resize_round(_cuda(coarse), out[0], out[1]).cpu().numpy()
    assert np.array_equal(got, want)


def _score_with_peak(shape, where, seed, smooth):
    rng = np.random.default_rng(seed)
    P, R, C = shape
    if smooth:
        zz = np.arange(P, dtype=np.float32)[:, None, None]
        h = synth.height_field(P, R, C)[None]
        score = (1000 * np.exp(-(zz - h) ** 2 / 8) + rng.random(shape) * 30).astype(np.float32)
    else:
        score = (rng.random(shape) * 100).astype(np.float32)
    r, c = where
    score[rng.integers(0, P), r % R, c % C] = 5000.0
    return score


@pytest.mark.parametrize("shape,where,smooth", [
    ((7, 23, 31), (11, 15), False), ((7, 23, 31), (0, 0), False), ((7, 23, 31), (22, 30), False),
    ((5, 16, 16), (0, 15), False), ((9, 30, 12), (29, 0), True), ((4, 1, 40), (0, 17), False),
    ((4, 40, 1), (20, 0), False), ((3, 9, 9), (4, 4), False), ((12, 64, 80), (31, 40), True),
    ((2, 33, 35), (16, 17), False), ((1, 12, 12), (3, 9), False), ((20, 150, 170), (10, 160), True),
    # edges longer than one thread walks (composition of the transition tables over 32 runs): random and smooth
    # scores, start in a corner / on an edge / inside, and more planes than one lane round (P > 32, P > 64)
    ((6, 300, 420), (150, 200), False), ((40, 260, 300), (0, 0), True), ((200, 130, 140), (129, 70), False),
    ((70, 200, 333), (100, 332), True)])
def test_manifold_equals_reference_algorithm(nat, shape, where, smooth):
    """Integer result of a sequential algorithm: must be identical, including the wrap of row -1 (start rows that
    make the left edge span the whole height) and the truncated mean of two-apart neighbours."""
    score = _score_with_peak(shape, where, sum(shape), smooth)
    want = orc.build_continues_manifold(score)
    got = nat.build_manifold(score)
    assert got.dtype == np.int64 and got.shape == want.shape
    assert np.array_equal(got, want), "differs at %d of %d cells" % ((got != want).sum(), want.size)


def test_manifold_with_exact_ties(nat):
    rng = np.random.default_rng(5)
    score = rng.integers(0, 3, size=(6, 28, 26)).astype(np.float32)         # ties everywhere: first maximum wins
    assert np.array_equal(nat.build_manifold(score), orc.build_continues_manifold(score))


def test_manifold_golden_case(golden, tsp):
    """The `manifold` case frozen from the real reference (oracle/make_golden.py)."""
    case = [c for c in golden_cases.CASES if c[3].get("build_manifold")][0]
    name, build, axes, kw = case
    _, arrays = golden
    got_proj, got_zmap = tsp.time_point_surface_projection(build(), axes, mode="bitexact", **kw)
    assert np.array_equal(got_zmap, arrays[name + "/zmap"])
    assert np.array_equal(got_proj, arrays[name + "/projection"].astype(np.float64))
    alt_proj, alt_zmap = tsp.time_point_surface_projection(build(), axes, mode="fast", **kw)
    assert (alt_zmap != got_zmap).mean() < 0.01


@pytest.mark.parametrize("mode", ["bitexact", "exact", "fast"])
@pytest.mark.parametrize("method,b,C,shift", [("max_averages", 2, 1, 0), ("max_averages", 3, 2, 2), ("max_std", 2, 1, 0),
                                             ("max_std", 4, 2, -1), ("multi_channel", 2, 2, 0),
                                             ("multi_channel", 3, 3, 1), ("multi_channel", 3, 3, -2)])
def test_binned_methods_against_oracle(tsp, mode, method, b, C, shift):
    Z, Y, X = 14, 70, 94
    img = synth.synth_stack(Z, Y, X, C=C, seed=40 + b)
    if img.ndim == 3:
        img = img[None]
    img = img[None]
    kw = dict(reference_channel=0, airyscan=False, z_map=True, method=method, bin_size=b, atoh_shift=shift)
    try:
        (want_proj, want_zmap), score = orc.time_point_surface_projection(img, "TCZYX", return_score=True, **kw)
    except IndexError:                     # the shifted height map leaves the stack (SP:68-69): same error here
        with pytest.raises(IndexError):
            tsp.time_point_surface_projection(img, "TCZYX", mode=mode, **kw)
        return
    got_proj, got_zmap = tsp.time_point_surface_projection(img, "TCZYX", mode=mode, **kw)
    assert got_proj.dtype == np.float64 and got_zmap.dtype == np.int64
    stats = compare_frame(got_proj, got_zmap, want_proj, want_zmap, orc.top2_relative_gap(score))
    print(method, b, mode, stats)


def test_binned_airyscan_zcrop(tsp):
    img = synth.synth_stack(18, 64, 80, C=2, seed=9, airyscan=True)[None]
    kw = dict(reference_channel=1, airyscan=True, z_map=True, method="max_averages", bin_size=2, min_z=2, max_z=16)
    with pytest.raises(IndexError):        # min_z is added to the height map but the band indexes the cropped stack
        orc.time_point_surface_projection(img, "TCZYX", **dict(kw, min_z=9, max_z=18))
    with pytest.raises(IndexError):
        tsp.time_point_surface_projection(img, "TCZYX", mode="exact", **dict(kw, min_z=9, max_z=18))
    kw = dict(kw, min_z=0, max_z=15)
    (want_proj, want_zmap), score = orc.time_point_surface_projection(img, "TCZYX", return_score=True, **kw)
    got_proj, got_zmap = tsp.time_point_surface_projection(img, "TCZYX", mode="exact", **kw)
    print(compare_frame(got_proj, got_zmap, want_proj, want_zmap, orc.top2_relative_gap(score)))


def test_unknown_method_raises_like_the_reference(tsp):
    img = synth.synth_stack(6, 32, 32, seed=1)
    img = img.reshape((1, 1) + img.shape[-3:])
    with pytest.raises(TypeError):
        tsp.time_point_surface_projection(img, "TCZYX", 0, airyscan=False, method="nope", bin_size=2)
    # bin_size == 1 never looks at the method (SP:39)
    tsp.time_point_surface_projection(img, "TCZYX", 0, airyscan=False, method="nope", bin_size=1)


@pytest.mark.parametrize("b,shift", [(1, 0), (1, 2), (2, 0), (3, -1)])
def test_manifold_operator_against_oracle(tsp, b, shift):
    """bitexact scores -> the region growing sees the reference's numbers: identical height map (bin_size 1); with
    bins the block means differ in the last float32 bit, so a few near-tie cells may grow differently."""
    Z, Y, X, C = 10, 48, 60, 2
    img = synth.synth_stack(Z, Y, X, C=C, seed=77)[None]
    kw = dict(reference_channel=0, airyscan=False, z_map=True, build_manifold=True, bin_size=b, atoh_shift=shift,
              method="max_averages")
    want_proj, want_zmap = orc.time_point_surface_projection(img, "TCZYX", **kw)
    got_proj, got_zmap = tsp.time_point_surface_projection(img, "TCZYX", mode="bitexact", **kw)
    mismatch = (got_zmap != want_zmap).mean()
    print("manifold operator", b, shift, mismatch)
    if b == 1:
        assert mismatch == 0
        assert np.array_equal(got_proj, want_proj)
    else:
        assert mismatch < 0.02
        if mismatch == 0:
            np.testing.assert_allclose(got_proj, want_proj, rtol=1e-5, atol=0)
