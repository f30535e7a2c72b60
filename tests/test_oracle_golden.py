"""The oracle restatement against the golden outputs frozen from the real reference
(reference surface_projection.py:17-85, surface_proj_m.py:14-35).  CPU only."""
import contextlib
import io

import numpy as np
import pytest

from oracle import golden_cases
from oracle import surface_projection_oracle as orc


@pytest.mark.parametrize("case", golden_cases.CASES, ids=lambda c: c[0])
def test_operator_matches_reference_bit_for_bit(case, golden):
    name, build, axes, kw = case
    manifest, arrays = golden
    res = orc.time_point_surface_projection(build(), axes, **kw)
    proj, zmap = res if kw.get("z_map") else (res, None)
    assert proj.dtype == np.float64
    assert np.array_equal(proj, arrays[name + "/projection"].astype(np.float64))
    if zmap is not None:
        assert str(zmap.dtype) == manifest["cases"][name]["zmap_dtype"]
        assert np.array_equal(zmap, arrays[name + "/zmap"])


@pytest.mark.parametrize("case", golden_cases.ERROR_CASES, ids=lambda c: c[0])
def test_operator_raises_like_reference(case, golden):
    name, build, axes, kw = case
    manifest, _ = golden
    with pytest.raises(Exception) as info:
        orc.time_point_surface_projection(build(), axes, **kw)
    assert type(info.value).__name__ == manifest["cases"][name]["raises"]


@pytest.mark.parametrize("case", golden_cases.SPM_CASES, ids=lambda c: c[0])
def test_surface_projection_m_matches_reference(case, golden):
    name, build, axes, kw = case
    _, arrays = golden
    with contextlib.redirect_stdout(io.StringIO()):
        out = orc.surface_projection_m(build(), axes, **kw)
    assert out.dtype == np.uint16
    assert np.array_equal(out, arrays[name + "/projection"])


def test_scipy_pass_restatement_is_bit_exact():
    """SURVEY trap T8: float64 accumulate per line, one cast per pass, z -> y -> x."""
    rng = np.random.default_rng(0)
    vol = rng.integers(0, 4096, size=(9, 70, 65)).astype(np.float32)
    for sig in (orc.SIGMA_PRE, orc.SIGMA_SCORE, orc.SIGMA_MASK):
        assert np.array_equal(orc.gaussian_filter_restated(vol, sig), orc.blur_image(vol, sig))
    u16 = rng.integers(0, 60000, size=(9, 70, 65)).astype(np.uint16)
    assert np.array_equal(orc.gaussian_filter_restated(u16, orc.SIGMA_M), orc.blur_image(u16, orc.SIGMA_M))


def test_band_mask_closed_form_is_bit_exact():
    """SURVEY trap T9 incl. z edge replication (surface at plane 0 and Z-1)."""
    rng = np.random.default_rng(1)
    for Z in (1, 3, 12):
        cz = rng.integers(0, Z, size=(37, 41))
        cz[:5] = 0
        cz[-5:] = Z - 1
        assert np.array_equal(orc.band_mask(cz, Z), orc.band_mask_closed_form(cz, Z))


def _percentile_f32_restated(sorted_vals, q=95):
    """numpy/lib/_function_base_impl.py (2.3.5) 'linear' method, every step in float32."""
    n = sorted_vals.size
    qf = np.float32(q) / np.float32(100)
    vi = np.float32(np.float32(n - 1) * qf)
    if vi >= np.float32(n - 1):
        return sorted_vals[-1]
    lo = np.floor(vi)
    hi = np.float32(lo + np.float32(1))
    g = np.float32(vi - lo)
    a, b = sorted_vals[int(lo)], sorted_vals[int(hi)]
    d = np.float32(b - a)
    if g >= 0.5:
        return np.float32(b - np.float32(d * np.float32(np.float32(1) - g)))
    return np.float32(a + np.float32(d * g))


def test_percentile_rank_is_float32():
    """SURVEY trap T1: numpy evaluates the virtual index (n-1)*0.95 in float32 for float32 data,
    so above 2**24 elements the rank is quantised."""
    n = 20_000_003
    step = 19_000_002                       # float64 rank is 19000001.9, float32 rank 19000002
    data = np.zeros(n, dtype=np.float32)
    data[step:] = 1
    got = np.percentile(data, 95)
    assert got == _percentile_f32_restated(data)
    assert got == np.float32(1.0)
    assert np.percentile(data.astype(np.float64), 95) == pytest.approx(0.9)


def test_percentile_restatement_small_counts():
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 20, 21, 1000, 4097):
        v = np.sort(rng.integers(1, 50, size=n).astype(np.float32))
        assert np.percentile(v, 95) == _percentile_f32_restated(v), n


def test_large_golden_fixtures_are_complete():
    """Every case of oracle/large_cases.py has its frozen reference output (digests, height map, near-tie mask, rows)."""
    import json
    import os
    from oracle import large_cases
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "large")
    for name, _, kw in large_cases.LARGE_CASES:
        with open(os.path.join(root, name + ".json")) as f:
            meta = json.load(f)
        arr = np.load(os.path.join(root, name + ".npz"))
        Y, X = arr["zmap"].shape
        assert meta["shape"][-2:] == [Y, X] and meta["kwargs"] == kw
        assert arr["near_tie_bits"].size == (Y * X + 7) // 8
        assert arr["proj_rows"].shape == (meta["shape"][1], arr["proj_row_index"].size, X)
        assert len(meta["zmap_sha256"]) == 64 and len(meta["proj_sha256"]) == 64


@pytest.mark.parametrize("name", ["cfg1_512x512x32", "zcrop_min3_768x768x24"])
def test_oracle_matches_large_reference_golden(name):
    """The oracle against the frozen reference run at BASELINE configs[0] and on the min_z > 0 corner (trap T4: the
    band sits min_z planes too deep, no IndexError)."""
    import hashlib
    import json
    import os
    from oracle import large_cases
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "large")
    case = {c[0]: c for c in large_cases.LARGE_CASES}[name]
    with open(os.path.join(root, name + ".json")) as f:
        meta = json.load(f)
    chunk = case[1]()
    assert hashlib.sha256(chunk.tobytes()).hexdigest() == meta["input_sha256"]
    proj, zmap = orc.time_point_surface_projection(chunk, "TCZYX", **case[2])
    assert hashlib.sha256(zmap.tobytes()).hexdigest() == meta["zmap_sha256"]
    assert hashlib.sha256(proj.astype(np.float32).tobytes()).hexdigest() == meta["proj_sha256"]
