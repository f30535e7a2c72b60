"""Parity at the sizes bench.py quotes: the five BASELINE configs (and three stress / corner inputs) against outputs
of the UNMODIFIED reference frozen by oracle/make_golden_large.py (reference surface_projection.py:17-85 run once in
the build container on the same seeded numpy inputs).

  bitexact   sha256 of the whole int64 height map and of the whole float32 projection must equal the reference's
             digests - every pixel of the frame is pinned to the reference, not to another mode of this library;
  exact/fast north-star rule (tests/parity.compare_frame) against the reference's height map with the ORACLE's
             near-tie mask, and against the reference's stored projection rows (every 8th row + the borders).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import large_cases
from scipy.ndimage import maximum_filter

from tests.parity import PROJ_RTOL, compare_frame

pytestmark = pytest.mark.gpu

LARGE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "large")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _load(name):
    with open(os.path.join(LARGE_DIR, name + ".json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(LARGE_DIR, name + ".npz"))


@pytest.fixture(scope="module")
def tsp():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import tissue_image_processing_b200 as pkg
    from tissue_image_processing_b200 import _native
    _native.load_library()
    return pkg


@pytest.mark.parametrize("case", large_cases.LARGE_CASES, ids=lambda c: c[0])
def test_reference_golden_at_bench_size(case, tsp):
    name, build, kw = case
    meta, arr = _load(name)
    chunk = build()
    assert list(chunk.shape) == meta["shape"]
    assert _sha(chunk) == meta["input_sha256"], "the seeded numpy generator did not reproduce the frozen input"
    want_zmap = arr["zmap"].astype(np.int64)
    Y, X = want_zmap.shape
    near_tie = np.unpackbits(arr["near_tie_bits"], count=Y * X).reshape(Y, X).astype(bool)
    gap = np.where(near_tie, 0.0, 1.0)
    gap.ravel()[arr["low_gap_index"]] = arr["low_gap_value"]
    rows = arr["proj_row_index"]
    want_rows = arr["proj_rows"].astype(np.float64)

    # bit-exact mode == the reference, whole frame, by digest
    proj, zmap = tsp.time_point_surface_projection(chunk, "TCZYX", mode="bitexact", **kw)
    assert proj.dtype == np.float64 and zmap.dtype == np.int64
    assert np.array_equal(zmap, want_zmap), "bitexact height map differs at %d pixels" % (zmap != want_zmap).sum()
    assert _sha(zmap) == meta["zmap_sha256"]
    p32 = proj.astype(np.float32)
    assert np.array_equal(p32.astype(np.float64), proj)
    if _sha(p32) != meta["proj_sha256"]:
        bad = p32[:, rows] != arr["proj_rows"]
        raise AssertionError("bitexact projection digest differs from the reference (%d of %d stored row pixels differ)"
                             % (bad.sum(), bad.size))
    ref_proj = proj                                   # = the reference's projection (digest), all pixels

    for mode in ("exact", "fast"):
        got_proj, got_zmap = tsp.time_point_surface_projection(chunk, "TCZYX", mode=mode, **kw)
        stats = compare_frame(got_proj, got_zmap, ref_proj, want_zmap, gap)
        # and directly against the stored reference rows (independent of the bitexact run above), away from the
        # +-8 px neighbourhood of tolerated height differences
        diff = got_zmap != want_zmap
        clean = ~(maximum_filter(diff.astype(np.uint8), size=17, mode="constant") > 0)[rows]
        rel = np.abs(got_proj[:, rows] - want_rows) / np.maximum(np.abs(want_rows), 1e-300)
        row_max = float(rel[:, clean].max())
        assert row_max <= PROJ_RTOL, (mode, row_max)
        viol = int((diff & ~near_tie).sum())
        print(name, mode, stats, "rule violations", viol, "row check max rel", row_max)
