"""Parity rules of BASELINE.json's north_star, shared by the GPU tests, smoke() and bench.py:
  * height map bit-exact wherever the oracle's top-two focus scores differ by more than 1e-4 relative;
  * projection within 1e-5 relative (float32), evaluated away from tolerated height-map differences
    (a differing pixel changes the sigma=2 blurred mask inside +-8 px).
"""
import numpy as np
from scipy.ndimage import maximum_filter

GAP_RULE = 1e-4
PROJ_RTOL = 1e-5


def compare_frame(got_proj, got_zmap, want_proj, want_zmap, gap, exact_zmap=False, proj_rtol=PROJ_RTOL):
    """Returns a dict of statistics; raises AssertionError when a rule is violated.
    gap: (Y,X) relative top-2 gap of the oracle score (oracle.top2_relative_gap)."""
    got_zmap = np.asarray(got_zmap)
    want_zmap = np.asarray(want_zmap)
    assert got_zmap.shape == want_zmap.shape, (got_zmap.shape, want_zmap.shape)
    assert got_proj.shape == want_proj.shape, (got_proj.shape, want_proj.shape)
    diff = got_zmap != want_zmap
    stats = {"zmap_mismatch": int(diff.sum()), "pixels": int(diff.size),
             "near_tie_pixels": int((gap <= GAP_RULE).sum())}
    if exact_zmap:
        assert not diff.any(), "height map differs at %d pixels (bit-exact mode)" % diff.sum()
    binding = diff & (gap > GAP_RULE)
    assert not binding.any(), ("height map differs at %d pixels whose oracle top-2 gap exceeds %g (max gap %.3g)"
                               % (binding.sum(), GAP_RULE, gap[diff].max()))
    tainted = maximum_filter(diff.astype(np.uint8), size=17, mode="constant") > 0 if diff.any() else diff
    ok = ~tainted
    g = np.asarray(got_proj, dtype=np.float64)[..., ok]
    w = np.asarray(want_proj, dtype=np.float64)[..., ok]
    err = np.abs(g - w)
    tol = proj_rtol * np.abs(w)
    bad = err > tol
    stats["proj_max_rel"] = float((err / np.maximum(np.abs(w), 1e-300)).max()) if w.size else 0.0
    assert not bad.any(), ("projection differs beyond %g relative at %d pixels (max rel %.3g)"
                           % (proj_rtol, bad.sum(), stats["proj_max_rel"]))
    return stats
