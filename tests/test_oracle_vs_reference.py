"""Oracle vs the live reference import (only where /root/reference is mounted: the build
container).  Randomised shapes beyond the frozen goldens."""
import numpy as np
import pytest

from oracle import reference_runner, synth
from oracle import surface_projection_oracle as orc

pytestmark = pytest.mark.skipif(not reference_runner.available(),
                                reason="reference repository not mounted on this box")


@pytest.mark.parametrize("seed", range(4))
def test_random_shapes_bit_exact(seed):
    sp = reference_runner.load_surface_projection()
    rng = np.random.default_rng(100 + seed)
    Z, Y, X, C = int(rng.integers(2, 14)), int(rng.integers(20, 90)), int(rng.integers(20, 90)), int(rng.integers(1, 4))
    img = synth.synth_stack(Z, Y, X, C=C, seed=seed, airyscan=bool(seed % 2))[None]
    kw = dict(reference_channel=int(rng.integers(0, C)), airyscan=bool(seed % 2), z_map=True,
              atoh_shift=int(rng.integers(-1, 2)) if Z > 4 else 0)
    try:
        want = sp.time_point_surface_projection(img, "TCZYX", **kw)
    except IndexError:
        with pytest.raises(IndexError):
            orc.time_point_surface_projection(img, "TCZYX", **kw)
        return
    got = orc.time_point_surface_projection(img, "TCZYX", **kw)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert got[0].dtype == want[0].dtype and got[1].dtype == want[1].dtype


def test_goldens_are_current(golden):
    """The committed fixtures are what the mounted reference produces today."""
    from oracle import golden_cases
    sp = reference_runner.load_surface_projection()
    _, arrays = golden
    for name, build, axes, kw in golden_cases.CASES[:6]:
        res = sp.time_point_surface_projection(build(), axes, **kw)
        proj = res[0] if kw.get("z_map") else res
        assert np.array_equal(proj, arrays[name + "/projection"].astype(np.float64)), name
