"""The fast-mode filter design (own design, csrc/fast.cu) against its numpy model (tools/multirate_model.py) and the
design study (tools/design_multirate.py).  CPU only: the taps come out of the library through
tsp_debug_coarse_taps, no kernel runs."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def library_taps():
    import __graft_entry__ as entry
    entry.build()
    from tissue_image_processing_b200 import _native
    lib = _native.load_library()
    buf = (ctypes.c_double * 32)()
    n = lib.tsp_debug_coarse_taps(buf, 32)
    assert n == 17
    assert lib.tsp_debug_coarse_taps(buf, 4) < 0          # capacity too small: an error code, not an overrun
    return np.array(buf[:n])


def test_library_taps_equal_the_model(library_taps):
    import multirate_model as mm
    c, l1 = mm.coarse_taps()
    assert c.shape == library_taps.shape
    assert np.abs(c - library_taps).max() < 1e-15
    assert abs(library_taps[0] + 2 * library_taps[1:].sum() - 1.0) < 1e-15      # unit DC gain
    assert l1 < 6.5e-5                                    # worst-case (L1) operator error per axis, DESIGN.md section 4


def test_library_taps_equal_the_design_study(library_taps):
    import design_multirate as dm
    _, c, _, l1 = dm.design(order_d=4, order_u=4, rc=16, verbose=False)
    c = c / (c[0] + 2.0 * c[1:].sum())
    assert np.abs(c - library_taps).max() < 1e-12        # a different least-squares formulation of the same fit
    assert l1.max() < 6.5e-5


@pytest.mark.parametrize("n", [96, 333, 520])
def test_factorised_axis_operator_stays_within_budget(n):
    """U C D against the exact edge-replicated sigma=1 then sigma=30 operator, borders included: the row-wise L1
    error bounds the relative score error on any non-negative line."""
    import multirate_model as mm
    err = np.abs(mm.axis_operator(n) - mm.exact_axis_operator(n)).sum(axis=1)
    assert err.max() < 2e-4, err.max()
    assert np.allclose(mm.axis_operator(n).sum(axis=1), 1.0, atol=1e-6)         # constants pass through unchanged
