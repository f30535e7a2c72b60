"""Catalogue of the golden cases frozen under tests/golden/ (TEST INFRASTRUCTURE ONLY).

Each case is (name, input builder, operator kwargs).  Inputs are rebuilt from seeds by
``oracle/synth.py`` so only the reference OUTPUTS are stored.
"""
import numpy as np

from . import synth


def _structured(Z, Y, X, C=1, seed=1, airyscan=False):
    return synth.synth_stack(Z, Y, X, C=C, seed=seed, airyscan=airyscan)[None]   # (1,C,Z,Y,X)


def _white(Z, Y, X, C=1, seed=7):
    return synth.white_noise_stack(Z, Y, X, C=C, seed=seed)[None]


def _sparse(Z, Y, X, C=1, seed=9):
    return synth.sparse_spike_stack(Z, Y, X, C=C, seed=seed)[None]


def _zeros(Z, Y, X):
    return np.zeros((1, 1, Z, Y, X), dtype=np.uint16)


def _saturated(Z, Y, X, seed=11):
    """Large flat regions at identical values -> exact score ties across z (trap T3)."""
    rng = np.random.default_rng(seed)
    a = np.full((1, 1, Z, Y, X), 500, dtype=np.uint16)
    a[0, 0, Z // 3: Z // 3 + 2, : Y // 2] = 4000
    a[0, 0, :, :, X // 2:] += rng.integers(0, 3, size=(Z, Y, X - X // 2)).astype(np.uint16)
    return a


CASES = [
    # name, builder, axes, kwargs
    ("structured_small", lambda: _structured(12, 96, 112), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True)),
    ("structured_interior", lambda: _structured(16, 288, 320, seed=2), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True)),
    ("white_noise", lambda: _white(10, 80, 72), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True)),
    ("sparse_spikes", lambda: _sparse(10, 96, 96), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True)),
    ("airyscan_default", lambda: _structured(12, 64, 80, seed=3, airyscan=True), "TCZYX",
     dict(reference_channel=0, z_map=True)),                          # airyscan defaults to True
    ("airyscan_all_below_pedestal", lambda: _structured(8, 48, 40, seed=4), "TCZYX",
     dict(reference_channel=0, airyscan=True, z_map=True)),           # everything clamps to 0 (T2)
    ("all_zero", lambda: _zeros(6, 40, 56), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True)),
    ("two_channel_shift0", lambda: _structured(14, 72, 88, C=2, seed=5), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True, atoh_shift=0)),
    ("two_channel_shift2_ref1", lambda: _structured(14, 72, 88, C=2, seed=6), "TCZYX",
     dict(reference_channel=1, airyscan=False, z_map=True, atoh_shift=2)),
    ("three_channel_shift_neg", lambda: _structured(12, 64, 64, C=3, seed=8), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True, atoh_shift=-3)),
    ("zcrop_min0", lambda: _structured(20, 64, 72, seed=10), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True, min_z=0, max_z=14)),
    # trap T4 without the IndexError: chosen_z carries +min_z but indexes the CROPPED stack (band min_z planes too deep)
    ("zcrop_min3_band_deeper", lambda: _structured(20, 64, 72, seed=20), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True, min_z=3, max_z=20)),
    ("zcrop_min2_two_channels_shift", lambda: _structured(22, 56, 64, C=2, seed=24), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True, min_z=2, max_z=21, atoh_shift=1)),
    ("saturated_ties", lambda: _saturated(9, 64, 64), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True)),
    ("no_zmap_return", lambda: _structured(8, 48, 48, seed=12), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=False)),
    ("axes_czyx", lambda: _structured(8, 48, 56, C=2, seed=13)[0], "CZYX",
     dict(reference_channel=0, airyscan=False, z_map=True)),
    ("axes_zcyx_transposed", lambda: np.ascontiguousarray(
        np.moveaxis(_structured(8, 40, 56, C=2, seed=14)[0], 0, 1)), "ZCYX",
     dict(reference_channel=1, airyscan=False, z_map=True)),          # output is (C, X, Y) (T7)
    ("odd_sizes", lambda: _structured(7, 53, 67, seed=15), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True)),
    ("single_plane", lambda: _structured(1, 40, 40, seed=16), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True)),
    ("manifold", lambda: _structured(10, 40, 44, seed=17), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True, build_manifold=True)),
]

# cases where the reference raises instead of returning (frozen as the exception type)
ERROR_CASES = [
    ("zcrop_min2_indexerror", lambda: _structured(20, 48, 48, seed=18), "TCZYX",
     dict(reference_channel=0, airyscan=False, z_map=True, min_z=12, max_z=20)),
    ("axes_zyx_runtimeerror", lambda: _structured(6, 32, 32, seed=19)[0, 0], "ZYX",
     dict(reference_channel=0, airyscan=False, z_map=True)),
]

SPM_CASES = [
    ("spm_mean_bin4", lambda: _structured(12, 96, 112, C=2, seed=21)[0], "CZYX",
     dict(reference_channel=0, min_z=0, max_z=12, method="max_averages", bin_size=4)),
    ("spm_var_bin5_crop", lambda: _structured(14, 90, 75, C=2, seed=22)[0], "CZYX",
     dict(reference_channel=1, min_z=2, max_z=11, method="max_std", bin_size=5)),
    ("spm_bin1", lambda: _white(8, 48, 40, seed=23)[0], "CZYX",
     dict(reference_channel=0, min_z=0, max_z=8, method="max_averages", bin_size=1)),
]
