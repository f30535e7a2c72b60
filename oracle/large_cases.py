"""Full-size parity cases = the five BASELINE.json configs at the sizes bench.py quotes (TEST INFRASTRUCTURE ONLY).

Inputs are rebuilt from seeds (``oracle/synth.py``, SURVEY 8(d): cfg1 seed 1, cfg2 seed 2, cfg3 seed 3 t=0.., cfg4
seed 4, cfg5 seed 5 in 2048 tiles); ``oracle/make_golden_large.py`` runs the UNMODIFIED reference on them once in
the build container and freezes digests + height maps under ``tests/golden/large/``.
"""
from . import synth

KW = dict(reference_channel=0, airyscan=False, z_map=True)


def _stack(Z, Y, X, C=1, seed=0, t=0):
    return lambda: synth.synth_stack(Z, Y, X, C=C, seed=seed, t=t)[None]


LARGE_CASES = [
    # name, builder of the (1,C,Z,Y,X) uint16 chunk, operator kwargs
    ("cfg1_512x512x32", _stack(32, 512, 512, seed=1), KW),
    ("cfg2_2048x2048x64", _stack(64, 2048, 2048, seed=2), KW),
    ("cfg3_1024x1024x48_t0", _stack(48, 1024, 1024, seed=3, t=0), KW),
    ("cfg3_1024x1024x48_t1", _stack(48, 1024, 1024, seed=3, t=1), KW),
    ("cfg3_1024x1024x48_t2", _stack(48, 1024, 1024, seed=3, t=2), KW),
    ("cfg3_1024x1024x48_t3", _stack(48, 1024, 1024, seed=3, t=3), KW),
    ("cfg4_2ch_2048x2048x64", _stack(64, 2048, 2048, C=2, seed=4), dict(KW, atoh_shift=0)),
    ("cfg5_tile00_2048x2048x128", lambda: synth.synth_tile(128, 2048, 4096, (0, 0), seed=5)[None], KW),
    ("cfg5_tile11_2048x2048x128", lambda: synth.synth_tile(128, 2048, 4096, (2048, 2048), seed=5)[None], KW),
    # stress inputs at sizes where the 241-tap kernel leaves the border regime
    ("sparse_spikes_640x640x16", lambda: synth.sparse_spike_stack(16, 640, 640, seed=9)[None], KW),
    ("white_noise_512x512x24", lambda: synth.white_noise_stack(24, 512, 512, seed=7)[None], KW),
    # trap T4 without the IndexError: min_z > 0, the band sits min_z planes too deep (SP:61 vs SP:66-69)
    ("zcrop_min3_768x768x24", _stack(24, 768, 768, seed=6), dict(KW, min_z=3, max_z=24)),
]
