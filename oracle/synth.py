"""Synthetic confocal stacks used by tests, smoke and bench.

The generator is deterministic in ``(Z, Y, X, C, seed, t)`` so the GPU box and the
build container create identical inputs without shipping data.
"""
import numpy as np


def height_field(Z, Y, X, t=0):
    """Smooth tissue surface h(y, x) in plane units."""
    y = np.arange(Y, dtype=np.float64)[:, None]
    x = np.arange(X, dtype=np.float64)[None, :]
    return (Z / 2.0
            + 0.15 * Z * np.sin(2 * np.pi * 1.5 * y / Y + 0.1 * t)
            + 0.10 * Z * np.cos(2 * np.pi * x / X))


def _blob_texture(rng, Y, X):
    """Low-frequency blob pattern in [0.5, 1] for the second channel."""
    gy, gx = max(2, Y // 24), max(2, X // 24)
    coarse = rng.random((gy, gx))
    reps = (-(-Y // gy), -(-X // gx))
    return 0.5 + 0.5 * np.kron(coarse, np.ones(reps))[:Y, :X]


def synth_stack(Z, Y, X, C=1, seed=0, t=0, airyscan=False, dtype=np.uint16):
    """Return a ``(C, Z, Y, X)`` uint16 stack with a bright sheet at ``height_field``.

    Channel 0 (reference): 300 + 2500 * exp(-(z-h)^2 / (2*2^2)) * tex, tex a sparse
    binary texture; noise 8*Poisson(sig/8) + N(0, 20).  Channel c>0: the sheet is one
    plane deeper per channel, amplitude 1500, blob texture.  ``airyscan`` adds the
    10000-count pedestal that reference surface_projection.py:27-29 subtracts.
    """
    rng = np.random.default_rng(seed + 1000 * t)
    h = height_field(Z, Y, X, t)
    z = np.arange(Z, dtype=np.float64)[:, None, None]
    out = np.empty((C, Z, Y, X), dtype=dtype)
    for c in range(C):
        if c == 0:
            tex = 0.5 + 0.5 * (rng.random((Y, X)) < 0.15)
            amp = 2500.0
        else:
            tex = _blob_texture(rng, Y, X)
            amp = 1500.0
        sig = 300.0 + amp * np.exp(-(z - (h + c)[None]) ** 2 / (2 * 2.0 ** 2)) * tex[None]
        img = 8.0 * rng.poisson(sig / 8.0) + rng.normal(0.0, 20.0, size=sig.shape)
        if airyscan:
            img = img + 10000.0
        out[c] = np.clip(np.rint(img), 0, 65535).astype(dtype)
    return out


def white_noise_stack(Z, Y, X, C=1, seed=0, high=4096):
    """Stress input: uniform white noise, near-ties in the focus score everywhere."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, high, size=(C, Z, Y, X), dtype=np.uint16)


def sparse_spike_stack(Z, Y, X, C=1, seed=0, density=0.01, amp=4000):
    """Stress input: isolated bright voxels on a zero background (exercises the p95 clip
    over non-zero voxels and the hard 4-sigma cut-off of the sigma=30 kernel)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((C, Z, Y, X)) < density
    vals = rng.integers(amp // 4, amp, size=(C, Z, Y, X))
    return np.where(mask, vals, 0).astype(np.uint16)


def synth_tile(Z, tile, full, origin, seed=0, C=1):
    """One ``tile x tile`` XY tile of a ``full x full`` frame (BASELINE configs[4]: the 4096x4096x128 stack is
    projected in independent ``chunk_size`` tiles, reference surface_projection.py:294-301).  The sheet follows
    the height field of the WHOLE frame evaluated at the tile's global coordinates; texture and noise come from
    a generator seeded per tile, so a tile is reproducible without building the multi-GB frame around it."""
    y0, x0 = origin
    rng = np.random.default_rng(seed + 7919 * (1 + (y0 // tile) * (-(-full // tile)) + x0 // tile))
    h = height_field(Z, full, full)[y0:y0 + tile, x0:x0 + tile]
    z = np.arange(Z, dtype=np.float64)[:, None, None]
    out = np.empty((C, Z) + h.shape, dtype=np.uint16)
    for c in range(C):
        tex = 0.5 + 0.5 * (rng.random(h.shape) < 0.15) if c == 0 else _blob_texture(rng, *h.shape)
        sig = 300.0 + (2500.0 if c == 0 else 1500.0) * np.exp(-(z - (h + c)[None]) ** 2 / 8.0) * tex[None]
        img = 8.0 * rng.poisson(sig / 8.0) + rng.normal(0.0, 20.0, size=sig.shape)
        out[c] = np.clip(np.rint(img), 0, 65535).astype(np.uint16)
    return out
