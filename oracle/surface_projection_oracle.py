"""numpy/scipy restatement of the reference surface-projection hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference lines it restates (paths relative to the reference repository root;
``SP`` = tissue_analyzing_tool/surface_projection.py, ``BIM`` =
tissue_analyzing_tool/basic_image_manipulations.py, ``SPM`` =
tissue_analyzing_tool/surface_proj_m.py).

Third-party arithmetic the reference delegates to (no version pins exist in the
reference; the versions of this image define the numbers):
  * scipy.ndimage.gaussian_filter, scipy 1.18.1 (BIM:389)
  * numpy.percentile 'linear', numpy 2.3.5 (SP:35)
  * skimage.measure.block_reduce / skimage.transform.resize (SP:41-65, SPM:23-25):
    scikit-image is absent -> restated from its published behaviour, PARITY UNPINNED.

Parity status: pinned for bin_size == 1 (all BASELINE configs), build_manifold and
``surface_projection_m`` with a block_reduce restatement, by running the unmodified
reference in the build container (``oracle/reference_runner.py``) - see
``tests/golden`` and ``tests/test_oracle_vs_reference.py``.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import gaussian_filter

AIRYSCAN_PEDESTAL = 10000          # SP:28
CLIP_PERCENTILE = 95               # SP:35
SIGMA_PRE = (0.5, 1.0, 1.0)        # SP:37
SIGMA_SCORE = (0.5, 30.0, 30.0)    # SP:55
SIGMA_MASK = (1.0, 2.0, 2.0)       # SP:70-71
SIGMA_M = (5.0, 5.0, 3.0)          # SPM:18


# --------------------------------------------------------------------------- #
# primitives
# --------------------------------------------------------------------------- #
def blur_image(image, std):
    """BIM:373-390 - Gaussian blur with edge replication ('nearest')."""
    return gaussian_filter(image, std, mode="nearest")


def gaussian_taps(sigma, truncate=4.0):
    """Weights scipy uses for one axis (scipy/ndimage/_filters.py: radius = int(truncate *
    sigma + 0.5); exp(-k^2 / (2 sigma^2)) normalised to sum 1, float64)."""
    radius = int(truncate * float(sigma) + 0.5)
    k = np.arange(-radius, radius + 1, dtype=np.float64)
    w = np.exp(-0.5 / (float(sigma) * float(sigma)) * k * k)
    return w / w.sum()


def correlate1d_nearest(volume, sigma, axis, out_dtype=None):
    """Independent restatement of one scipy pass: float64 accumulation along ``axis`` over an
    edge-replicated line, then ONE cast to the output dtype (float32 rounds to nearest, integer
    dtypes truncate toward zero like the C cast scipy performs).  Used to cross-check that the
    oracle understands what ``gaussian_filter`` does (SURVEY trap T8), not on the timed path."""
    w = gaussian_taps(sigma)
    r = (len(w) - 1) // 2
    out_dtype = np.dtype(out_dtype or volume.dtype)
    src = np.moveaxis(np.asarray(volume), axis, -1).astype(np.float64)
    n = src.shape[-1]
    idx = np.clip(np.arange(-r, n + r), 0, n - 1)
    padded = src[..., idx]
    acc = np.zeros_like(src)
    for k in range(2 * r + 1):
        acc += w[k] * padded[..., k:k + n]
    if out_dtype.kind in "ui":
        acc = np.trunc(acc)
    return np.moveaxis(acc.astype(out_dtype), -1, axis)


def gaussian_filter_restated(volume, sigmas, out_dtype=None):
    """scipy.ndimage.gaussian_filter(mode='nearest') as a chain of the passes above, axis 0
    first, storing the output dtype between passes."""
    out = np.asarray(volume)
    out_dtype = np.dtype(out_dtype or out.dtype)
    for axis, s in enumerate(sigmas):
        out = correlate1d_nearest(out, s, axis, out_dtype)
    return out


def put_channel_axis_first(image, axes):
    """BIM:199-231.  Only transposes when 'C' is present and not already first; the target order
    is (C, [T], [Z], X, Y) - X before Y (SURVEY trap T7)."""
    c = axes.find("C")
    if c <= 0:
        return image, tuple(range(len(axes)))
    order = [c]
    for name in ("T", "Z"):
        pos = axes.find(name)
        if pos >= 0:
            order.append(pos)
    order += [axes.find("X"), axes.find("Y")]
    return np.transpose(image, axes=order), tuple(order)


def block_reduce(volume, block, func):
    """skimage.measure.block_reduce restated (PARITY UNPINNED): zero-pad every axis up to a
    multiple of the block, then reduce each block with ``func``."""
    volume = np.asarray(volume)
    pad = [(0, (-s) % b) for s, b in zip(volume.shape, block)]
    if any(p[1] for p in pad):
        volume = np.pad(volume, pad, mode="constant", constant_values=0)
    shape = []
    for s, b in zip(volume.shape, block):
        shape += [s // b, b]
    view = volume.reshape(shape)
    return func(view, axis=tuple(range(1, 2 * volume.ndim, 2)))


def resize_order1(image, out_shape):
    """skimage.transform.resize(order=1) stand-in (PARITY UNPINNED): bilinear resampling on
    pixel centres with edge reflection, no anti-aliasing (we only ever upsample here)."""
    from scipy.ndimage import zoom
    image = np.asarray(image)
    factors = [o / s for o, s in zip(out_shape, image.shape)]
    return zoom(image, factors, order=1, mode="mirror", grid_mode=True)


# --------------------------------------------------------------------------- #
# stages of time_point_surface_projection (bin_size == 1)
# --------------------------------------------------------------------------- #
def prepare_image(time_point, axes, airyscan, min_z, max_z):
    """SP:21-31: drop T, channel first, float32, airyscan pedestal, z-crop."""
    if axes.find("T") >= 0:
        time_point = time_point.reshape(time_point.shape[1:])
        axes = axes[1:]
    image, _ = put_channel_axis_first(time_point, axes)
    image = image.astype("float32")
    if airyscan:
        image -= AIRYSCAN_PEDESTAL
        image[image < 0] = 0
    if max_z > 0:
        image = image[:, min_z:max_z, :, :]
    return image


def clip_reference_channel(channel):
    """SP:32-36: clip above the 95th percentile of the non-zero voxels.  Returns (clipped, p95 or
    None when the channel has no non-zero voxel - SURVEY trap T2)."""
    pc = np.copy(channel)
    nz = pc[pc > 0]
    if nz.size == 0:
        return pc, None
    p = np.percentile(nz, CLIP_PERCENTILE)
    pc[pc > p] = p
    return pc, p


def focus_score(image, reference_channel):
    """SP:32-37 + SP:55: clip, sigma=(0.5,1,1) blur, sigma=(0.5,30,30) blur."""
    pc, _ = clip_reference_channel(image[reference_channel])
    pc = blur_image(pc, SIGMA_PRE)
    return blur_image(pc, SIGMA_SCORE)


def top2_relative_gap(score):
    """Per-pixel (best - second best) / best of the focus score along z; the north-star parity
    rule makes the height map binding only where this exceeds 1e-4."""
    if score.shape[0] < 2:
        return np.full(score.shape[1:], np.inf)
    part = np.partition(score, score.shape[0] - 2, axis=0)
    best, second = part[-1].astype(np.float64), part[-2].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        gap = (best - second) / np.abs(best)
    gap[~np.isfinite(gap)] = 0.0
    return gap


def band_mask(chosen_z, z_size):
    """SP:66-71: one-hot volume at the height map, blurred with sigma=(1,2,2).  Raises
    IndexError exactly when the reference does (SURVEY trap T4)."""
    y_size, x_size = chosen_z.shape
    onehot = np.zeros((z_size, y_size * x_size), dtype="float32")
    onehot[chosen_z.flatten(), np.arange(x_size * y_size)] = 1
    return blur_image(onehot.reshape((z_size, y_size, x_size)), SIGMA_MASK)


def band_mask_closed_form(chosen_z, z_size):
    """SURVEY trap T9: the same mask without materialising the one-hot volume first - z taps
    with edge replication evaluated from the height map, then the y and x passes."""
    wz = gaussian_taps(SIGMA_MASK[0])
    r = (len(wz) - 1) // 2
    z = np.arange(z_size)[:, None, None]
    acc = np.zeros((z_size,) + chosen_z.shape, dtype=np.float64)
    for k in range(-r, r + 1):
        acc += wz[k + r] * (np.clip(z + k, 0, z_size - 1) == chosen_z[None])
    mz = acc.astype(np.float32)
    my = correlate1d_nearest(mz, SIGMA_MASK[1], 1, np.float32)
    return correlate1d_nearest(my, SIGMA_MASK[2], 2, np.float32)


def project_channels(image, reference_channel, mask, mask_other):
    """SP:72-79: per channel max over z of intensity * mask (float32 product, float64 output)."""
    channels = image.shape[0]
    projection = np.zeros((channels,) + image.shape[-2:])
    for c in range(channels):
        m = mask if c == reference_channel else mask_other
        projection[c] = np.max(image[c] * m, axis=0)
    return projection


# --------------------------------------------------------------------------- #
# continuous manifold (SP:87-165)
# --------------------------------------------------------------------------- #
def find_pixel_plane(score, chosen, row, col, n_rows, n_cols, n_planes):
    """SP:130-165.  Looks at the up/down/left/right neighbours that already have a plane and
    restricts the search to +-1 of them.  Quirks kept: the ``row >= 0`` test lets row -1 wrap
    (SP:133-134) and a two-apart neighbour pair returns their float mean (SP:165)."""
    first = second = None

    def consider(plane):
        nonlocal first, second
        if plane >= 0:
            if first is None:
                first = plane
            else:
                second = plane

    if row >= 0:
        plane = chosen[row - 1, col]
        if plane >= 0:
            first = plane
    if row < n_rows - 1:
        consider(chosen[row + 1, col])
    if second is None and col > 0:
        consider(chosen[row, col - 1])
    if second is None and col < n_cols - 1:
        consider(chosen[row, col + 1])

    if second is None or first == second:
        lo = max(0, first - 1)
        return lo + np.argmax(score[lo:min(n_planes, first + 2), row, col])
    if np.abs(first - second) == 1:
        lo = max(0, min(first, second))
        return lo + np.argmax(score[lo:min(n_planes, min(first, second) + 2), row, col])
    return (first + second) / 2


def build_continues_manifold(score):
    """SP:87-128: grow the height map outwards from the global maximum of the score in square
    rings, visiting each ring in the reference's order (right edge lower half, bottom edge
    right-to-left, left edge bottom-to-top, top edge left-to-right, right edge upper half)."""
    n_planes, n_rows, n_cols = score.shape
    chosen = -1 * np.ones((n_rows, n_cols)).astype(int)
    p0, r0, c0 = np.unravel_index(np.argmax(score), score.shape)
    chosen[r0, c0] = p0
    reach = np.max(np.abs(np.array([c0, r0, c0, r0]) - np.array([0, 0, n_cols - 1, n_rows - 1])))

    def visit(row, col):
        chosen[row, col] = find_pixel_plane(score, chosen, row, col, n_rows, n_cols, n_planes)

    for d in range(1, reach + 1):
        if c0 + d < n_cols:
            for row in range(r0, r0 + d + 1):
                if row < n_rows:
                    visit(row, c0 + d)
        if r0 + d < n_rows:
            for col in range(c0 + d - 1, c0 - d - 1, -1):
                if 0 <= col < n_cols:
                    visit(r0 + d, col)
        if c0 - d >= 0:
            for row in range(r0 + d - 1, r0 - d - 1, -1):
                if 0 <= row < n_rows:
                    visit(row, c0 - d)
        if r0 - d >= 0:
            for col in range(c0 - d + 1, c0 + d + 1):
                if 0 <= col < n_cols:
                    visit(r0 - d, col)
        if c0 + d < n_cols:
            for row in range(r0 - d + 1, r0):
                if row >= 0:
                    visit(row, c0 + d)
    return chosen


# --------------------------------------------------------------------------- #
# the operator
# --------------------------------------------------------------------------- #
def time_point_surface_projection(time_point, axes, reference_channel, min_z=0, max_z=0,
                                  method="max_averages", bin_size=1, airyscan=True, z_map=False,
                                  atoh_shift=0, build_manifold=False, return_score=False):
    """SP:17-85 (behavioural spec in SURVEY.md section 3.3).  ``return_score`` is an oracle-only
    extra used by the parity harness to evaluate the top-2 gap rule."""
    image = prepare_image(time_point, axes, airyscan, min_z, max_z)
    pc, _ = clip_reference_channel(image[reference_channel])
    pc = blur_image(pc, SIGMA_PRE)
    z_size, y_size, x_size = image.shape[-3:]
    if bin_size > 1:
        blk = (1, bin_size, bin_size)
        if method == "max_averages":
            score = block_reduce(blur_image(pc, SIGMA_SCORE), blk, np.mean)
        elif method == "max_std":
            score = block_reduce(pc, blk, np.var)
        elif method == "multi_channel":
            other = np.copy(image[(reference_channel + 1) % image.shape[0]])
            p_other = np.percentile(other, CLIP_PERCENTILE)          # SP:46 - zeros included
            other[other > p_other] = p_other
            other = blur_image(other, SIGMA_PRE)
            score = (block_reduce(blur_image(other, SIGMA_SCORE), blk, np.mean)
                     * block_reduce(pc, blk, np.var))
        else:
            raise TypeError("exceptions must derive from BaseException")   # SP:53 raises a str
    else:
        score = blur_image(pc, SIGMA_SCORE)
    if build_manifold:
        chosen_z = build_continues_manifold(score)
    else:
        if score.shape[1:] != (y_size, x_size):
            score = resize_order1(score.astype("float32"), (z_size, y_size, x_size))
        chosen_z = min_z + np.argmax(score, axis=0)
    if atoh_shift == 0:
        chosen_z_other = np.copy(chosen_z)
    else:
        chosen_z_other = np.clip(chosen_z + atoh_shift, 0, score.shape[0])   # SP:62, inclusive
    if chosen_z.shape != (y_size, x_size):
        chosen_z = np.round(resize_order1(chosen_z.astype("float32"), (y_size, x_size))).astype("int")
        chosen_z_other = np.round(
            resize_order1(chosen_z_other.astype("float32"), (y_size, x_size))).astype("int")
    mask = band_mask(chosen_z, z_size)
    mask_other = band_mask(chosen_z_other, z_size)
    if axes.find("C") >= 0:
        projection = project_channels(image, reference_channel, mask, mask_other)
    else:
        projection = np.max(image * mask, axis=0)
    out = (projection, chosen_z) if z_map else projection
    if return_score:
        return out, score
    return out


# --------------------------------------------------------------------------- #
# surface_projection_m (SPM:14-47, 81-100)
# --------------------------------------------------------------------------- #
def expand_score(score, bin_size):
    """SPM:81-100 without the Python triple loop: nearest-neighbour upsampling of every plane by
    ``bin_size`` in both in-plane axes, returned with z LAST: (rows, cols, Z)."""
    up = np.repeat(np.repeat(score, bin_size, axis=1), bin_size, axis=2)
    return np.moveaxis(up, 0, 2)


def surface_projection_m(time_point, axes, reference_channel, min_z, max_z, method, bin_size):
    """SPM:14-35 with the ``put_cannel_axis_first`` typo (SPM:15) read as the function it means.
    The blur keeps the uint16 dtype, so every one of the three passes truncates;
    ``np.choose`` limits Z to 64 planes."""
    image, _ = put_channel_axis_first(time_point, axes)
    stack = image[reference_channel][min_z:max_z]
    stack = blur_image(stack, SIGMA_M)
    blk = (1, bin_size, bin_size)
    if method == "max_averages":
        score = block_reduce(stack, blk, np.mean)
    elif method == "max_std":
        score = block_reduce(stack, blk, np.var)
    else:
        raise TypeError("exceptions must derive from BaseException")       # SPM:27 raises a str
    z, rows, cols = stack.shape
    fixed = expand_score(score, bin_size)[:rows, :cols, :z]                  # SPM:44-47
    best_z = np.argmax(fixed, axis=2)
    return np.choose(best_z, stack).reshape((rows, cols))                    # SPM:37-41
