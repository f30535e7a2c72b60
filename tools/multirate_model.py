"""numpy model of the ``fast`` focus-score stage (own design; the reference evaluates
surface_projection.py:37 + :55 as two full-rate scipy Gaussian blurs).  Design and validation tool only: the library
builds its tables in C++ (csrc/fast.cu); tests/test_multirate_design.py ties the two together (same coarse taps to
1e-15, operator error of the factorisation within the stated bound).

Along one in-plane axis of length N the reference applies ``H30c @ H1c`` (sigma=1 then sigma=30,
both edge-replicated, truncated at 4 sigma and renormalised).  The fast path factors it as

    U (B-spline reconstruction, stride 8 -> 1)  @  C (coarse FIR)  @  D (decimating prefilter)

  * D rows are ``d * (Rep @ H1c)``: a B-spline prefilter applied to the sigma=1-blurred,
    edge-replicated line - border rows are computed exactly here, so no clamp logic is left
    for the kernels beyond index clamping on the coarse grid;
  * C is fitted by least squares so that U C D reproduces the *truncated* sigma=30 kernel
    (including its renormalisation) for all 8 output phases;
  * along z the two sigma=0.5 passes are one exact banded Z x Z matrix.

Everything is float64 numpy on tiny 1-D problems.
"""
from __future__ import annotations

import functools

import numpy as np

STRIDE = 8
ORDER_D = 4           # prefilter = (box8)^4 -> 29 taps
ORDER_U = 4           # reconstruction = cubic B-spline on the stride-8 grid
RC = 16               # coarse FIR radius (coarse samples)
PAD_LO = 2            # coarse samples stored before position 0 (the first is fully outside)
D_TAPS = 37           # 29 (B-spline) + 8 (sigma=1 kernel radius 4 on both sides)
D_RADIUS = 18
ACC = 5               # coarse rows a fine row can contribute to


def gaussian_taps(sigma, truncate=4.0):
    """scipy.ndimage weights (scipy/ndimage/_filters.py): radius int(truncate*sigma+0.5)."""
    radius = int(truncate * float(sigma) + 0.5)
    k = np.arange(-radius, radius + 1, dtype=np.float64)
    w = np.exp(-0.5 / (float(sigma) ** 2) * k * k)
    return w / w.sum()


def bspline_taps(order, stride=STRIDE):
    k = np.ones(1)
    for _ in range(order):
        k = np.convolve(k, np.ones(stride) / stride)
    return k


def clamp_matrix(w, n):
    """n x n matrix of a centred FIR with edge replication ('nearest')."""
    r = (len(w) - 1) // 2
    m = np.zeros((n, n))
    rows = np.arange(n)
    for k in range(-r, r + 1):
        np.add.at(m, (rows, np.clip(rows + k, 0, n - 1)), w[k + r])
    return m


@functools.lru_cache(maxsize=None)
def coarse_taps(sigma_score=30.0, order_d=ORDER_D, order_u=ORDER_U, rc=RC):
    """Least-squares coarse FIR c[0..rc] (symmetric) such that, in the shift-invariant interior,
    sum_{j,i} u_p[j] c[i] d[.] matches the truncated/renormalised Gaussian for every phase p."""
    h = gaussian_taps(sigma_score)
    r_h = (len(h) - 1) // 2
    d = bspline_taps(order_d)
    u = bspline_taps(order_u) * STRIDE
    rd, ru = (len(d) - 1) // 2, (len(u) - 1) // 2
    span = STRIDE * rc + rd + ru + STRIDE
    n = 2 * span + 1                       # fine offsets s in [-span, span]
    # explicit enumeration (tiny problem): weight of P[x+s] in out[x], x = p
    A = np.zeros((STRIDE, n, rc + 1))
    for p in range(STRIDE):
        for m in range(-(ru // STRIDE) - 1, ru // STRIDE + 2):
            tu = p - STRIDE * m
            if abs(tu) > ru:
                continue
            for i in range(-rc, rc + 1):
                base = STRIDE * (m + i)
                for t in range(-rd, rd + 1):
                    s = base + t - p
                    if abs(s) <= span:
                        A[p, s + span, abs(i)] += u[tu + ru] * d[t + rd]
    b = np.zeros((STRIDE, n))
    b[:, span - r_h: span + r_h + 1] = h
    c, *_ = np.linalg.lstsq(A.reshape(-1, rc + 1), b.reshape(-1), rcond=None)
    resid = (A.reshape(-1, rc + 1) @ c - b.reshape(-1)).reshape(STRIDE, n)
    c = c / (c[0] + 2.0 * c[1:].sum())                 # unit DC gain (B-splines are a partition of unity)
    return c, float(np.abs(resid).sum(axis=1).max())


def coarse_len(n):
    """Stored coarse samples for a fine axis of length n.  Index mi < coarse_len-1 is the prefiltered
    sample at fine position 8*(mi - PAD_LO): mi = 0 lies fully before the line (it equals the
    replicated value P[0]), the last true sample is the last one whose B-spline support still
    touches the line.  Index coarse_len-1 is a stand-alone sample holding the replicated value
    P[n-1]; the coarse FIR clamps its reads to [0, coarse_len-1]."""
    return (n + 29) // STRIDE + 2


def decimation_matrix(n, sigma_pre=1.0, order_d=ORDER_D):
    """(coarse_len(n) x n) exact rows of  d * (Rep @ H1c)."""
    h1 = clamp_matrix(gaussian_taps(sigma_pre), n)
    d = bspline_taps(order_d)
    rd = (len(d) - 1) // 2
    mc = coarse_len(n)
    out = np.zeros((mc, n))
    for mi in range(mc - 1):
        centre = STRIDE * (mi - PAD_LO)
        for t in range(-rd, rd + 1):
            out[mi] += d[t + rd] * h1[min(max(centre + t, 0), n - 1)]
    assert STRIDE * (mc - 1 - PAD_LO) - rd > n - 1        # the next regular sample would be fully outside
    out[mc - 1] = h1[n - 1]
    return out


def coarse_matrix(mc_out_lo, mc_out_n, mc, c):
    """Rows m = mc_out_lo .. of the coarse FIR reading the stored coarse line with clamping."""
    rc = len(c) - 1
    out = np.zeros((mc_out_n, mc))
    for r in range(mc_out_n):
        m = mc_out_lo + r
        for i in range(-rc, rc + 1):
            out[r, min(max(m + i + PAD_LO, 0), mc - 1)] += c[abs(i)]
    return out


def reconstruction_matrix(n, order_u=ORDER_U):
    """(n x coarse_len(n)) B-spline reconstruction; coarse column index = m + PAD_LO."""
    u = bspline_taps(order_u) * STRIDE
    ru = (len(u) - 1) // 2
    mc = coarse_len(n)
    out = np.zeros((n, mc))
    for x in range(n):
        for mi in range(mc):
            t = x - STRIDE * (mi - PAD_LO)
            if abs(t) <= ru:
                out[x, mi] = u[t + ru]
    return out


def axis_operator(n, sigma_pre=1.0, sigma_score=30.0):
    """Dense n x n matrix of the fast path along one in-plane axis (for validation only)."""
    c, _ = coarse_taps(sigma_score)
    mc = coarse_len(n)
    return reconstruction_matrix(n) @ coarse_matrix(-PAD_LO, mc, mc, c) @ decimation_matrix(n, sigma_pre)


def exact_axis_operator(n, sigma_pre=1.0, sigma_score=30.0):
    return clamp_matrix(gaussian_taps(sigma_score), n) @ clamp_matrix(gaussian_taps(sigma_pre), n)


def z_operator(z, sigma=0.5, passes=2):
    """Exact banded Z x Z matrix of ``passes`` edge-replicated sigma=0.5 blurs."""
    b = clamp_matrix(gaussian_taps(sigma), z)
    return np.linalg.matrix_power(b, passes)
