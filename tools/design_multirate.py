"""Design study for the multirate focus-score stage (own design, not in the reference).

Approximates the 1-D operator  H30c @ H1c  (scipy 'nearest' Gaussian sigma=1 then sigma=30,
reference surface_projection.py:37,55) by  U @ C @ Dt :
  Dt  decimate-by-8 prefilter d (B-spline) applied to the sigma=1-blurred, edge-replicated line
  C   coarse-grid FIR found by least squares
  U   B-spline reconstruction back to the fine grid.
Prints the worst-case (L1) and typical errors of the composite.
"""
import numpy as np


def gaussian_taps(sigma, truncate=4.0):
    """scipy.ndimage.gaussian_filter1d's weights: radius int(truncate * sigma + 0.5), normalised to sum 1."""
    radius = int(truncate * float(sigma) + 0.5)
    k = np.arange(-radius, radius + 1, dtype=np.float64)
    w = np.exp(-0.5 / (float(sigma) * float(sigma)) * k * k)
    return w / w.sum()

S = 8


def bspline_box(order, s=S):
    k = np.ones(1)
    for _ in range(order):
        k = np.convolve(k, np.ones(s) / s)
    return k          # length order*(s-1)+1, sums to 1


def clamp_conv_matrix(w, n):
    r = (len(w) - 1) // 2
    M = np.zeros((n, n))
    for k in range(-r, r + 1):
        idx = np.clip(np.arange(n) + k, 0, n - 1)
        np.add.at(M, (np.arange(n), idx), w[k + r])
    return M


def composite_kernels(d, c, u_order):
    """K_p[s] for the shift-invariant interior, p = 0..7.  d: fine taps (centered), c: coarse taps
    (centered), U: centred B-spline of given order sampled at fine offsets."""
    u = bspline_box(u_order) * S          # interpolation kernel on fine grid (partition of unity)
    # zero-stuffed coarse filter on fine grid
    cz = np.zeros((len(c) - 1) * S + 1)
    cz[::S] = c
    full = np.convolve(np.convolve(d, cz), u)       # response of fine impulse -> fine output, phase-mixed
    return full, u


def build_problem(d, rc, u_order):
    """Linear map from symmetric coarse taps (rc+1 unknowns) to stacked per-phase composite
    kernels, so C can be found by least squares against h30."""
    u = bspline_box(u_order) * S
    ru = (len(u) - 1) / 2.0
    rd = (len(d) - 1) / 2.0
    # output at fine position n = 8m+p; a[m'] = sum_t d[t] P[8m' + t - rd_off]
    # we work with explicit index bookkeeping on a long line
    L = 8 * 80
    centre = L // 2 - (L // 2) % 8 + 8 * 0
    rows = []
    return None


def design(order_d=6, order_u=6, rc=12, n_line=1024, verbose=True, extra_d=None):
    h30 = gaussian_taps(30.0)
    h1 = gaussian_taps(1.0)
    d = bspline_box(order_d) if extra_d is None else extra_d
    u = bspline_box(order_u) * S
    # Work on a long zero-padded line with plain convolutions; impulse position sweeps one coarse period
    # Operator pieces as matrices on a line of length n_line (interior only is inspected).
    n = n_line
    mc = n // S
    # centred alignment: make d and u odd-length symmetric about integer centres.
    assert len(d) % 2 == 1 and len(u) % 2 == 1, (len(d), len(u))
    rd = (len(d) - 1) // 2
    ru = (len(u) - 1) // 2
    D = np.zeros((mc, n))
    for m in range(mc):
        for t in range(-rd, rd + 1):
            j = S * m + t
            if 0 <= j < n:
                D[m, j] = d[t + rd]
    U = np.zeros((n, mc))
    for x in range(n):
        for m in range(mc):
            t = x - S * m
            if -ru <= t <= ru:
                U[x, m] = u[t + ru]
    # unknown symmetric c: C = sum_i c_i * Shift_i
    basis = []
    for i in range(rc + 1):
        Ci = np.zeros((mc, mc))
        for m in range(mc):
            for sgn in ((1, -1) if i else (1,)):
                mm = m + sgn * i
                if 0 <= mm < mc:
                    Ci[m, mm] = 1.0
        basis.append(U @ Ci @ D)          # n x n
    # target: plain h30 convolution rows, interior outputs x in one coarse period near the middle
    x0 = (n // 2 // S) * S
    rows = range(x0, x0 + S)
    r30 = 120
    A = []
    b = []
    for x in rows:
        tgt = np.zeros(n)
        tgt[x - r30:x + r30 + 1] = h30
        A.append(np.stack([B[x] for B in basis], axis=1))
        b.append(tgt)
    A = np.concatenate(A, axis=0)
    b = np.concatenate(b)
    c, *_ = np.linalg.lstsq(A, b, rcond=None)
    res = (A @ c - b).reshape(S, n)
    l1 = np.abs(res).sum(axis=1)
    if verbose:
        print(f"d order {order_d} ({len(d)} taps) u order {order_u} rc {rc}: "
              f"L1 err per phase max {l1.max():.3e} mean {l1.mean():.3e}; sum c {c[0] + 2 * c[1:].sum():.8f}")
    return d, c, u, l1


if __name__ == "__main__":
    for od, ou in ((4, 4), (4, 6), (6, 6), (6, 4), (8, 6)):
        for rc in (10, 12, 14, 16):
            design(od, ou, rc)
