"""Empirical check of the error band behind the search kernels' fp32 pre-filter (elimaloc_b200/csrc/icp_kernels.cu,
`visit_points`): the kernel scans candidates with fp32 distances, keeps the minimum m, and treats the fp32 argmin as the unique
exact winner only if the second-smallest value lies above  T(m) = ((sqrt(m) (1 + 2^-20) + 2^-20 |p|_1)^2) (1 + 2^-18) + 1e-30.
Exactness needs:  every candidate j that is at least as close as the fp32 argmin IN EXACT ARITHMETIC has d32_j <= T(m).
This test re-computes d32 with the kernel's operation order in numpy float32 (the two FFMAs are evaluated in float64 and
rounded once, which can differ from a true fused operation by one fp32 ulp at most — far inside the 2-4x slack of the band)
and the exact distances with Python integers, over adversarial populations: near-ties down to one ulp, coordinates up to
1e5 m, distances from 1e-4 m to 10 m, voxel means rounded to fp32 (the VGICP variant)."""
from fractions import Fraction

import numpy as np
import pytest


def d32_like_the_kernel(q32, p64):
    f = p64.astype(np.float32)                                           # Query: fp32 rounding of the fp64 query
    dx, dy, dz = (q32[..., k] - f[..., k] for k in range(3))              # three FADDs (float32 arrays: one rounding each)
    t = (dx * dx).astype(np.float32)                                      # FMUL
    t = (dy.astype(np.float64) * dy.astype(np.float64) + t.astype(np.float64)).astype(np.float32)   # FFMA
    t = (dz.astype(np.float64) * dz.astype(np.float64) + t.astype(np.float64)).astype(np.float32)   # FFMA
    band = ((np.abs(f[..., 0]) + np.abs(f[..., 1])).astype(np.float32) + np.abs(f[..., 2])).astype(np.float32) * np.float32(2.0 ** -20)
    return t, band


def threshold(m, band):
    sm = (np.sqrt(m).astype(np.float64) * float(np.float32(1.0 + 2.0 ** -20)) + band.astype(np.float64)).astype(np.float32)
    return ((sm * sm).astype(np.float32).astype(np.float64) * float(np.float32(1.0 + 2.0 ** -18)) + 1e-30).astype(np.float32)


def exact_d2(q, p):
    """exact rational squared distance between a float32 (or float64) candidate and the float64 query"""
    return sum((Fraction(float(a)) - Fraction(float(b))) ** 2 for a, b in zip(q, p))


@pytest.mark.parametrize("scale", [1.0, 300.0, 1e4, 1e5])
@pytest.mark.parametrize("radius", [1e-4, 0.05, 0.7, 10.0])
@pytest.mark.parametrize("means", [False, True])
def test_no_true_winner_falls_outside_the_band(scale, radius, means):
    rng = np.random.default_rng(int(scale) * 7 + int(radius * 1e4) + means)
    n_q, n_c = 300, 12
    p = (rng.random((n_q, 3)) * 2 - 1) * scale                                            # fp64 queries
    # candidates on a sphere of (almost) equal radius around the query: near-ties of every size, down to a few ulps
    d = rng.normal(size=(n_q, n_c, 3))
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    r = radius * (1.0 + rng.choice([0.0, 1e-9, 1e-7, 3e-6, 1e-4, 1e-2], size=(n_q, n_c)) * rng.normal(size=(n_q, n_c)))
    c64 = p[:, None, :] + d * r[..., None]
    c32 = c64.astype(np.float32)             # stored points (or voxel means rounded for the pre-filter)
    exact_pos = c64 if means else c32.astype(np.float64)   # VGICP decides on the fp64 mean, P2P/GICP on the fp32 point itself
    d32, band = d32_like_the_kernel(c32, np.broadcast_to(p[:, None, :], c64.shape))
    m = d32.min(axis=1)
    T = threshold(m, band[:, 0])
    checked = 0
    for i in range(n_q):
        i_star = int(np.argmin(d32[i]))
        ex = [exact_d2(exact_pos[i, j], p[i]) for j in range(n_c)]
        for j in range(n_c):
            if ex[j] <= ex[i_star]:              # j would win or tie in exact arithmetic ...
                assert d32[i, j] <= T[i], (i, j, float(d32[i, j]), float(T[i]))   # ... so the filter must keep it
                checked += 1
    assert checked >= n_q                        # (every query contributes at least its own argmin)


def test_band_is_not_vacuous():
    """the band is tight enough to be useful: well-separated candidates are resolved without the exact re-scan"""
    rng = np.random.default_rng(0)
    p = (rng.random((2000, 3)) * 2 - 1) * 100.0
    c32 = (p[:, None, :] + rng.normal(size=(2000, 27, 3)) * 0.5).astype(np.float32)       # a voxel column's worth of candidates
    d32, band = d32_like_the_kernel(c32, np.broadcast_to(p[:, None, :], c32.shape))
    s = np.sort(d32, axis=1)
    T = threshold(s[:, 0], band[:, 0])
    assert (s[:, 1] > T).mean() > 0.995
