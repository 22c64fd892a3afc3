"""Model check of the VGICP search's quantised pre-filter (elimaloc_b200/csrc/icp_kernels.cu `nearest_mean_27`, voxel_key.hpp
`pack_vcand`): the <= 27 voxel means of a directory entry are stored as 13-bit fixed-point offsets from the entry's key, in voxel
sizes, over [-2, 2] (step 4 / 8191, error <= 2.45e-4 per axis, 4.3e-4 on the vector); the kernel scans them with fp32 distances
to the query's in-cell fraction and takes the fp32 argmin as THE answer of GetCorrespondencesCov (voxel_hash_map.cpp:90-151) only if
the second-smallest distance lies outside a band of 1.2e-3 voxel sizes:  s2 > (sqrt(m) + 1.2e-3)^2 (1 + 1e-5); otherwise every
candidate is decided again with its exact fp64 mean in visit order.
Property checked: whenever the fast path fires, its winner is the exact fp64 argmin (and no other candidate ties with it).
The quantisation and the fp32 arithmetic are restated in numpy with the kernel's operation order; populations: random means, means
pushed to near-ties of every size around the band, several voxel sizes and keys on both sides of the origin."""
import numpy as np
import pytest

F = np.float32
AXIS_MAX = 8191


def pack_axis(o):
    t = (o + 2.0) * (AXIS_MAX / 4.0) + 0.5
    return np.clip(t, 0.0, float(AXIS_MAX)).astype(np.int64)          # static_cast<uint64_t>: truncation of a non-negative value


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)   # (one rounding, like FFMA)


def kernel_distances(qint, frac32):
    step = F(4.0) / F(AXIS_MAX)
    c = [fma32(qint[..., k].astype(np.float32), np.full(qint.shape[:-1], step, np.float32), np.full(qint.shape[:-1], -2.0, np.float32)) for k in range(3)]
    dx, dy, dz = ((c[k] - frac32[..., k]).astype(np.float32) for k in range(3))
    return fma32(dz, dz, fma32(dy, dy, (dx * dx).astype(np.float32)))


@pytest.mark.parametrize("vs", [1.0, 0.5, 0.37, 2.5])
@pytest.mark.parametrize("key", [(7, 3, 12), (-4, 0, 2), (0, 0, 0), (-1, -6, -3)])
def test_the_fast_path_returns_the_exact_argmin(vs, key):
    rng = np.random.default_rng(abs(hash((vs, key))) % (2 ** 31))
    n_q, n_c = 4000, 27
    key = np.array(key, np.float64)
    frac = rng.random((n_q, 3))                                                             # the query inside its cell, voxel units
    p = (key + frac) * vs                                                                   # the query in metres (fp64)
    off = rng.random((n_q, n_c, 3)) * 3.8 - 1.9                                             # means in (-2, 2) voxel sizes around the key
    # near-ties of every size: some candidates are moved onto (almost) the sphere of candidate 0 around the query
    r0 = np.linalg.norm(off[:, 0] - frac, axis=1)
    for j in range(1, 6):
        d = rng.normal(size=(n_q, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        eps = rng.choice([0.0, 1e-9, 1e-6, 1e-4, 5e-4, 1.1e-3, 1.3e-3, 3e-3], size=n_q) * rng.choice([-1.0, 1.0], size=n_q)
        cand = frac + d * (r0 + eps)[:, None]
        ok = (np.abs(cand) < 1.9).all(axis=1)
        off[ok, j] = cand[ok]
    means = (key + off) * vs                                                                # exact fp64 voxel means in metres
    # what the kernel sees: offsets of the means from the key in voxel sizes, quantised; the in-cell fraction of the query in fp32
    qint = pack_axis((means / vs) - key)
    frac32 = ((p / vs) - np.floor(p / vs)).astype(np.float32)
    d32 = kernel_distances(qint, np.broadcast_to(frac32[:, None, :], qint.shape))
    s = np.sort(d32, axis=1)
    m, s2 = s[:, 0], s[:, 1]
    sd = (np.sqrt(m).astype(np.float32) + F(1.2e-3)).astype(np.float32)
    fast = s2 > ((sd * sd).astype(np.float32) * F(1.00001)).astype(np.float32)
    exact = ((means - p[:, None, :]) ** 2).sum(axis=2)                                       # (fp64; the separation asserted below dwarfs its rounding)
    win = d32.argmin(axis=1)
    order = np.sort(exact, axis=1)
    assert fast.mean() > 0.5 and (~fast).sum() > 50                                         # both paths occur in this population
    assert (exact.argmin(axis=1)[fast] == win[fast]).all()
    # ... and with room to spare: the runner-up is farther by more than the fp64 rounding of this check could hide
    gap = np.sqrt(order[fast, 1]) - np.sqrt(order[fast, 0])
    assert gap.min() > 1e-4 * vs


def test_the_quantisation_error_is_what_the_band_assumes():
    rng = np.random.default_rng(5)
    o = rng.random(200_000) * 4.0 - 2.0
    back = pack_axis(o).astype(np.float64) * (4.0 / AXIS_MAX) - 2.0
    assert np.abs(back - o).max() <= 2.45e-4
    assert np.sqrt(3.0) * 2.45e-4 < 4.3e-4 and 2 * 4.3e-4 < 1.2e-3                          # two candidates, each off by <= 4.3e-4: inside the band
