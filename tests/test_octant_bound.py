"""Empirical check of the OCTANT selection of the warm-started search's refresh path (elimaloc_b200/csrc/icp_kernels.cu:
`interval_gap2`, `axis_halves`, `octant_runs`; voxel_key.hpp `axis_half`): a refresh reads, of the query's 27 voxels, only the octants
(half-voxel cells; the points of a voxel are stored sorted by octant) that can hold a point within R of the query — per axis a half of
a voxel is needed unless its under-estimated fp32 gap to the query exceeds the bound, a voxel unless the sum of its three span gaps
does, a z-column unless its two xy gaps plus the smallest z gap do.  bound = float_ru(R^2) / vs^2 * 1.00001.
Property: a stored point within R (exact, fp64) of the query is never left out: its half is selected on every axis, its voxel and its
column pass.  The spans follow the truncation rule of the insert key (voxel_hash_map.cpp:275): a voxel with stored key c holds
p / vs in [c, c+1) for c > 0, (c-1, c] for c < 0, (-1, 1) for c == 0, and its upper half starts at the middle of that span.
The kernel's float32 formulas are restated in numpy; points are drawn with emphasis on the faces and the middles of the spans."""
import numpy as np
import pytest

F = np.float32


def interval_gap2(f, a, b):
    g = max(max(max(F(a - f), F(f - b)), F(0.0)) - F(1e-5), F(0.0))
    return F(F(g) * F(g))


def axis_halves(kq, f, bound):
    """-> (mask: bit 2 (o + 1) + h, gaps of the three whole spans)"""
    m, gv = 0, []
    for o in (-1, 0, 1):
        c = kq + o
        lo, hi = F(o - (1 if c <= 0 else 0)), F(o + (1 if c >= 0 else 0))
        mid = F(F(0.5) * F(lo + hi))
        g0, g1 = interval_gap2(f, lo, mid), interval_gap2(f, mid, hi)
        gv.append(min(g0, g1))
        if not F(g0 * F(0.9999)) > bound:
            m |= 1 << (2 * (o + 1))
        if not F(g1 * F(0.9999)) > bound:
            m |= 2 << (2 * (o + 1))
    return m, gv


def axis_half(q, c):
    mid = c + 0.5 if c > 0 else (c - 0.5 if c < 0 else 0.0)
    return 1 if q >= mid else 0


def span(c):
    return (c, c + 1) if c > 0 else ((c - 1, c) if c < 0 else (-1, 1))


@pytest.mark.parametrize("vs", [1.0, 0.5, 0.3, 12.0])
def test_a_point_within_R_is_never_left_out(vs):
    rng = np.random.default_rng(int(vs * 1000))
    inv_vs2_up = F(F(1.0 / (vs * vs)) * F(1.00001))
    checked, left_out_possible = 0, 0
    for _ in range(600):
        k = rng.integers(-2, 3, size=3)                                     # query cells around the origin: both signs, zero
        q = k + rng.random(3)
        if rng.random() < 0.3:
            ax = rng.integers(0, 3)
            q[ax] = k[ax] + rng.choice([0.0, 1e-12, 1 - 1e-12, 0.5, 0.25])
        p = q * vs
        kq = np.floor(p / vs).astype(int)
        f = (p / vs - np.floor(p / vs)).astype(F)
        R = rng.choice([0.02, 0.15, 0.4, 0.9, 1.6]) * vs * rng.random()
        r2 = R * R
        bound = F(np.nextafter(F(r2), F(np.inf)) if F(r2) < r2 else F(r2)) * inv_vs2_up        # __double2float_ru(R * R) * inv_vs2_up
        masks, gaps = zip(*(axis_halves(int(kq[a]), f[a], bound) for a in range(3)))
        for L in range(27):
            o = np.array((L // 9 - 1, (L // 3) % 3 - 1, L % 3 - 1))
            c = kq + o
            # stored points of that voxel: the closest storable position, the middles and faces of the spans, random ones
            lo_hi = [np.array(span(int(c[a])), np.float64) * vs for a in range(3)]
            nearest = np.array([min(max(p[a], lo_hi[a][0]), lo_hi[a][1]) for a in range(3)])
            cands = [nearest] + [np.array([rng.uniform(*lo_hi[a]) for a in range(3)]) for _ in range(4)]
            mids = np.array([0.5 * (lo_hi[a][0] + lo_hi[a][1]) for a in range(3)])
            cands += [np.where(rng.random(3) < 0.5, mids + rng.choice([0.0, 1e-7, -1e-7]) * vs, nearest)]
            for s in cands:
                s32 = s.astype(F).astype(np.float64)
                if not np.array_equal((s32 / vs).astype(np.int64), c):      # (rounding to float32 crossed a face: another voxel's point)
                    continue
                if float(((s32 - p) ** 2).sum()) > r2:
                    left_out_possible += 1
                    continue                                                # farther than R: may be skipped, nothing to prove
                checked += 1
                for a in range(3):
                    h = axis_half(s32[a] / vs, int(c[a]))
                    assert masks[a] & ((1 << h) << (2 * (o[a] + 1))), ("half not selected", vs, kq, o, a, s32, p, R)
                voxel_lb = F(F(F(gaps[0][o[0] + 1] + gaps[1][o[1] + 1]) + gaps[2][o[2] + 1]) * F(0.9999))
                assert not voxel_lb > bound, ("voxel skipped", vs, kq, o, s32, p, R)
                col_lb = F(F(F(gaps[0][o[0] + 1] + gaps[1][o[1] + 1]) + min(gaps[2])) * F(0.9999))
                assert not col_lb > bound, ("column skipped", vs, kq, o, s32, p, R)
    assert checked > 2000 and left_out_possible > 2000


def test_the_selection_does_prune():
    """sanity: a bound of a tenth of a voxel around a query in the lower corner region of its cell needs one octant of 27 x 8"""
    f = np.array([0.25, 0.25, 0.25], F)
    bound = F(0.01) * F(1.00001)
    masks = [axis_halves(5, f[a], bound)[0] for a in range(3)]
    assert masks == [0b000100, 0b000100, 0b000100]                          # only the lower half of the centre voxel, on every axis
