"""Empirical check of the exact-pruning bound of the P2P/GICP search (elimaloc_b200/csrc/icp_kernels.cu, `axis_gap2` /
`voxels_to_visit`): a voxel of the 27-neighbourhood is skipped only if  0.9999 (gx^2 + gy^2 + gz^2) > float_ru(best_d2) / vs^2 * 1.00001,
g = per-axis gap (fp32, voxel units, minus 1e-5) between the query's in-cell fraction and the interval of coordinates that can
be STORED under the voxel's key — insert keys truncate toward zero (voxel_hash_map.cpp:275): [c, c+1) for c > 0, (c-1, c] for
c < 0, (-1, 1) for c == 0.  Property: a skipped voxel cannot hold a point with exact squared distance <= best_d2.
The kernel's float32 formulas are restated in numpy; stored points are drawn inside every neighbour voxel with emphasis on the
faces, edges and corners nearest to the query, for several voxel sizes and for keys around zero (where the asymmetry lives)."""
import numpy as np
import pytest

F = np.float32


def axis_gap2(kq, o, f):
    c = kq + o
    lo = F(o - (1 if c <= 0 else 0))
    hi = F(o + (1 if c >= 0 else 0))
    g = max(max(max(F(lo - f), F(f - hi)), F(0.0)) - F(1e-5), F(0.0))
    return F(F(g) * F(g))


def stored_interval(c, vs):
    """closed hull of the coordinates whose truncated key is c"""
    if c > 0:
        return c * vs, (c + 1) * vs
    if c < 0:
        return (c - 1) * vs, c * vs
    return -vs, vs


@pytest.mark.parametrize("vs", [1.0, 0.5, 0.3, 12.0])
def test_a_skipped_voxel_cannot_hold_a_closer_point(vs):
    rng = np.random.default_rng(int(vs * 100))
    inv_vs2_up = F(F(1.0 / (vs * vs)) * F(1.00001))
    violations, skipped_total = 0, 0
    for _ in range(400):
        k = rng.integers(-2, 3, size=3)                                   # query cell around the origin (both signs, zero)
        p = (k + rng.random(3)) * vs
        if rng.random() < 0.3:                                            # queries sitting (almost) on cell faces
            ax = rng.integers(0, 3)
            p[ax] = (k[ax] + rng.choice([0.0, 1e-12, 1 - 1e-12, 0.5])) * vs
        q = p / vs
        kq = np.floor(q).astype(int)
        f = (q - np.floor(q)).astype(F)
        best_d2 = (rng.choice([0.02, 0.2, 0.6, 1.1]) * vs * rng.random()) ** 2
        bound = F(np.nextafter(F(best_d2), F(np.inf)) if F(best_d2) < best_d2 else F(best_d2)) * inv_vs2_up   # __double2float_ru
        g = [[axis_gap2(int(kq[a]), o, f[a]) for o in (-1, 0, 1)] for a in range(3)]
        for L in range(27):
            o = (L // 9 - 1, (L // 3) % 3 - 1, L % 3 - 1)
            lb = F(F(F(g[0][o[0] + 1] + g[1][o[1] + 1]) + g[2][o[2] + 1]) * F(0.9999))
            if not lb > bound:
                continue                                                  # voxel visited: nothing to prove
            skipped_total += 1
            # the closest storable position of that voxel, plus random ones, as float32 stored points
            c = kq + np.array(o)
            lo_hi = [stored_interval(int(c[a]), vs) for a in range(3)]
            nearest = np.array([min(max(p[a], lo_hi[a][0]), lo_hi[a][1]) for a in range(3)])
            cands = [nearest] + [np.array([rng.uniform(*lo_hi[a]) for a in range(3)]) for _ in range(6)]
            for s in cands:
                s32 = s.astype(F).astype(np.float64)
                # keep only float32 points whose truncated key really is c (rounding to float32 can cross a face)
                if not np.array_equal((s32 / vs).astype(np.int64), c):    # astype(int64) truncates toward zero like static_cast<int>
                    continue
                d2 = float(((s32 - p) ** 2).sum())
                if d2 <= best_d2:
                    violations += 1
    assert skipped_total > 1000 and violations == 0


def test_the_bound_does_prune():
    """sanity: with a best distance of a tenth of a voxel, most of the 26 other voxels are skipped for a query in mid-cell"""
    vs = 1.0
    inv = F(F(1.0) * F(1.00001))
    f = np.array([0.5, 0.5, 0.5], F)
    g = [[axis_gap2(5, o, f[a]) for o in (-1, 0, 1)] for a in range(3)]
    bound = F(0.01) * inv
    need = sum(not (F(F(F(g[0][L // 9] + g[1][(L // 3) % 3]) + g[2][L % 3]) * F(0.9999)) > bound) for L in range(27))
    assert need == 1
