"""Model check of the exactness argument behind the warm-started P2P / GICP search (elimaloc_b200/csrc/icp_kernels.cu:
`warm_refresh`, `warm_list_answers`) — the rule that lets iterations 2..n of a call answer a nearest-neighbour query
from the query's own candidate list instead of searching the 27 voxels again:

  refresh at q0   the previous match p bounds the nearest distance: R = (|p - q0| + margin)(1 + 1e-9); the LIST = every stored point of
                  the 27 voxels of key(q0) within R of q0; the memo keeps fp32(q0), Rf = R - |q0 - fp32(q0)| (rounded down) and key(q0)
  reuse at q      allowed iff a previous match exists and  d' + |q - fp32(q0)| <= Rf  with d' = |p - q| (each side rounded against us), and
                  either key(q) == key(q0), or both keys are >= 2 on every axis and Rf < 0.99 voxel (insert keys truncate toward zero
                  — voxel_hash_map.cpp:275 — so only there "within one voxel size" implies "inside the 27 voxels")
  claim           whenever reuse is allowed, the nearest list entry (exact distance, smallest canonical rank among equals) IS what
                  GetCorrespondencePoints returns at q: the nearest stored point of the 27 voxels of key(q) (voxel_hash_map.cpp:31-88).

The rule is restated in numpy (float32 where the kernel uses float32) and run against a brute-force search over the truncation-keyed
voxels on random maps — positive, negative and straddling the origin — along random walks from sub-millimetre steps to jumps across
voxels.  The GPU tests (tests/test_gpu_warm.py) check the kernels; this checks the argument itself, where no GPU is needed."""
import numpy as np
import pytest

F = np.float32
MARGIN_VOX = 0.08  # elm_registration::warm_margin_vox


def sq3(a, b):
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    return (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]


class World:
    def __init__(self, pts, vs):
        self.vs = vs
        self.pts = np.asarray(pts, np.float32)
        keys = np.trunc(self.pts.astype(np.float64) / vs).astype(np.int64)          # insert key: truncation toward zero
        order = np.lexsort((np.arange(len(keys)), keys[:, 2], keys[:, 1], keys[:, 0]))  # canonical: voxels by (x, y, z), then arrival
        self.rank = np.empty(len(keys), np.int64)
        self.rank[order] = np.arange(len(keys))
        self.vox = {}
        for i, k in enumerate(map(tuple, keys)):
            self.vox.setdefault(k, []).append(i)

    def key(self, q):
        return tuple(int(v) for v in np.floor(np.asarray(q, np.float64) / self.vs))   # query key: floor

    def neighbourhood(self, k):
        out = []
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    out += self.vox.get((k[0] + dx, k[1] + dy, k[2] + dz), [])
        return out

    def nearest(self, q, cand):
        """exact nearest of `cand` (indices), smallest canonical rank among equal distances; None if empty"""
        best, bd, br = None, None, None
        for i in cand:
            d = sq3(self.pts[i], q)
            if best is None or d < bd or (d == bd and self.rank[i] < br):
                best, bd, br = i, d, self.rank[i]
        return best


def refresh(w, q0, prev):
    """-> (match, list, memo) as warm_refresh builds them; prev = index of the previous match"""
    k0 = w.key(q0)
    cand = w.neighbourhood(k0)
    prev_ok = prev is not None and all(abs(int(np.trunc(float(w.pts[prev][a]) / w.vs)) - k0[a]) <= 1 for a in range(3))
    match = w.nearest(q0, cand)
    if not prev_ok:
        return match, None, None
    R = (np.sqrt(sq3(w.pts[prev], q0)) + MARGIN_VOX * w.vs) * (1.0 + 1e-9)
    lst = [i for i in cand if sq3(w.pts[i], q0) <= R * R]
    q0f = np.asarray(q0, np.float64).astype(np.float32)
    e = np.asarray(q0, np.float64) - q0f.astype(np.float64)
    Rf = np.nextafter(F(R - np.sqrt(e @ e) * (1.0 + 1e-9)), F(-np.inf))          # (rounded down, one ulp to spare)
    return match, lst, dict(q0f=q0f, Rf=Rf, key=k0)


def reuse_allowed(w, q, prev, memo):
    if prev is None or memo is None:
        return False
    k = w.key(q)
    if k != memo["key"]:
        if not (min(k) >= 2 and min(memo["key"]) >= 2 and memo["Rf"] < F(0.99) * F(w.vs)):
            return False
    d2_prev = sq3(w.pts[prev], q)
    d = (np.asarray(q, np.float64) - memo["q0f"].astype(np.float64)).astype(np.float32)
    delta = F(np.sqrt(F(d[2] * d[2] + F(d[1] * d[1] + F(d[0] * d[0]))))) * F(1.000001) + F(1e-30)
    room = F(F(memo["Rf"] - delta) * F(0.999999))
    return bool(room > 0 and d2_prev <= float(room) * float(room))


@pytest.mark.parametrize("origin, vs", [(4.0, 1.0), (-9.0, 1.0), (-2.5, 1.0), (3.0, 0.5), (-1.7, 0.37)])
def test_a_reused_list_gives_the_reference_answer(origin, vs):
    rng = np.random.default_rng(int(abs(origin) * 100 + vs * 1000))
    box = 5.0 * vs
    pts = (origin + rng.random((1500, 3)) * box).astype(np.float32)
    pts[:40] = pts[40:80]                                                              # exact duplicates: ties decided by rank
    w = World(pts, vs)
    reused_same, reused_moved, refreshed = 0, 0, 0
    for walk in range(60):
        q = origin + rng.random(3) * box
        match = w.nearest(q, w.neighbourhood(w.key(q)))                                # the cold search of iteration 1
        lst = memo = None
        step = vs * 10.0 ** rng.uniform(-4, 0.3)                                       # 0.1 mm ... 2 voxels per iteration
        for it in range(12):
            q = q + rng.standard_normal(3) * step
            want = w.nearest(q, w.neighbourhood(w.key(q)))                             # GetCorrespondencePoints at q
            if reuse_allowed(w, q, match, memo):
                got = w.nearest(q, lst)
                assert got == want, (origin, vs, walk, it)
                reused_same += w.key(q) == memo["key"]
                reused_moved += w.key(q) != memo["key"]
                match = got
            else:
                match, lst, memo = refresh(w, q, match)
                assert match == want
                refreshed += 1
    assert reused_same > 50 and refreshed > 50                                         # both paths were exercised
    if origin >= 2.0 * vs:
        assert reused_moved > 0                                                        # lists that survived a change of voxel
    else:
        assert reused_moved == 0 or origin + box > 2.0 * vs                            # never among keys < 2


def test_without_the_key_condition_the_argument_fails_near_the_origin():
    """Why the `keys >= 2` condition is there: among negative coordinates the stored key of a point is one higher than its floor key,
    so a point within d' of the query can lie OUTSIDE the 27 voxels the reference visits, and a list carried across a change of voxel
    would answer with a point GetCorrespondencePoints never sees.  Constructed directly on the x axis."""
    vs = 1.0
    w = World(np.array([[-0.9, 0.5, 0.5], [-1.6, 0.5, 0.5]], np.float32), vs)   # stored keys (truncation): 0 and -1
    q0, q = np.array([-0.95, 0.5, 0.5]), np.array([-1.05, 0.5, 0.5])             # floor keys -1 and -2: the query crosses x = -1
    assert w.nearest(q0, w.neighbourhood(w.key(q0))) == 0                        # key -1 visits stored keys -2..0: the point 0.05 away
    match, lst, memo = refresh(w, q0, 0)
    assert match == 0 and sorted(lst) == [0] and memo["Rf"] < 0.99               # (a small R: the margin only)
    want = w.nearest(q, w.neighbourhood(w.key(q)))
    assert want == 1                                                             # key -2 visits stored keys -3..-1: NOT the point 0.15 away
    assert w.nearest(q, lst) == 0 != want                                        # the carried list would give the wrong answer ...
    assert not reuse_allowed(w, q, match, memo)                                  # ... and the rule refuses it (keys < 2)
    memo_far = dict(memo, key=(5, 5, 5))
    assert not reuse_allowed(w, q, match, memo_far)                              # (either key below 2 is enough)
