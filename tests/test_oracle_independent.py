"""An INDEPENDENT restatement of the three correspondence searches — numpy + a plain dict, none of the oracle's code — checked
against the oracle (the reference ships no vectors; besides the pin on the reference's own sources, tests/test_reference_build.py,
the C++ restatement is cross-examined by a second, differently structured one that shares none of its linear algebra).  Follows voxel_hash_map.cpp:31-243 directly:
floor query key (voxel_hash_map.hpp:176-180), GetAdjacentVoxels order (x outer, y, z inner / c,+x,-x,+y,-y,+z,-z), strict <
(first of equals wins), the default-constructed neighbour at the origin when nothing is found, the max-distance gate."""
import numpy as np
import pytest

from elimaloc_b200 import synth
from oracle import oracle as O

P2P, GICP, VGICP, AVGICP = 0, 1, 2, 3


def transform_exact(T, s):
    """(T [s; 1]).head3 in the association order of a 4x4 * 4x1 product: ((T0 x + T1 y) + T2 z) + T3, one rounding per op."""
    x, y, z = (s[:, k].astype(np.float64) for k in range(3))
    return np.stack([((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3] for r in range(3)], axis=1)


def brute_force(export, scan, T, method, max_dist, vs):
    keys, counts, pxyz = export["keys"], export["counts"], export["pxyz"].astype(np.float64)
    starts = np.concatenate([[0], np.cumsum(counts)])
    voxel = {tuple(int(c) for c in k): v for v, k in enumerate(keys)}
    p = transform_exact(np.asarray(T, np.float64), scan)
    K = 7 if method == AVGICP else 1
    cnt = np.zeros(len(scan), np.int32)
    tgt = np.zeros((len(scan), K, 3))
    for i, q in enumerate(p):
        k = np.floor(q / vs).astype(np.int64)
        if method == AVGICP:
            c = 0
            for d in ((0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
                v = voxel.get((k[0] + d[0], k[1] + d[1], k[2] + d[2]))
                if v is None:
                    continue
                m = export["vmean"][v]
                e = m - q
                if (e[0] * e[0] + e[1] * e[1]) + e[2] * e[2] < max_dist * max_dist:
                    tgt[i, c] = m
                    c += 1
            cnt[i] = c
            continue
        cand = []
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    v = voxel.get((k[0] + dx, k[1] + dy, k[2] + dz))
                    if v is None:
                        continue
                    cand.append(export["vmean"][v][None, :] if method == VGICP else pxyz[starts[v]:starts[v + 1]])
        best = np.zeros(3)  # the default-constructed neighbour sits at the origin
        if cand:
            c = np.concatenate(cand)
            e = c - q[None, :]
            d2 = (e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1]) + e[:, 2] * e[:, 2]
            best = c[int(np.argmin(d2))]  # argmin returns the FIRST minimum: the strict < of the reference
        e = best - q
        if (e[0] * e[0] + e[1] * e[1]) + e[2] * e[2] < max_dist * max_dist:
            cnt[i] = 1
            tgt[i, 0] = best
    return cnt, tgt


@pytest.mark.parametrize("origin", [0.0, -6.5])
@pytest.mark.parametrize("method", [P2P, VGICP, AVGICP])
def test_oracle_searches_equal_an_independent_restatement(origin, method):
    raw = synth.map_u(25_000, 13.0, origin=origin)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    om.CalVoxelCovAll()
    ex = om.export()
    T = synth.se3([origin + 6.0, origin + 7.0, origin + 6.5], [0.02, -0.03, 0.4])
    rng = np.random.default_rng(11)
    scan = ((rng.random((500, 3)) * 2 - 1) * 9.0).astype(np.float32)       # inside, at the border of and outside the map
    scan[:40] = (np.linalg.inv(T) @ np.c_[ex["pxyz"][:40].astype(np.float64), np.ones(40)].T).T[:, :3].astype(np.float32)  # near-zero distances
    for max_dist in (5.0, 0.6):
        oc, ot = O.correspondences(om, scan, T, method, max_dist)
        bc, bt = brute_force(ex, scan, T, method, max_dist, 1.0)
        assert np.array_equal(oc, bc), (method, max_dist, np.flatnonzero(oc != bc)[:5])
        assert np.array_equal(ot, bt), (method, max_dist)


def test_oracle_origin_default_is_reproduced_by_the_restatement():
    """Q2: a query whose 27 voxels are empty is matched to the origin when it lies within max_dist of it."""
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(np.array([[40.0, 40.0, 40.0]], np.float32))
    om.CalVoxelCovAll()
    scan = np.array([[1.0, 2.0, 2.0], [4.0, 4.0, 4.0]], np.float32)
    for method in (P2P, VGICP):
        oc, ot = O.correspondences(om, scan, np.eye(4), method, 5.0)
        bc, bt = brute_force(om.export(), scan, np.eye(4), method, 5.0, 1.0)
        assert np.array_equal(oc, bc) and np.array_equal(ot, bt) and list(oc) == [1, 0]


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def linearize_numpy(scan, cnt, tgt, T, method, th, covs=None):
    """AlignCloudsLocal (registration.cpp:15-66) / AlignCloudsLocalVoxelCov (:154-225) written with numpy.linalg — a
    restatement that shares nothing with the oracle's hand-written small-matrix code."""
    Tinv = np.linalg.inv(T)
    Rinv = np.linalg.inv(T[:3, :3])
    JTJ, JTr, res, n = np.zeros((6, 6)), np.zeros(6), 0.0, 0
    for i in range(len(scan)):
        s = scan[i].astype(np.float64)
        for c in range(cnt[i]):
            t = tgt[i, c]
            r = (Tinv @ np.r_[t, 1.0])[:3] - s
            J = np.hstack([np.eye(3), -skew(s)])
            w = th * th / (th + r @ r) ** 2
            n += 1
            if method == P2P:
                JTJ += w * J.T @ J
                JTr += w * J.T @ r
                res += np.linalg.norm(r)
            else:
                if w < 0.01:
                    continue
                M = np.linalg.inv(Rinv @ covs[i][c] @ Rinv.T)
                JTJ += w * J.T @ M @ J
                JTr += w * J.T @ M @ r
                res += np.linalg.norm(r)
    return JTJ, JTr, res, n


@pytest.mark.parametrize("method", [P2P, VGICP, AVGICP])
def test_oracle_linearisation_equals_a_numpy_restatement(method):
    raw = synth.map_s(40_000, 30.0) if method != P2P else synth.map_u(25_000, 13.0, origin=-2.0)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    om.CalVoxelCovAll()
    ex = om.export()
    T_true = synth.se3([6.0, 7.0, 2.0], [0.02, -0.03, 0.4])
    scan = synth.scan_m(ex["pxyz"], 400, T_true)
    T = T_true @ synth.canonical_offset()
    th = 5.0
    cnt, tgt = brute_force(ex, scan, T, method, th, 1.0)
    covs = None
    if method != P2P:  # the covariance that belongs to each emitted voxel mean (means are unique per voxel)
        by_mean = {tuple(m): c for m, c in zip(ex["vmean"], ex["vcov"])}
        covs = [[by_mean.get(tuple(tgt[i, c]), np.eye(3)) for c in range(cnt[i])] for i in range(len(scan))]
    JTJ, JTr, res, n = linearize_numpy(scan, cnt, tgt, T, method, th, covs)
    lin = O.Registration().linearize(scan, om, T, O.make_config(icp_method=method, max_search_dist=th))
    assert lin["n_corr"] == n and n > 300
    assert np.abs(lin["JTJ"] - JTJ).max() <= 1e-10 * np.abs(JTJ).max()
    assert np.abs(lin["JTr"] - JTr).max() <= 1e-10 * max(np.abs(JTr).max(), 1e-300) + 1e-9
    assert abs(lin["residual_sum"] - res) <= 1e-10 * res


def test_oracle_gicp_linearisation_equals_a_numpy_restatement():
    """AlignCloudsLocalPointCov (registration.cpp:68-152): the nearest neighbour is chosen by POSITION, the residual goes to
    that point's neighbourhood MEAN (Q4), weight 0.8 w + 0.2 (Q8), M from the point's covariance, fitness = |r . n| with n the
    eigenvector of the smallest eigenvalue (numpy.linalg.eigh here; the sign of n does not matter)."""
    raw = synth.map_s(40_000, 30.0)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    om.CalVoxelCovAll()
    om.CalPointCovAll(0.4)
    ex = om.export()
    T_true = synth.se3([6.0, 7.0, 2.0], [0.02, -0.03, 0.4])
    scan = synth.scan_m(ex["pxyz"], 400, T_true)
    T = T_true @ synth.canonical_offset()
    th = 5.0
    cnt, tgt = brute_force(ex, scan, T, P2P, th, 1.0)
    index = {tuple(p): i for i, p in enumerate(ex["pxyz"].astype(np.float64))}   # stored positions are unique (spacing filter)
    Tinv, Rinv = np.linalg.inv(T), np.linalg.inv(T[:3, :3])
    JTJ, JTr, res, n = np.zeros((6, 6)), np.zeros(6), 0.0, 0
    for i in range(len(scan)):
        if not cnt[i]:
            continue
        m = index[tuple(tgt[i, 0])]
        s = scan[i].astype(np.float64)
        r = (Tinv @ np.r_[ex["pmean"][m], 1.0])[:3] - s
        J = np.hstack([np.eye(3), -skew(s)])
        w = th * th / (th + r @ r) ** 2 * 0.8 + 0.2
        M = np.linalg.inv(Rinv @ ex["pcov"][m] @ Rinv.T)
        JTJ += w * J.T @ M @ J
        JTr += w * J.T @ M @ r
        nl = Rinv @ np.linalg.eigh(ex["pcov"][m])[1][:, 0]
        res += abs(r @ (nl / np.linalg.norm(nl)))
        n += 1
    lin = O.Registration().linearize(scan, om, T, O.make_config(icp_method=GICP, max_search_dist=th))
    assert lin["n_corr"] == n and n > 300
    assert np.abs(lin["JTJ"] - JTJ).max() <= 1e-10 * np.abs(JTJ).max()
    assert np.abs(lin["JTr"] - JTr).max() <= 1e-10 * np.abs(JTr).max() + 1e-9
    assert abs(lin["residual_sum"] - res) <= 1e-8 * res


def regularize_svd(cov):
    """U diag(1, 1, 1e-3) V^T with numpy's SVD (voxel_hash_map.hpp:141-144); well defined for full-rank covariances"""
    U, s, Vt = np.linalg.svd(cov)
    return U @ np.diag([1.0, 1.0, 1e-3]) @ Vt, s


@pytest.mark.parametrize("origin", [0.0, -5.5])
def test_oracle_map_build_and_covariances_equal_a_python_restatement(origin):
    """AddPoints / AddPointWithSpacing (voxel_hash_map.cpp:270-285, .hpp:106-113) as a dict of lists filled one point at a time;
    CalVoxelCov (.hpp:114-148) and ProcessVoxelBlock (.hpp:195-250, self counted twice) with numpy SVD.  Covariances whose
    sample covariance is (numerically) rank deficient are excluded: their null-space basis is implementation-defined."""
    vs, cap = 1.0, 30
    raw = synth.map_u(12_000, 9.0, origin=origin)
    om = O.VoxelHashMap(vs, cap)
    om.AddPoints(raw)
    om.CalVoxelCovAll()
    om.CalPointCovAll(0.4)
    ex = om.export()
    # ---- sequential insert
    res = np.sqrt(vs * vs / cap)
    vox = {}
    for p in raw.astype(np.float64):
        key = tuple(int(c) for c in (p / vs))                       # static_cast<int>: truncation toward zero
        pts = vox.setdefault(key, [])
        if not pts:
            pts.append(p)                                           # the first point of a voxel is always kept
        elif len(pts) < cap and all(np.sqrt(((q - p) ** 2)[0] + ((q - p) ** 2)[1] + ((q - p) ** 2)[2]) >= res for q in pts):
            pts.append(p)
    keys = sorted(vox)
    assert np.array_equal(ex["keys"], np.array(keys, np.int32))
    assert np.array_equal(ex["counts"], np.array([len(vox[k]) for k in keys], np.int32))
    assert np.array_equal(ex["pxyz"].astype(np.float64), np.concatenate([np.array(vox[k]) for k in keys]))
    # ---- voxel covariances
    checked = 0
    for v, k in enumerate(keys):
        P = np.array(vox[k])
        if len(P) == 1:
            assert np.array_equal(ex["vmean"][v], P[0]) and np.array_equal(ex["vcov"][v], np.eye(3))
            continue
        mean = P.mean(axis=0)
        cov, s = regularize_svd((P - mean).T @ (P - mean) / (len(P) - 1))
        assert np.allclose(ex["vmean"][v], mean, rtol=0, atol=1e-12)
        if s[2] > 1e-6 * s[0] and s[1] - s[2] > 1e-6 * s[0]:
            assert np.abs(ex["vcov"][v] - cov).max() < 1e-8
            checked += 1
    assert checked > 0.5 * len(keys)
    # ---- point covariances (a sample of points): neighbours = self + every stored point of the 27 FLOOR-keyed voxels within 0.4 m
    starts = np.concatenate([[0], np.cumsum(ex["counts"])])
    stored = ex["pxyz"].astype(np.float64)
    index = {k: v for v, k in enumerate(keys)}
    checked = 0
    for p in range(0, len(stored), 37):
        x = stored[p]
        kf = np.floor(x / vs).astype(int)
        nb = [x]
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    v = index.get((kf[0] + dx, kf[1] + dy, kf[2] + dz))
                    if v is None:
                        continue
                    for q in stored[starts[v]:starts[v + 1]]:
                        e = q - x
                        if (e[0] * e[0] + e[1] * e[1]) + e[2] * e[2] <= 0.4 * 0.4:
                            nb.append(q)
        N = np.array(nb)
        mean = N.mean(axis=0)
        cov, s = regularize_svd((N - mean).T @ (N - mean) / (len(N) - 1))
        assert np.allclose(ex["pmean"][p], mean, rtol=0, atol=1e-12)
        if s[2] > 1e-6 * s[0] and s[1] - s[2] > 1e-6 * s[0]:
            assert np.abs(ex["pcov"][p] - cov).max() < 1e-8
            checked += 1
    assert checked > 20


def rodrigues(w):
    """AngleAxisd(|w|, w / |w|).toRotationMatrix() (registration.cpp:58-62)"""
    a = np.linalg.norm(w)
    if a == 0.0:
        return np.eye(3)
    k = skew(w / a)
    return np.eye(3) + np.sin(a) * k + (1.0 - np.cos(a)) * (k @ k)


@pytest.mark.parametrize("method", [P2P, VGICP])
def test_oracle_run_register_equals_an_independent_icp_loop(method):
    """RunRegister (registration.cpp:274-418) end to end, rebuilt from the independent pieces above: search, linearise,
    x = (JtJ + lambda diag(JtJ))^-1 Jtr (numpy.linalg.solve instead of LDLT), right-multiplied update T <- T dT with
    x = [t; w], termination test AFTER the update, fitness = residual sum / pairs.  Six iterations from the canonical offset."""
    raw = synth.map_s(30_000, 24.0) if method == VGICP else synth.map_u(25_000, 13.0, origin=-2.0)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    om.CalVoxelCovAll()
    ex = om.export()
    T_true = synth.se3([6.0, 7.0, 2.0], [0.02, -0.03, 0.4])
    scan = synth.scan_m(ex["pxyz"], 300, T_true)
    T = T_true @ synth.canonical_offset()
    th, lam, thr, iters = 5.0, 0.5, 1e-4, 6
    by_mean = {tuple(m): c for m, c in zip(ex["vmean"], ex["vcov"])}
    fitness, done_at = None, iters
    Tk = T.copy()
    for it in range(iters):
        cnt, tgt = brute_force(ex, scan, Tk, method, th, 1.0)
        covs = [[by_mean.get(tuple(tgt[i, c]), np.eye(3)) for c in range(cnt[i])] for i in range(len(scan))] if method == VGICP else None
        JTJ, JTr, res, n = linearize_numpy(scan, cnt, tgt, Tk, method, th, covs)
        fitness = res / n
        x = np.linalg.solve(JTJ + lam * np.diag(np.diag(JTJ)), JTr)
        dT = np.eye(4)
        dT[:3, :3] = rodrigues(x[3:])
        dT[:3, 3] = x[:3]
        Tk = Tk @ dT
        if np.linalg.norm(x[3:]) + np.linalg.norm(x[:3]) < thr:   # angle of dR == |w| for |w| < pi
            done_at = it + 1
            break
    cfg = O.make_config(icp_method=method, max_iteration=iters, max_search_dist=th, lm_lambda=lam, icp_termination_threshold_m=thr,
                        min_overlap_ratio=0.0, max_fitness_score=1e30)
    o = O.Registration().RunRegister(scan, om, T, cfg)
    assert o["is_success"]
    assert np.abs(o["pose"] - Tk).max() <= 1e-9 * np.abs(Tk).max()
    assert abs(o["fitness_score"] - fitness) <= 1e-9 * fitness
    # ... and the loop really moves towards the truth (LM damping 0.5: about a third of the error is removed per iteration)
    assert np.linalg.norm(Tk[:3, 3] - T_true[:3, 3]) < 0.5 * np.linalg.norm(T[:3, 3] - T_true[:3, 3])


def test_oracle_deskew_equals_a_numpy_restatement():
    """DeskewPoint / FindRotation / FindPosition (pcm_matching.cpp:731-824) in float64 numpy: table lookup with linear
    interpolation, odometry ratio, the Q3 quirk (z translation takes the interpolated yaw integral), Rz Ry Rx.  The oracle
    works in float32 like the reference, so the comparison allows float32 rounding of 80 m coordinates."""
    rng = np.random.default_rng(5)
    t_end, span = 500.0, 0.1
    t_cur = t_end - span
    stamps = np.arange(t_cur - 0.04, t_end + 0.04, 0.005)
    gyro = np.tile([0.08, -0.05, 0.7], (len(stamps), 1)) + rng.normal(0, 0.01, (len(stamps), 3))
    start = np.array([3.0, -2.0, 0.5, 0.02, -0.01, 0.3])
    end = start + np.array([0.9, 0.05, 0.02, 0.008, -0.005, 0.07])
    tab = O.deskew_tables(stamps, gyro, t_cur, t_end, start, t_cur, end, t_end)
    xyz = ((rng.random((3000, 3), dtype=np.float32) * 2 - 1) * np.float32(80.0)).astype(np.float32)
    rel = np.sort(rng.random(3000).astype(np.float32) * np.float32(span))
    got = O.deskew_points(tab, xyz, rel)

    cur = tab["imu_pointer_cur"]
    tt, rx, ry, rz = (np.asarray(tab[k], np.float64) for k in ("imu_time", "imu_rot_x", "imu_rot_y", "imu_rot_z"))
    inc = np.asarray(tab["odom_incre"], np.float64)
    want = np.zeros((len(xyz), 3))
    for i, (p, dt) in enumerate(zip(xyz.astype(np.float64), rel.astype(np.float64))):
        t = tab["time_scan_cur"] + dt
        front = 0
        while front < cur and not t < tt[front]:
            front += 1
        if t > tt[front] or front == 0:
            rot = np.array([rx[front], ry[front], rz[front]])
        else:
            back = front - 1
            a = (t - tt[back]) / (tt[front] - tt[back])
            b = (tt[front] - t) / (tt[front] - tt[back])
            rot = np.array([rx[front] * a + rx[back] * b, ry[front] * a + ry[back] * b, rz[front] * a + rz[back] * b])
        pos = dt / (tab["time_scan_end"] - tab["time_scan_cur"]) * inc if tab["odom_available"] else np.zeros(3)
        d_rot = rot - np.array([rx[cur], ry[cur], rz[cur]])
        d_pos = np.array([pos[0] - inc[0], pos[1] - inc[1], rot[2] - inc[2]])   # Q3: rot z, not pos z (pcm_matching.cpp:804)
        R = synth.exp_so3([0, 0, d_rot[2]]) @ synth.exp_so3([0, d_rot[1], 0]) @ synth.exp_so3([d_rot[0], 0, 0])
        want[i] = R @ p + d_pos
    assert np.abs(got - want).max() < 1e-4          # float32 arithmetic on |x| <= 80 m
    assert np.abs(got - xyz).max() > 0.05           # the correction is not a no-op


def test_oracle_imu_prediction_equals_a_numpy_restatement():
    """EkfAlgorithm::RunPredictionImu (ekf_algorithm.cpp:167-316) for one step from a random state, complementary filter off:
    strap-down propagation with bias and gravity, Q (diagonal blocks x dt^2), F (identity + sparse blocks incl. the gravity
    column and -dExp/dgyro), P <- F P F^T + Q — written with numpy, compared with the oracle's hand-rolled 27x27 code."""
    from elimaloc_b200 import _capi, ekf as pekf
    rng = np.random.default_rng(42)
    cfg = pekf.make_ekf_config(use_complementary_filter=0)
    f = O.EkfAlgorithm(cfg, _capi.EkfState)
    N = 27
    A = rng.normal(size=(N, N))
    P0 = A @ A.T / N + np.eye(N) * 0.05
    R0 = synth.exp_so3([0.1, -0.2, 0.7])
    w = np.sqrt(1 + np.trace(R0)) / 2
    q0 = np.array([w, (R0[2, 1] - R0[1, 2]) / (4 * w), (R0[0, 2] - R0[2, 0]) / (4 * w), (R0[1, 0] - R0[0, 1]) / (4 * w)])
    st = dict(pos=[1.0, -2.0, 0.5], vel=[7.0, 0.3, -0.1], bg=[0.002, -0.001, 0.003], ba=[0.05, -0.02, 0.01], grav=[0.0, 0.0, 9.79])
    for k, v in st.items():
        getattr(f.s, k)[:] = v
    f.s.rot[:] = list(q0)
    f.s.P[:] = list(P0.reshape(-1))
    f.s.state_initialized = 1
    f.s.reset_for_init_prediction = 0
    f.s.pcm_init_on_going = 0
    t0, dt = 50.0, 0.01
    f.s.prev_timestamp = t0
    gyro, acc = np.array([0.02, -0.01, 0.3]), np.array([0.4, 2.1, 9.9])
    assert f.RunPredictionImu(t0 + dt, gyro, acc)

    bg, ba, grav, pos, vel = (np.array(st[k]) for k in ("bg", "ba", "grav", "pos", "vel"))
    wg = gyro - bg
    R1 = R0 @ rodrigues(wg * dt)
    a_g = R0 @ (acc - ba) - grav
    pos1 = pos + vel * dt + 0.5 * a_g * dt * dt
    vel1 = vel + a_g * dt
    Q = np.zeros((N, N))
    d2r = np.pi / 180.0
    for idx, sd in ((0, cfg.state_std_pos_m), (3, cfg.state_std_rot_deg * d2r), (6, cfg.state_std_vel_mps), (9, cfg.imu_std_gyro_dps * d2r),
                    (12, cfg.imu_std_acc_mps), (15, cfg.imu_bias_cov_gyro), (18, cfg.imu_bias_cov_acc), (21, cfg.imu_bias_cov_acc),
                    (24, cfg.state_std_rot_deg * d2r)):
        Q[idx:idx + 3, idx:idx + 3] = np.eye(3) * sd ** 2 * dt * dt
    F = np.eye(N)
    F[0:3, 6:9] = np.eye(3) * dt
    F[0:3, 18:21] = -0.5 * R0 * dt * dt
    om = wg * dt
    th = np.linalg.norm(om)
    K = skew(om / th)
    F[3:6, 15:18] = -dt * (np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * K @ K)
    F[6:9, 18:21] = -R0 * dt
    F[9:12, 15:18] = -np.eye(3)
    F[12:15, 18:21] = -R0
    F[2, 23], F[8, 23], F[14, 23] = -0.5 * dt * dt, -dt, -1.0      # imu_estimate_gravity: z column only
    P1 = F @ P0 @ F.T + Q

    s = pekf.state_to_dict(f.s)
    qw, qx, qy, qz = s["rot"]
    Rq = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                   [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                   [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])
    np.testing.assert_allclose(Rq, R1, rtol=0, atol=1e-12)
    np.testing.assert_allclose(s["pos"], pos1, rtol=1e-13)
    np.testing.assert_allclose(s["vel"], vel1, rtol=1e-13)
    np.testing.assert_allclose(s["gyro"], wg, rtol=1e-13)
    np.testing.assert_allclose(s["acc"], a_g, rtol=1e-12)
    np.testing.assert_allclose(np.array(s["P"]).reshape(N, N), P1, rtol=1e-11, atol=1e-14)
    assert s["prev_timestamp"] == t0 + dt and s["predictions"] == 1


def quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def R_to_quat(R):
    w = np.sqrt(1 + np.trace(R)) / 2
    return np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])


def rot_to_vec(R):
    """RotToVec (localization_functions.hpp:312-333), non-gimbal branch: roll, pitch, yaw"""
    pitch = np.arcsin(-R[2, 0])
    return np.array([np.arctan2(R[2, 1] / np.cos(pitch), R[2, 2] / np.cos(pitch)), pitch, np.arctan2(R[1, 0] / np.cos(pitch), R[0, 0] / np.cos(pitch))])


def test_oracle_pcm_update_equals_a_numpy_restatement():
    """EkfAlgorithm::RunGnssUpdate for a PCM measurement (ekf_algorithm.cpp:366-428) + UpdateEkfState (ekf_algorithm.hpp:116-145):
    Y = [dpos; Euler(meas) - Euler(state)] (CalEulerResidualFromQuat), K = P H^T (H P H^T + R)^-1, EVERY state block updated from
    K Y, rot and imu_rot right-multiplied by the angle-axis quaternions of their slices, P <- P - K H P."""
    from elimaloc_b200 import _capi, ekf as pekf
    rng = np.random.default_rng(9)
    N = 27
    f = O.EkfAlgorithm(pekf.make_ekf_config(use_complementary_filter=0), _capi.EkfState)
    A = rng.normal(size=(N, N))
    P0 = A @ A.T / N + np.eye(N) * 0.05
    R0 = synth.exp_so3([0.03, -0.05, 0.8])
    Ri = synth.exp_so3([0.01, 0.02, -0.03])
    st = dict(pos=[4.0, -1.0, 0.2], vel=[6.0, 0.5, 0.0], gyro=[0.0, 0.01, 0.2], acc=[0.1, 0.2, 0.0], bg=[1e-3, 2e-3, -1e-3],
              ba=[0.02, 0.0, -0.01], grav=[0.0, 0.0, 9.8])
    for k, v in st.items():
        getattr(f.s, k)[:] = v
    f.s.rot[:] = list(R_to_quat(R0))
    f.s.imu_rot[:] = list(R_to_quat(Ri))
    f.s.P[:] = list(P0.reshape(-1))
    f.s.state_initialized = 1
    Rm = synth.exp_so3([0.0, 0.0, 0.83]) @ synth.exp_so3([0.0, -0.04, 0.0]) @ synth.exp_so3([0.02, 0.0, 0.0])
    zpos = np.array([4.2, -1.1, 0.25])
    R6 = np.diag([0.04, 0.05, 0.09, 2e-4, 1e-4, 4e-4])
    assert f.RunGnssUpdate(pekf.make_measurement(7.0, zpos, R_to_quat(Rm), R6[:3, :3], R6[3:, 3:], source=pekf.PCM))

    H = np.zeros((6, N))
    H[:6, :6] = np.eye(6)
    K = P0 @ H.T @ np.linalg.inv(H @ P0 @ H.T + R6)
    res = rot_to_vec(Rm) - rot_to_vec(R0)
    res = (res + np.pi) % (2 * np.pi) - np.pi
    Y = np.r_[zpos - np.array(st["pos"]), res]
    du = K @ Y
    s = pekf.state_to_dict(f.s)
    for name, lo in (("pos", 0), ("vel", 6), ("gyro", 9), ("acc", 12), ("bg", 15), ("ba", 18), ("grav", 21)):
        np.testing.assert_allclose(s[name], np.array(st[name]) + du[lo:lo + 3], rtol=1e-10, atol=1e-13, err_msg=name)
    np.testing.assert_allclose(quat_to_R(s["rot"]), R0 @ rodrigues(du[3:6]), rtol=0, atol=1e-12)
    np.testing.assert_allclose(quat_to_R(s["imu_rot"]), Ri @ rodrigues(du[24:27]), rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.array(s["P"]).reshape(N, N), P0 - K @ H @ P0, rtol=1e-10, atol=1e-13)


def predict_numpy(st, P0, gyro, acc, dt, cfg, N=27):
    """the strap-down step + F / Q of test_oracle_imu_prediction..., as a function: returns (new state dict, P1)"""
    d2r = np.pi / 180.0
    R0 = st["R"]
    wg = gyro - st["bg"]
    a_g = R0 @ (acc - st["ba"]) - st["grav"]
    new = dict(st, R=R0 @ rodrigues(wg * dt), pos=st["pos"] + st["vel"] * dt + 0.5 * a_g * dt * dt, vel=st["vel"] + a_g * dt, gyro=wg, acc=a_g)
    Q = np.zeros((N, N))
    for idx, sd in ((0, cfg.state_std_pos_m), (3, cfg.state_std_rot_deg * d2r), (6, cfg.state_std_vel_mps), (9, cfg.imu_std_gyro_dps * d2r),
                    (12, cfg.imu_std_acc_mps), (15, cfg.imu_bias_cov_gyro), (18, cfg.imu_bias_cov_acc), (21, cfg.imu_bias_cov_acc),
                    (24, cfg.state_std_rot_deg * d2r)):
        Q[idx:idx + 3, idx:idx + 3] = np.eye(3) * sd ** 2 * dt * dt
    F = np.eye(N)
    F[0:3, 6:9] = np.eye(3) * dt
    F[0:3, 18:21] = -0.5 * R0 * dt * dt
    om = wg * dt
    th = np.linalg.norm(om)
    if th >= 1e-5:
        K = skew(om / th)
        F[3:6, 15:18] = -dt * (np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * K @ K)
    F[6:9, 18:21] = -R0 * dt
    F[9:12, 15:18] = -np.eye(3)
    F[12:15, 18:21] = -R0
    F[2, 23], F[8, 23], F[14, 23] = -0.5 * dt * dt, -dt, -1.0
    return new, F @ P0 @ F.T + Q


def apply_update(st, P, K, Y, H):
    """UpdateEkfState (ekf_algorithm.hpp:116-145)"""
    du = K @ Y
    out = dict(st)
    for name, lo in (("pos", 0), ("vel", 6), ("gyro", 9), ("acc", 12), ("bg", 15), ("ba", 18), ("grav", 21)):
        out[name] = st[name] + du[lo:lo + 3]
    out["R"] = st["R"] @ rodrigues(du[3:6])
    out["Ri"] = st["Ri"] @ rodrigues(du[24:27])
    return out, P - K @ H @ P


@pytest.mark.parametrize("p_scale", [1.0, 1e-7])
def test_oracle_complementary_filter_equals_a_numpy_restatement(p_scale):
    """ComplementaryKalmanFilter (ekf_algorithm.cpp:597-701) inside RunPredictionImu: roll / pitch pseudo-measurement from the
    bias-corrected accelerometer with the centripetal (v_x * yaw rate) and — once the rotation is stabilised (P small) — the
    longitudinal (d v_x / dt, function-static memory) accelerations removed, noise scaled by the dynamics, 2-row update.
    Two predictions: the first only latches the statics (dt = 0), the second applies the update."""
    from elimaloc_b200 import _capi, ekf as pekf
    rng = np.random.default_rng(4)
    N = 27
    cfg = pekf.make_ekf_config(use_complementary_filter=1)
    f = O.EkfAlgorithm(cfg, _capi.EkfState)
    A = rng.normal(size=(N, N))
    P = (A @ A.T / N + np.eye(N) * 0.05) * p_scale
    st = dict(R=synth.exp_so3([0.02, -0.03, 0.5]), Ri=np.eye(3), pos=np.array([1.0, 2.0, 0.3]), vel=np.array([7.0, 1.0, 0.05]),
              gyro=np.zeros(3), acc=np.zeros(3), bg=np.array([1e-3, -2e-3, 5e-4]), ba=np.array([0.03, -0.01, 0.02]),
              grav=np.array([0.0, 0.0, 9.80]))
    for k in ("pos", "vel", "bg", "ba", "grav"):
        getattr(f.s, k)[:] = list(st[k])
    f.s.rot[:] = list(R_to_quat(st["R"]))
    f.s.P[:] = list(P.reshape(-1))
    f.s.state_initialized = 1
    f.s.reset_for_init_prediction = 0
    f.s.pcm_init_on_going = 0
    t0, dt = 20.0, 0.01
    f.s.prev_timestamp = t0
    imu = [(np.array([0.01, -0.02, 0.25]), np.array([0.9, 1.9, 9.7])), (np.array([0.012, -0.018, 0.26]), np.array([1.1, 2.0, 9.75]))]
    prev = None
    for k, (gyro, acc) in enumerate(imu):
        t = t0 + (k + 1) * dt
        stabilised = all(np.sqrt(P[i, i]) < 0.2 * np.pi / 180 for i in (3, 4, 5))     # CheckRotationStabilized, before the step
        assert f.RunPredictionImu(t, gyro, acc)
        st, P = predict_numpy(st, P, gyro, acc, dt, cfg)
        # ---- ComplementaryKalmanFilter
        a_meas = acc - st["ba"]
        v_local = st["R"].T @ st["vel"]
        centrip = v_local[0] * st["gyro"][2]
        if prev is None:
            prev = (v_local[0], t)
            continue                                                                   # first call: dt == 0 -> return
        est_ax = (v_local[0] - prev[0]) / (t - prev[1])
        prev = (v_local[0], t)
        comp = a_meas - np.array([0.0, centrip, 0.0])
        if stabilised:
            comp = comp - np.array([est_ax, 0.0, 0.0])
        gdir = comp / np.linalg.norm(comp)
        z = np.array([np.arctan2(gdir[1], gdir[2]), -np.arcsin(gdir[0])])
        innov = z - rot_to_vec(st["R"])[:2]
        innov = (innov + np.pi) % (2 * np.pi) - np.pi
        H = np.zeros((2, N))
        H[0, 3] = H[1, 4] = 1.0
        base = 1.0 * np.pi / 180
        diff = abs(np.linalg.norm(a_meas) - np.linalg.norm(st["grav"])) / 9.81 * 10
        lat, lon = 1 + diff + abs(centrip) / 9.81 * 10, 1 + diff + abs(est_ax) / 9.81 * 10
        Rm = np.diag([max((base * lat) ** 2, base ** 2), max((base * lon) ** 2, base ** 2)])
        K = P @ H.T @ np.linalg.inv(H @ P @ H.T + Rm)
        st, P = apply_update(st, P, K, innov, H)
    s = pekf.state_to_dict(f.s)
    np.testing.assert_allclose(quat_to_R(s["rot"]), st["R"], rtol=0, atol=1e-11)
    for name in ("pos", "vel", "gyro", "acc", "bg", "ba", "grav"):
        np.testing.assert_allclose(s[name], st[name], rtol=1e-9, atol=1e-12, err_msg=name)
    np.testing.assert_allclose(np.array(s["P"]).reshape(N, N), P, rtol=1e-9, atol=1e-16 * p_scale + 1e-20)
    assert s["ckf_has_prev"] == 1 and s["predictions"] == 2
