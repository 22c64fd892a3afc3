"""The pin of the oracle: oracle/liboracle.so (our CPU restatement) against oracle/_ref/libref.so — the REFERENCE's own
registration.cpp + voxel_hash_map.{hpp,cpp}, compiled unmodified from /root/reference against stand-in third-party headers
(oracle/ref_build/stubs; Eigen3 / oneTBB / PCL are absent from this image).  Runs without a GPU.

What agreement here means: the reference's control flow — insertion keys (truncation, Q1) vs query keys (floor), the spacing
test, visit order and first-wins tie-breaks, the default-constructed "origin" neighbour (Q2), the three searches and their
emission order, weights, the AVGICP weight cut, the LM step, early-outs, fitness persistence — is what the oracle restates.
Index-level results must be identical; floating-point results agree to rounding (the stand-in evaluates Eigen expressions
eagerly, the oracle has its own summation order; Eigen's decomposition kernels are restated in both, oracle/smallmat.hpp).

In the build container the library is (re)built from the sources; on a box without /root/reference the prebuilt file is used;
with neither the module is skipped."""
import numpy as np
import pytest

from elimaloc_b200 import synth
from oracle import oracle as O
from oracle import reference_build as R

pytestmark = pytest.mark.skipif(not R.available(), reason="neither /root/reference nor a prebuilt oracle/_ref/libref.so is here")

METHODS = [O.P2P, O.GICP, O.VGICP, O.AVGICP]


def build_both(raw, voxel_size=1.0, cap=30, cov=True, radius=0.4):
    maps = []
    for impl in (O, R):
        m = impl.VoxelHashMap(voxel_size, cap)
        m.AddPoints(raw)
        if cov:
            m.CalVoxelCovAll()
            m.CalPointCovAll(radius)
        maps.append(m)
    return maps


def assert_same_map(om, rm, cov=True):
    eo, er = om.export(), rm.export()
    for k in ("keys", "counts", "pxyz"):
        assert np.array_equal(eo[k], er[k]), k
    if cov:
        for k in ("vmean", "pmean"):
            assert np.array_equal(eo[k], er[k]), k
        for k in ("vcov", "pcov"):
            assert np.abs(eo[k] - er[k]).max() < 1e-12, k


@pytest.fixture(scope="module")
def world():
    raw = synth.map_u(60_000, 18.0, origin=-7.0)  # straddles the origin: truncation vs floor keys differ on the negative side
    om, rm = build_both(raw)
    T_true = synth.se3([2.0, 1.5, 1.0], [0.02, -0.01, 0.3])
    scan = synth.scan_m(om.export()["pxyz"], 2048, T_true)
    return dict(om=om, rm=rm, scan=scan, T_true=T_true, T0=T_true @ synth.canonical_offset())


def test_map_build_identical(world):
    assert_same_map(world["om"], world["rm"])
    assert world["om"].num_voxels() == world["rm"].num_voxels() and world["om"].num_points() == world["rm"].num_points()


@pytest.mark.parametrize("voxel_size,cap,n,box", [(0.5, 30, 20_000, 6.0), (1.0, 5, 30_000, 5.0), (2.0, 30, 20_000, 12.0), (1.0, 1, 5_000, 6.0)])
def test_map_build_other_shapes(voxel_size, cap, n, box):
    """other voxel sizes, a cap that fills up (cap 5 / cap 1 in a dense box), incremental AddPoints calls"""
    raw = synth.map_u(n, box, seed=77, origin=-box / 2)
    om, rm = build_both(raw[: n // 2], voxel_size, cap, cov=False)
    om.AddPoints(raw[n // 2:])
    rm.AddPoints(raw[n // 2:])
    for m in (om, rm):
        m.CalVoxelCovAll()
        m.CalPointCovAll(0.4 * voxel_size)
    assert_same_map(om, rm)


def test_insert_key_truncates_and_query_key_floors():
    """Q1: AddPoints keys by static_cast<int>(p / vs) (truncation), the searches by floor: a point at x = -0.5 is stored under
    key 0 and is therefore invisible to a query at x = -1.6 (query key -2: neighbourhood -3..-1)."""
    pts = np.array([[-0.5, 0.2, 0.2], [3.5, 0.2, 0.2]], np.float32)
    om, rm = build_both(pts, cov=False)
    assert np.array_equal(om.export()["keys"], rm.export()["keys"])
    assert np.array_equal(rm.export()["keys"], np.array([[0, 0, 0], [3, 0, 0]], np.int32))
    q = np.array([[-1.6, 0.2, 0.2]], np.float32)
    for impl, m in ((O, om), (R, rm)):
        cnt, tgt = impl.correspondences(m, q, np.eye(4), O.P2P, 5.0)
        assert cnt[0] == 1 and np.array_equal(tgt[0, 0], [0.0, 0.0, 0.0])  # the default-constructed neighbour (Q2), not the point


@pytest.mark.parametrize("method", METHODS)
def test_correspondences_identical(world, method):
    for pose in (world["T0"], world["T_true"], np.eye(4)):
        co, to = O.correspondences(world["om"], world["scan"], pose, method, 5.0)
        cr, tr = R.correspondences(world["rm"], world["scan"], pose, method, 5.0)
        assert np.array_equal(co, cr) and np.array_equal(to, tr)


@pytest.mark.parametrize("method", METHODS)
def test_correspondences_of_a_scan_that_leaves_the_map(world, method):
    """queries outside the map: empty neighbourhoods fall back to the default neighbour at the origin (Q2), accepted when the
    query lies within max_dist of (0, 0, 0)"""
    q = synth.scan_u(3000, 30.0, seed=5)
    co, to = O.correspondences(world["om"], q, np.eye(4), method, 5.0)
    cr, tr = R.correspondences(world["rm"], q, np.eye(4), method, 5.0)
    assert np.array_equal(co, cr) and np.array_equal(to, tr)
    assert 0 < co.sum() < len(q) * (7 if method == O.AVGICP else 1)
    for max_dist in (0.3, 1.7, 12.0):  # larger than the 27-voxel reach: the search still only sees 27 voxels
        co, to = O.correspondences(world["om"], q[:500], np.eye(4), method, max_dist)
        cr, tr = R.correspondences(world["rm"], q[:500], np.eye(4), method, max_dist)
        assert np.array_equal(co, cr) and np.array_equal(to, tr)


@pytest.mark.parametrize("method", [O.P2P, O.VGICP])
def test_equidistant_candidates_first_visited_wins(method):
    """a lattice map and queries at cell centres / face centres: many exactly equidistant candidates; the strict `<` of the
    reference keeps the first one in visit order (x outer, y, z inner; insertion order inside a voxel)"""
    g = np.arange(-3, 4, dtype=np.float32)
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + np.float32(0.25)
    rng = np.random.default_rng(3)
    lattice = lattice[rng.permutation(len(lattice))]
    extra = lattice + np.float32(0.5)  # second point in most voxels: insertion order matters
    om, rm = build_both(np.vstack([lattice, extra]))
    q = np.array([[0.5, 0.5, 0.5], [0.0, 0.0, 0.0], [0.75, 0.25, 0.25], [-0.25, -0.25, 0.5], [1.0, 1.0, 1.0], [0.25, 0.75, -1.25]], np.float32)
    co, to = O.correspondences(om, q, np.eye(4), method, 5.0)
    cr, tr = R.correspondences(rm, q, np.eye(4), method, 5.0)
    assert np.array_equal(co, cr) and np.array_equal(to, tr)


@pytest.mark.parametrize("method", METHODS)
def test_whole_scan_emission_order(world, method):
    """one search over the whole scan: pairs come out in scan order (AVGICP: up to 7 per point, in the order centre, +x, -x,
    +y, -y, +z, -z), for 1 and for 4 threads of the oneTBB stand-in"""
    co, to = O.correspondences(world["om"], world["scan"], world["T0"], method, 5.0)
    want_idx = np.repeat(np.arange(len(co)), co)
    want_tgt = np.concatenate([to[i, :c] for i, c in enumerate(co)]) if co.sum() else np.zeros((0, 3))
    for threads in (1, 4):
        R.set_threads(threads)
        try:
            idx, tgt = R.search_pairs(world["rm"], world["scan"], world["T0"], method, 5.0)
        finally:
            R.set_threads(1)
        assert np.array_equal(idx, want_idx) and np.array_equal(tgt, want_tgt)


@pytest.mark.parametrize("method", METHODS)
def test_linearisation_agrees(world, method):
    """JTJ / JTr / residual sum of one AlignClouds* call (read from the system the reference hands to ldlt().solve with
    lm_lambda = 0)"""
    for pose in (world["T0"], world["T_true"]):
        cfg = O.make_config(icp_method=method)
        lo = O.Registration().linearize(world["scan"], world["om"], pose, cfg)
        lr = R.Registration().linearize(world["scan"], world["rm"], pose, cfg)
        assert lo["n_corr"] == lr["n_corr"] > 0
        assert np.abs(lo["JTJ"] - lr["JTJ"]).max() <= 1e-11 * np.abs(lr["JTJ"]).max()
        assert np.abs(lo["JTr"] - lr["JTr"]).max() <= 1e-11 * np.abs(lr["JTr"]).max()
        assert abs(lo["residual_sum"] - lr["residual_sum"]) <= 1e-12 * abs(lr["residual_sum"])


def compare_runs(ro, rr, lm_lambda=0.5, pose_tol=1e-10):
    assert ro["is_success"] == rr["is_success"]
    assert ro["n_iter"] == rr["n_iter"]
    assert np.abs(ro["pose"] - rr["pose"]).max() < pose_tol
    if np.isnan(rr["fitness_score"]):
        assert np.isnan(ro["fitness_score"])
    else:
        assert abs(ro["fitness_score"] - rr["fitness_score"]) <= 1e-11 * max(1.0, abs(rr["fitness_score"]))
    fin = np.isfinite(rr["local_cov"])  # GICP on an empty scan inverts the zero matrix: the non-finite pattern must match too
    assert np.array_equal(fin, np.isfinite(ro["local_cov"]))
    if fin.any():
        assert np.abs(ro["local_cov"][fin] - rr["local_cov"][fin]).max() <= 1e-9 * max(1.0, np.abs(rr["local_cov"][fin]).max())
    for k in range(rr["n_iter"]):  # per iteration: the reference's regularised system vs the oracle's raw sums
        A = ro["trace"]["JTJ"][k] + lm_lambda * np.diag(np.diag(ro["trace"]["JTJ"][k]))
        assert np.abs(A - rr["trace"]["A"][k]).max() <= 1e-9 * np.abs(A).max(), k
        # Jtr cancels to ~0 at convergence while its terms stay as large as the entries of JtJ: the rounding scale is |A|, not |b|
        assert np.abs(ro["trace"]["JTr"][k] - rr["trace"]["b"][k]).max() <= 1e-9 * max(1.0, np.abs(rr["trace"]["b"][k]).max()) + 1e-12 * np.abs(A).max(), k


@pytest.mark.parametrize("method", METHODS)
def test_run_register_forced_iterations(world, method):
    cfg = O.make_config(icp_method=method, max_iteration=10, **synth.timing_knobs())
    compare_runs(O.Registration().RunRegister(world["scan"], world["om"], world["T0"], cfg),
                 R.Registration().RunRegister(world["scan"], world["rm"], world["T0"], cfg))


@pytest.mark.parametrize("method", METHODS)
def test_run_register_default_ini_config(world, method):
    """the reference's own knobs (config/localization.ini): termination threshold, overlap gate, fitness gate"""
    cfg = O.make_config(icp_method=method)
    ro = O.Registration().RunRegister(world["scan"], world["om"], world["T0"], cfg, fitness_in=-1.0)
    rr = R.Registration().RunRegister(world["scan"], world["rm"], world["T0"], cfg, fitness_in=-1.0)
    compare_runs(ro, rr)
    if method in (O.P2P, O.GICP):
        assert rr["is_success"] and rr["n_iter"] < 10  # converged before the iteration cap
        err = np.linalg.inv(world["T_true"]) @ rr["pose"]
        assert np.linalg.norm(err[:3, 3]) < 0.02


def test_early_outs(world):
    scan, T0 = world["scan"], world["T0"]
    # empty map: failure, pose = initial guess, fitness argument untouched
    eo, er = O.VoxelHashMap(1.0, 30), R.VoxelHashMap(1.0, 30)
    assert eo.Empty() and er.Empty()
    cfg = O.make_config(icp_method=O.GICP)
    compare_runs(O.Registration().RunRegister(scan, eo, T0, cfg, fitness_in=7.0), R.Registration().RunRegister(scan, er, T0, cfg, fitness_in=7.0))
    rr = R.Registration().RunRegister(scan, er, T0, cfg, fitness_in=7.0)
    assert not rr["is_success"] and np.array_equal(rr["pose"], T0) and rr["fitness_score"] == 7.0 and rr["n_iter"] == 0
    # overlap gate: a scan far from the map
    far = scan + np.float32(500.0)
    for method in METHODS:
        cfg = O.make_config(icp_method=method)
        ro = O.Registration().RunRegister(far, world["om"], T0, cfg, fitness_in=3.0)
        rr = R.Registration().RunRegister(far, world["rm"], T0, cfg, fitness_in=3.0)
        compare_runs(ro, rr)
        assert not rr["is_success"] and rr["fitness_score"] == 3.0
    # fitness gate: every iteration runs, then the result is rejected; the pose returned is the refined one
    for method in METHODS:
        cfg = O.make_config(icp_method=method, max_fitness_score=1e-6)
        ro = O.Registration().RunRegister(scan, world["om"], T0, cfg, fitness_in=3.0)
        rr = R.Registration().RunRegister(scan, world["rm"], T0, cfg, fitness_in=3.0)
        compare_runs(ro, rr)
        assert not rr["is_success"] and rr["fitness_score"] == 3.0 and not np.array_equal(rr["pose"], T0)
    # max_iteration = 0: no iteration, success with the initial pose (d_fitness_score_ = 0 passes the gate)
    cfg = O.make_config(icp_method=O.P2P, max_iteration=0)
    compare_runs(O.Registration().RunRegister(scan, world["om"], T0, cfg), R.Registration().RunRegister(scan, world["rm"], T0, cfg))


def test_fitness_score_member_persists_across_calls(world):
    """d_fitness_score_ is a member (registration.hpp:229): a call that runs no iteration is gated on the PREVIOUS call's value"""
    oreg, rreg = O.Registration(), R.Registration()
    cfg = O.make_config(icp_method=O.VGICP, max_iteration=3, **synth.timing_knobs())
    compare_runs(oreg.RunRegister(world["scan"], world["om"], world["T0"], cfg), rreg.RunRegister(world["scan"], world["rm"], world["T0"], cfg))
    cfg0 = O.make_config(icp_method=O.P2P, max_iteration=0, max_fitness_score=0.1)  # VGICP's fitness above was ~0.5
    ro = oreg.RunRegister(world["scan"], world["om"], world["T0"], cfg0, fitness_in=9.0)
    rr = rreg.RunRegister(world["scan"], world["rm"], world["T0"], cfg0, fitness_in=9.0)
    compare_runs(ro, rr)
    assert not rr["is_success"] and rr["fitness_score"] == 9.0


def test_empty_scan(world):
    """0 points: 0/0 overlap ratio is NaN, which passes `ratio < min` — the loop runs on empty sums"""
    empty = np.zeros((0, 3), np.float32)
    for method in METHODS:
        cfg = O.make_config(icp_method=method)
        compare_runs(O.Registration().RunRegister(empty, world["om"], world["T0"], cfg, fitness_in=2.0),
                     R.Registration().RunRegister(empty, world["rm"], world["T0"], cfg, fitness_in=2.0))


def test_surface_map_all_methods():
    """Map-S (ground + walls): the geometry VGICP / AVGICP are made for; converging runs"""
    raw = synth.map_s(60_000, 30.0)
    om, rm = build_both(raw)
    assert_same_map(om, rm)
    T_true = synth.se3([12.0, 14.0, 1.5], [0.01, -0.02, 0.5])
    scan = synth.scan_m(om.export()["pxyz"], 3000, T_true)
    T0 = T_true @ synth.canonical_offset()
    for method in METHODS:
        cfg = O.make_config(icp_method=method, max_iteration=15)
        compare_runs(O.Registration().RunRegister(scan, om, T0, cfg), R.Registration().RunRegister(scan, rm, T0, cfg), pose_tol=1e-9)


def test_find_ground_height(world):
    for xy in ([0.0, 0.0], [3.3, -2.2], [-6.5, 8.0], [40.0, 40.0]):
        fo, zo = world["om"].FindGroundHeight(xy)
        fr, zr = world["rm"].FindGroundHeight(xy)
        assert fo == fr
        if fr:
            assert abs(zo - zr) < 1e-12
    assert not world["rm"].FindGroundHeight([40.0, 40.0])[0]


def test_voxel_downsample_keeps_the_first_point_of_every_floor_keyed_voxel():
    """the reference emits the survivors in hash-table order; as a set they are the first point of every voxel"""
    rng = np.random.default_rng(9)
    xyz = ((rng.random((20_000, 3)) - 0.5) * 30).astype(np.float32)
    for vs in (0.5, 1.0, 2.5):
        ref_idx = R.voxel_downsample(xyz, vs)
        assert len(np.unique(ref_idx)) == len(ref_idx)
        assert np.array_equal(np.sort(ref_idx), O.scan_preprocess(xyz, 0.0, vs))


def test_threads_of_the_tbb_stand_in_do_not_change_results(world):
    raw = synth.map_u(20_000, 10.0, seed=11, origin=-5.0)
    _, r1 = build_both(raw)
    R.set_threads(4)
    try:
        _, r4 = build_both(raw)
        cfg = O.make_config(icp_method=O.GICP)
        a = R.Registration().RunRegister(world["scan"], world["rm"], world["T0"], cfg)
    finally:
        R.set_threads(1)
    b = R.Registration().RunRegister(world["scan"], world["rm"], world["T0"], cfg)
    e1, e4 = r1.export(), r4.export()
    assert all(np.array_equal(e1[k], e4[k]) for k in e1)
    assert np.array_equal(a["pose"], b["pose"]) and a["n_iter"] == b["n_iter"]


@pytest.mark.parametrize("voxel_size,cap,n,box,radius", [(1.0, 30, 60_000, 18.0, 0.4), (0.5, 12, 30_000, 8.0, 0.3), (2.0, 5, 40_000, 16.0, 0.8)])
def test_product_host_map_builder_against_the_reference_sources(voxel_size, cap, n, box, radius):
    """the PRODUCT's map builder (elimaloc_b200/csrc/host_map.cpp: slab-parallel counting sort, no hash map; device = -1 keeps it
    on the host, no GPU needed) directly against the reference's AddPoints / CalVoxelCovAll / CalPointCovAll: same voxels, same
    stored points in the same per-voxel order, covariances to 1e-9 — without the oracle in between"""
    import elimaloc_b200 as E
    raw = synth.map_u(n, box, seed=21, origin=-0.4 * box)
    pm, rm = E.VoxelHashMap(voxel_size, cap, device=-1), R.VoxelHashMap(voxel_size, cap)
    for m in (pm, rm):
        m.AddPoints(raw[: n // 3])
        m.AddPoints(raw[n // 3:])
        m.CalVoxelCovAll()
        m.CalPointCovAll(radius)
    pe, re_ = pm.export(True, True), rm.export()
    for k in ("keys", "counts", "pxyz"):
        assert np.array_equal(pe[k], re_[k]), k
    for k in ("vmean", "vcov", "pmean", "pcov"):
        assert np.abs(pe[k] - re_[k]).max() < 1e-9, k
    rng = np.random.default_rng(2)
    for q in rng.uniform(-0.5 * box, 0.7 * box, (20, 2)):
        pf, pz = pm.FindGroundHeight(q)
        rf, rz = rm.FindGroundHeight(q)
        assert pf == rf and (not rf or abs(pz - rz) < 1e-12)


@pytest.mark.parametrize("seed", range(6))
def test_randomised_worlds(seed):
    """small random worlds: random voxel size, cap, covariance radius, map extent (straddling the origin or far from it), pose and
    max_dist; all four methods; oracle and reference sources must agree on everything"""
    rng = np.random.default_rng(1000 + seed)
    vs = float(rng.choice([0.4, 0.75, 1.0, 1.6]))
    cap = int(rng.choice([3, 10, 30]))
    box = float(rng.uniform(6, 14)) * vs
    origin = float(rng.choice([-0.5 * box, -3.0, 120.0, -250.0]))
    raw = synth.map_u(int(rng.integers(5_000, 25_000)), box, seed=seed, origin=origin)
    om, rm = build_both(raw, vs, cap, radius=0.4 * vs)
    assert_same_map(om, rm)
    c = origin + 0.5 * box
    T_true = synth.se3([c + rng.uniform(-1, 1), c + rng.uniform(-1, 1), c + rng.uniform(-0.5, 0.5)], rng.normal(0, 0.2, 3))
    scan = synth.scan_m(om.export()["pxyz"], 600, T_true, noise=0.03 * vs, seed=seed)
    T0 = T_true @ synth.se3(rng.normal(0, 0.15 * vs, 3), rng.normal(0, 0.01, 3))
    md = float(rng.choice([0.8, 2.0, 5.0])) * vs
    for method in METHODS:
        co, to = O.correspondences(om, scan, T0, method, md)
        cr, tr = R.correspondences(rm, scan, T0, method, md)
        assert np.array_equal(co, cr) and np.array_equal(to, tr)
        cfg = O.make_config(icp_method=method, max_search_dist=md, max_iteration=int(rng.integers(1, 12)), lm_lambda=float(rng.choice([0.0, 0.1, 0.5])))
        compare_runs(O.Registration().RunRegister(scan, om, T0, cfg, fitness_in=-3.0), R.Registration().RunRegister(scan, rm, T0, cfg, fitness_in=-3.0),
                     lm_lambda=cfg.lm_lambda, pose_tol=1e-9)


@pytest.mark.parametrize("seed", [0, 5, 20, 45, 50, 95, 105, 110])
def test_product_host_map_builder_on_lattices_and_duplicates(seed):
    """adversarial maps for the covariance passes: exact duplicates and points on a quarter-voxel lattice (exactly collinear /
    coplanar neighbourhoods, equal eigenvalues, dominant directions along lattice diagonals).  The rank-deficient cases follow the
    documented convention — and must do so independently of the last bits of the eigenvectors: an earlier
    version of the axis choice flipped on lattice diagonals (1 point in 1437), found by this sweep."""
    import elimaloc_b200 as E
    rng = np.random.default_rng(seed)
    vs = float(rng.choice([0.3, 0.5, 1.0, 2.0, 3.7]))
    cap = int(rng.choice([1, 2, 5, 12, 30, 100]))
    n = int(rng.integers(200, 30000))
    box = float(rng.uniform(2, 20)) * vs
    origin = float(rng.choice([-0.5 * box, -box, 0.0, 777.0, -12345.0]))
    raw = synth.map_u(n, box, seed=seed, origin=origin)
    if seed % 4 == 0:
        raw = np.vstack([raw, raw[: n // 3]])
    if seed % 5 == 0:
        raw = np.round(raw / (vs / 4)).astype(np.float32) * np.float32(vs / 4)
    maps = [E.VoxelHashMap(vs, cap, device=-1), R.VoxelHashMap(vs, cap), O.VoxelHashMap(vs, cap)]
    k = int(rng.integers(1, 4))
    rad = float(rng.uniform(0.2, 0.9)) * vs
    for m in maps:
        for part in np.array_split(raw, k):
            m.AddPoints(part)
        m.CalVoxelCovAll()
        m.CalPointCovAll(rad)
    pe, re_, oe = maps[0].export(True, True), maps[1].export(), maps[2].export()
    for other in (re_, oe):
        for key in ("keys", "counts", "pxyz"):
            assert np.array_equal(pe[key], other[key]), key
        for key in ("vmean", "vcov", "pmean", "pcov"):
            assert np.abs(pe[key] - other[key]).max() < 1e-9, key


def test_searches_on_lattice_maps_and_lattice_queries():
    """exact ties everywhere: maps and queries on half- / quarter-voxel lattices (random insertion order, caps that fill up,
    boxes on both sides of the origin, lattice-preserving poses), all four methods, two gate distances"""
    total = 0
    for seed in range(24):
        rng = np.random.default_rng(3000 + seed)
        vs = float(rng.choice([0.5, 1.0, 2.0]))
        cap = int(rng.choice([2, 8, 30]))
        step = vs / float(rng.choice([2, 4]))
        n = int(rng.integers(500, 6000))
        box = float(rng.uniform(3, 9)) * vs
        origin = float(rng.choice([-0.5 * box, -box - 0.25, 0.0]))
        raw = np.round(synth.map_u(n, box, seed=seed, origin=origin) / step).astype(np.float32) * np.float32(step)
        raw = raw[rng.permutation(len(raw))]
        om, rm = build_both(raw, vs, cap, radius=0.6 * vs)
        q = np.round(synth.scan_u(400, 0.7 * box, seed=seed) / (step / 2)).astype(np.float32) * np.float32(step / 2) + np.float32(origin + 0.5 * box)
        pose = synth.se3([step, -step, 0.0], [0, 0, np.pi / 2]) if seed % 3 == 0 else np.eye(4)
        for method in METHODS:
            for md in (0.7 * vs, 5.0):
                co, to = O.correspondences(om, q, pose, method, md)
                cr, tr = R.correspondences(rm, q, pose, method, md)
                assert np.array_equal(co, cr) and np.array_equal(to, tr), (seed, method, md)
                total += 1
    assert total == 24 * 4 * 2
