"""GPU edge cases and size-independent properties of the hot path (through the C ABI)."""
import numpy as np
import pytest

import elimaloc_b200 as E
from elimaloc_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
I4 = np.eye(4)


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def small():
    raw = synth.map_u(30_000, 12.0, origin=-4.0)
    gm = E.VoxelHashMap(1.0, 30, device=0)
    gm.AddPoints(raw)
    gm.CalVoxelCovAll()
    gm.CalPointCovAll(0.4)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    om.CalVoxelCovAll()
    om.CalPointCovAll(0.4)
    T_true = synth.se3([1.0, 2.0, 1.5], [0.01, -0.02, 0.2])
    return dict(gm=gm, om=om, T_true=T_true, T0=T_true @ synth.canonical_offset(), greg=E.Registration(device=0),
                oreg=O.Registration(), stored=gm.Pointcloud())


@pytest.mark.parametrize("n", [1, 3, 31, 255, 257, 1001, 4099])
@pytest.mark.parametrize("method", [E.P2P, E.GICP, E.VGICP, E.AVGICP])
def test_ragged_scan_sizes(small, n, method):
    """sizes that are not a multiple of the tile / of 4 points (the TMA bulk copy needs 16-byte multiples: the last
    tile falls back to plain loads)"""
    scan = synth.scan_m(small["stored"], n, small["T_true"], seed=100 + n)
    kw = dict(icp_method=method, **synth.timing_knobs())
    g = small["greg"].linearize(scan, small["gm"], small["T0"], E.RegistrationConfig(**kw))
    o = small["oreg"].linearize(scan, small["om"], small["T0"], O.make_config(**kw))
    assert g["n_corr"] == o["n_corr"]
    assert rel_err(g["JTJ"], o["JTJ"]) < 1e-5 and rel_err(g["JTr"], o["JTr"]) < 1e-5
    gc, gt = small["greg"].correspondences(scan, small["gm"], small["T0"], method, 5.0)
    oc, ot = O.correspondences(small["om"], scan, small["T0"], method, 5.0)
    assert np.array_equal(gc, oc) and np.array_equal(gt, ot)


def test_empty_scan_and_empty_map(small):
    cfg = E.RegistrationConfig(icp_method=E.P2P)
    T, ok, fit, cov = small["greg"].RunRegister(np.zeros((0, 3), np.float32), small["gm"], small["T0"], cfg, fitness_score=-3.0)
    assert not ok and np.array_equal(T, small["T0"]) and fit == -3.0 and np.array_equal(cov, np.eye(6))
    empty = E.VoxelHashMap(1.0, 30, device=0)
    scan = synth.scan_m(small["stored"], 64, small["T_true"])
    T, ok, fit, cov = small["greg"].RunRegister(scan, empty, small["T0"], cfg, fitness_score=-3.0)  # registration.cpp:291-295
    assert not ok and np.array_equal(T, small["T0"]) and fit == -3.0 and np.array_equal(cov, np.eye(6))


def test_early_outs_match_the_reference_semantics(small):
    scan = synth.scan_m(small["stored"], 512, small["T_true"])
    g, o = small["greg"], small["oreg"]
    # overlap gate (registration.cpp:351-356)
    far = scan + np.float32(500.0)
    T, ok, fit, _ = g.RunRegister(far, small["gm"], small["T0"], E.RegistrationConfig(icp_method=E.P2P), fitness_score=-3.0)
    assert not ok and np.array_equal(T, small["T0"]) and fit == -3.0
    # fitness gate (registration.cpp:405-409): final pose returned, fitness untouched
    Tg, okg, _, _ = g.RunRegister(scan, small["gm"], small["T0"], E.RegistrationConfig(icp_method=E.P2P))
    Tb, okb, fitb, _ = g.RunRegister(scan, small["gm"], small["T0"], E.RegistrationConfig(icp_method=E.P2P, max_fitness_score=1e-6),
                                     fitness_score=-3.0)
    assert okg and not okb and fitb == -3.0 and rel_err(Tb, Tg) < 1e-12
    # d_fitness_score_ persists across calls (registration.hpp:229): max_iteration = 0
    Tz, okz, fitz, _ = g.RunRegister(scan, small["gm"], small["T0"], E.RegistrationConfig(icp_method=E.P2P, max_iteration=0))
    oz = o.RunRegister(scan, small["om"], small["T0"], O.make_config(icp_method=O.P2P))
    assert okz and np.array_equal(Tz, small["T0"]) and abs(fitz - oz["fitness_score"]) < 1e-9
    # unsupported / misuse
    with pytest.raises(E.ElmError) as ei:
        g.RunRegister(scan, small["gm"], small["T0"], E.RegistrationConfig(icp_method=E.P2P, use_radar_cov=1))
    assert ei.value.status == 4
    bare = E.VoxelHashMap(1.0, 30, device=0)
    bare.AddPoints(synth.map_u(1000, 4.0))
    with pytest.raises(E.ElmError) as ei:
        g.RunRegister(scan, bare, small["T0"], E.RegistrationConfig(icp_method=E.GICP))
    assert ei.value.status == 6


def test_quirks_on_the_gpu():
    """Q1 (trunc insert / floor query), Q2 (origin default), Q6 (7-voxel AVGICP), Q7 (weight skip) through the kernels"""
    g, o = E.Registration(device=0), O.Registration()
    cases = [
        (np.array([[-0.5, -0.5, -0.5], [-1.5, 0.5, 2.5]], np.float32), 1.0, np.array([[-0.4, -0.4, -0.4], [-1.2, 0.2, 2.2]], np.float32), 5.0),
        (np.array([[50.0, 50.0, 50.0]], np.float32), 1.0, np.array([[1.0, 2.0, 2.0], [4.0, 4.0, 4.0]], np.float32), 5.0),
        (np.array([[5.5, 5.5, 5.5], [6.5, 5.5, 5.5], [4.5, 5.5, 5.5], [5.5, 6.5, 5.5], [5.5, 4.5, 5.5], [5.5, 5.5, 6.5], [5.5, 5.5, 4.5],
                   [6.5, 6.5, 5.5]], np.float32), 1.0, np.array([[5.4, 5.6, 5.5]], np.float32), 5.0),
        (np.array([[0.5, 0.5, 0.5], [11.5, 11.5, 0.5]], np.float32), 12.0, np.array([[6.0, -4.8, 0.5], [6.0, 5.0, 0.5]], np.float32), 12.0),
    ]
    for pts, vs, scan, md in cases:
        gm = E.VoxelHashMap(vs, 30, device=0)
        gm.AddPoints(pts)
        gm.CalVoxelCovAll()
        gm.CalPointCovAll(0.4)
        om = O.VoxelHashMap(vs, 30)
        om.AddPoints(pts)
        om.CalVoxelCovAll()
        om.CalPointCovAll(0.4)
        for method in (E.P2P, E.GICP, E.VGICP, E.AVGICP):
            gc, gt = g.correspondences(scan, gm, I4, method, md)
            oc, ot = O.correspondences(om, scan, I4, method, md)
            assert np.array_equal(gc, oc) and np.array_equal(gt, ot), (method, gc, oc)
            kw = dict(icp_method=method, max_search_dist=md)
            gl = g.linearize(scan, gm, I4, E.RegistrationConfig(**kw))
            ol = o.linearize(scan, om, I4, O.make_config(**kw))
            assert gl["n_corr"] == ol["n_corr"]
            assert np.abs(gl["JTJ"] - ol["JTJ"]).max() <= 1e-5 * max(1e-300, np.abs(ol["JTJ"]).max())
            assert abs(gl["residual_sum"] - ol["residual_sum"]) < 1e-9


def test_comm_path_world_size_one_equals_fused_path(small):
    """elm_registration_set_comm with a single rank runs reduce -> ncclAllReduce -> separate solve kernel; the results
    must equal the fused single-GPU path bit for bit (same sums, same solve code)."""
    scan = synth.scan_m(small["stored"], 2048, small["T_true"])
    cfg = E.RegistrationConfig(icp_method=E.GICP, max_iteration=6, **synth.timing_knobs())
    a = E.Registration(device=0)
    b = E.Registration(device=0)
    b.set_comm(E.Registration.comm_unique_id(), 0, 1)
    Ta, oka, fa, ca = a.RunRegister(scan, small["gm"], small["T0"], cfg)
    Tb, okb, fb, cb = b.RunRegister(scan, small["gm"], small["T0"], cfg)
    assert oka == okb and np.array_equal(Ta, Tb) and fa == fb and np.array_equal(ca, cb)


@pytest.mark.parametrize("method", [E.P2P, E.AVGICP])
def test_peer_exchange_world_size_one_equals_fused_path(small, method):
    """elm_registration_peer_attach with a single rank runs the in-kernel mailbox exchange (write own slot, raise and
    wait for the flag, sum the slots) before the solve; the results must equal the plain single-GPU path bit for bit."""
    scan = synth.scan_m(small["stored"], 3000, small["T_true"])
    cfg = E.RegistrationConfig(icp_method=method, max_iteration=6, **synth.timing_knobs())
    a = E.Registration(device=0)
    b = E.Registration(device=0)
    b.peer_attach([b.peer_export()], 0, 1)
    for _ in range(2):  # twice: the mailbox sequence number keeps counting across calls
        Ta, oka, fa, ca = a.RunRegister(scan, small["gm"], small["T0"], cfg)
        Tb, okb, fb, cb = b.RunRegister(scan, small["gm"], small["T0"], cfg)
        assert oka == okb and np.array_equal(Ta, Tb) and fa == fb and np.array_equal(ca, cb)
    b.peer_detach()
    Tb, okb, fb, cb = b.RunRegister(scan, small["gm"], small["T0"], cfg)
    assert np.array_equal(Ta, Tb)


@pytest.mark.parametrize("method", [E.P2P, E.GICP])
def test_fused_kernel_matches_two_kernel_path(small, method):
    """elm_registration_set_fused(1): one kernel per iteration; sums equal the default path to rounding"""
    scan = synth.scan_m(small["stored"], 3000, small["T_true"])
    cfg = E.RegistrationConfig(icp_method=method, max_iteration=6, **synth.timing_knobs())
    a, b = E.Registration(device=0), E.Registration(device=0)
    b.set_fused(True)
    la, lb = a.linearize(scan, small["gm"], small["T0"], cfg), b.linearize(scan, small["gm"], small["T0"], cfg)
    assert la["n_corr"] == lb["n_corr"] and rel_err(la["JTJ"], lb["JTJ"]) < 1e-12 and rel_err(la["JTr"], lb["JTr"]) < 1e-10
    Ta, oka, fa, ca = a.RunRegister(scan, small["gm"], small["T0"], cfg)
    Tb, okb, fb, cb = b.RunRegister(scan, small["gm"], small["T0"], cfg)
    assert oka == okb and rel_err(Tb, Ta) < 1e-10 and abs(fa - fb) < 1e-10 and rel_err(cb, ca) < 1e-8


def test_run_to_run_bit_reproducible(small):
    scan = synth.scan_m(small["stored"], 5000, small["T_true"])
    cfg = E.RegistrationConfig(icp_method=E.P2P, max_iteration=8, **synth.timing_knobs())
    outs = [small["greg"].RunRegister(scan, small["gm"], small["T0"], cfg) for _ in range(3)]
    for r in outs[1:]:
        assert np.array_equal(r[0], outs[0][0]) and r[2] == outs[0][2]


def test_full_size_properties():
    """BASELINE config 2 size (131072-point Scan-U vs a 2 M-raw-point slab of Map-U; the oracle cannot finish the
    10 M map in seconds): size-independent properties.
      * exact pruning == exhaustive visit (index-level, bit exact)
      * a permuted scan gives the same matches (permuted) and the same sums to rounding
      * a 4096-point sample of the matches equals the oracle's, bit exact"""
    raw = synth.map_u(2_000_000, 58.5)
    gm = E.VoxelHashMap(1.0, 30, device=0)
    gm.AddPoints(raw)
    reg = E.Registration(device=0)
    scan = synth.scan_u(131072, 23.0)
    T = synth.se3([29.0, 29.0, 29.0], np.deg2rad([1.0, -2.0, 30.0]))
    c1, t1 = reg.correspondences(scan, gm, T, E.P2P, 5.0)
    reg.set_exhaustive(True)
    c2, t2 = reg.correspondences(scan, gm, T, E.P2P, 5.0)
    reg.set_exhaustive(False)
    assert np.array_equal(c1, c2) and np.array_equal(t1, t2)
    perm = np.random.default_rng(5).permutation(len(scan))
    c3, t3 = reg.correspondences(scan[perm], gm, T, E.P2P, 5.0)
    assert np.array_equal(c3, c1[perm]) and np.array_equal(t3, t1[perm])
    cfg = E.RegistrationConfig(icp_method=E.P2P, **synth.timing_knobs())
    la, lb = reg.linearize(scan, gm, T, cfg), reg.linearize(scan[perm], gm, T, cfg)
    assert la["n_corr"] == lb["n_corr"] == int(c1.sum())
    assert rel_err(la["JTJ"], lb["JTJ"]) < 1e-12 and rel_err(la["JTr"], lb["JTr"]) < 1e-10
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    idx = np.random.default_rng(6).choice(len(scan), 4096, replace=False)
    oc, ot = O.correspondences(om, scan[idx], T, O.P2P, 5.0)
    assert np.array_equal(oc, c1[idx]) and np.array_equal(ot, t1[idx])


@pytest.mark.parametrize("origin", [0.0, -8.0, 4096.0])
@pytest.mark.parametrize("exhaustive", [False, True])
def test_exact_ties_and_near_ties(origin, exhaustive):
    """Lattice map + queries at cell / face / edge centres: many candidates at EXACTLY equal distance, so the winner is
    decided by the reference's first-in-visit-order rule (strict <, voxel_hash_map.cpp:45); tiny offsets put further
    queries inside and just outside the fp32 pre-filter's error band (wide at origin 4096).  Bit-exact vs the oracle."""
    g = np.arange(16, dtype=np.float64) * 0.5 + 0.25
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(7)
    raw = (lat[rng.permutation(len(lat))] + origin).astype(np.float32)
    gm = E.VoxelHashMap(1.0, 30, device=0)
    gm.AddPoints(raw)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    base = rng.integers(2, 13, size=(600, 3)).astype(np.float64) * 0.5 + 0.25      # a lattice point
    kind = rng.integers(0, 4, size=600)
    off = np.zeros((600, 3))
    off[kind >= 1, 0] = 0.25                                                       # edge midpoint: 2 ties
    off[kind >= 2, 1] = 0.25                                                       # face centre: 4 ties
    off[kind >= 3, 2] = 0.25                                                       # cell centre: 8 ties
    eps = rng.choice([0.0, 1e-7, -1e-7, 1e-6, 3e-5, -2e-4], size=(600, 3))
    scan = (base + off + eps + origin).astype(np.float32)
    reg = E.Registration(device=0)
    reg.set_exhaustive(exhaustive)
    for T in (I4, synth.se3([0.0, 0.0, 0.0], [0.0, 0.0, 0.0]) @ I4):
        gc, gt = reg.correspondences(scan, gm, T, E.P2P, 5.0)
        oc, ot = O.correspondences(om, scan, T, E.P2P, 5.0)
        assert np.array_equal(gc, oc)
        assert np.array_equal(gt, ot), np.flatnonzero((gt != ot).any(axis=(1, 2)))[:10]
    # a pose that is not the identity: the transformed query is no longer an fp32 value
    T = synth.se3([0.125 + origin * 1e-3, -0.25, 0.0625], [0.0, 0.0, np.pi / 2])
    local = (np.linalg.inv(T) @ np.c_[scan.astype(np.float64), np.ones(len(scan))].T).T[:, :3].astype(np.float32)
    gc, gt = reg.correspondences(local, gm, T, E.P2P, 5.0)
    oc, ot = O.correspondences(om, local, T, E.P2P, 5.0)
    assert np.array_equal(gc, oc) and np.array_equal(gt, ot)


def test_registration_on_a_restored_map_is_bit_identical(small, tmp_path):
    """elm_map_save -> elm_map_load -> RunRegister: the restored device map (points, directory, covariances) gives the same
    bits as the map it was saved from, for every method."""
    path = tmp_path / "small.elm"
    small["gm"].Save(path)
    lm = E.VoxelHashMap.Load(path, device=0)
    scan = synth.scan_m(small["stored"], 2500, small["T_true"], seed=5)
    for method in (E.P2P, E.GICP, E.VGICP, E.AVGICP):
        cfg = E.RegistrationConfig(icp_method=method, max_iteration=5, **synth.timing_knobs())
        a = small["greg"].RunRegister(scan, small["gm"], small["T0"], cfg)
        b = small["greg"].RunRegister(scan, lm, small["T0"], cfg)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2] and np.array_equal(a[3], b[3]), method


@pytest.mark.parametrize("n", [1, 255, 5000, 200_000])
def test_scan_preprocess_matches_oracle(n):
    """FilterPointsByDistance + VoxelDownsample on the GPU (pcm_matching.cpp:451-465, voxel_hash_map.hpp:260-283) vs the
    oracle: same survivors, same (input) order, bit-identical coordinates; the relative time stamps travel with them."""
    rng = np.random.default_rng(n)
    xyz = ((rng.random((n, 3), dtype=np.float32) * 2 - 1) * np.float32(60.0)).astype(np.float32)
    xyz[::7] = np.round(xyz[::7] * 2) / 2            # points exactly on voxel faces of the 0.5 m / 1.5 m grids
    xyz[::11, 0] = np.float32(50.0)                  # ... and exactly at the distance limit along x
    xyz[::11, 1:] = 0
    rel = rng.random(n).astype(np.float32)
    reg = E.Registration(device=0)
    for max_dist, vs in [(50.0, 0.0), (0.0, 1.5), (50.0, 1.5), (45.5, 0.5), (0.0, 0.0)]:
        want = O.scan_preprocess(xyz, max_dist, vs)
        got_xyz, got_rel, got_idx = reg.PreprocessScan(xyz, max_dist, vs, aux=rel)
        assert np.array_equal(got_idx, want), (max_dist, vs)
        assert np.array_equal(got_xyz, xyz[want]) and np.array_equal(got_rel, rel[want])
    with pytest.raises(E.ElmError):                  # a voxel key that cannot be packed is reported, not silently dropped
        reg.PreprocessScan(np.array([[1e9, 0, 0]], np.float32), 0.0, 0.5)
