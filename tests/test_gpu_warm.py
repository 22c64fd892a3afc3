"""Warm-started P2P / GICP search (icp_search_warm_kernel) against the oracle, through the C ABI.

From the second ICP iteration on the search starts from the previous iteration's match and reads only the octants of the
27 voxels that lie within that distance.  It must return exactly what GetCorrespondencePoints returns at that pose
(voxel_hash_map.cpp:31-88: nearest stored point of the 27 voxels, first in visit order among equals) — whatever the
sequence of poses before it was: tiny steps, steps that move queries into other voxels, jumps that leave the old match
outside the new neighbourhood, sparse regions, negative coordinates (insert keys truncate, Q1), exact ties."""
import numpy as np
import pytest

import elimaloc_b200 as E
from elimaloc_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def make_world(raw, voxel=1.0, cap=30, cov=True):
    gm = E.VoxelHashMap(voxel, cap, device=0)
    gm.AddPoints(raw)
    om = O.VoxelHashMap(voxel, cap)
    om.AddPoints(raw)
    if cov:
        gm.CalPointCovAll(0.4)
        om.CalPointCovAll(0.4)
    return gm, om


def pose_walk(T, steps, rng, trans, rot):
    out = [T]
    for _ in range(steps):
        d = synth.se3(rng.normal(0, trans, 3), rng.normal(0, rot, 3))
        out.append(out[-1] @ d)
    return out


def check_sequence(greg, gm, om, scan, poses, method, max_dist=5.0):
    """every prefix of the pose sequence: the warm-started result at its last pose == the oracle at that pose"""
    for k in range(1, len(poses) + 1):
        gc, gt = greg.correspondences_sequence(scan, gm, poses[:k], method, max_dist)
        oc, ot = O.correspondences(om, scan, poses[k - 1], method, max_dist)
        assert np.array_equal(gc, oc), (k, int((gc != oc).sum()))
        assert np.array_equal(gt, ot), (k, int((gt != ot).any(axis=(1, 2)).sum()))


@pytest.fixture(scope="module")
def dense():
    raw = synth.map_u(120_000, 23.0, origin=-7.0)   # ~10 points per voxel, straddles the origin on every axis
    gm, om = make_world(raw)
    return dict(gm=gm, om=om, stored=gm.Pointcloud(), greg=E.Registration(device=0))


@pytest.mark.parametrize("method", [E.P2P, E.GICP])
@pytest.mark.parametrize("trans,rot", [(0.002, 0.0002), (0.03, 0.003), (0.3, 0.02), (1.2, 0.1)])
def test_warm_search_equals_the_oracle_along_pose_walks(dense, method, trans, rot):
    """millimetre steps (the converged loop), centimetre steps, steps that change most queries' voxel, jumps beyond a voxel"""
    rng = np.random.default_rng(int(trans * 1e4) + method)
    T_true = synth.se3([3.0, 4.0, 2.5], [0.02, -0.01, 0.3])
    scan = synth.scan_m(dense["stored"], 5000, T_true, seed=11)
    poses = pose_walk(T_true @ synth.canonical_offset(), 4, rng, trans, rot)
    check_sequence(dense["greg"], dense["gm"], dense["om"], scan, poses, method)


def test_warm_search_uniform_scan_with_queries_outside_the_map(dense):
    """Scan-U: queries anywhere, many of them in empty neighbourhoods (Q2 default match) or at the map's border"""
    rng = np.random.default_rng(5)
    scan = synth.scan_u(6000, 14.0, seed=3)
    T = synth.se3([4.5, 4.5, 4.5], [0.0, 0.0, 0.1])
    poses = pose_walk(T, 3, rng, 0.05, 0.004)
    check_sequence(dense["greg"], dense["gm"], dense["om"], scan, poses, E.P2P)


def test_warm_search_sparse_map_and_large_bounds():
    """1.5 points per voxel: matches are often a voxel away, bounds reach across the whole neighbourhood"""
    raw = synth.map_u(6000, 16.0, origin=-8.0)
    gm, om = make_world(raw, cov=False)
    rng = np.random.default_rng(9)
    scan = synth.scan_u(4000, 7.0, seed=4)
    poses = pose_walk(synth.se3([0.3, -0.2, 0.1], [0.0, 0.0, 0.0]), 4, rng, 0.08, 0.01)
    check_sequence(E.Registration(device=0), gm, om, scan, poses, E.P2P)


@pytest.mark.parametrize("voxel,cap", [(0.5, 30), (2.0, 60), (1.0, 300), (0.7, 8)])
def test_warm_search_other_voxel_sizes_and_caps(voxel, cap):
    """non-power-of-two voxel sizes (exact division path), a cap above 255 (no octant words: whole voxels)"""
    raw = synth.map_u(60_000, 12.0, origin=-5.0)
    gm, om = make_world(raw, voxel, cap, cov=False)
    rng = np.random.default_rng(cap)
    scan = synth.scan_m(gm.Pointcloud(), 3000, synth.se3([1.0, 0.5, 0.2], [0.0, 0.0, 0.2]), seed=8)
    poses = pose_walk(synth.se3([1.0, 0.5, 0.2], [0.0, 0.0, 0.2]) @ synth.canonical_offset(), 3, rng, 0.03, 0.003)
    check_sequence(E.Registration(device=0), gm, om, scan, poses, E.P2P)


def test_warm_search_exact_ties_on_a_lattice():
    """map points on a 0.25 m lattice, queries on lattice points, cell centres, face centres: exact distance ties between
    points of different octants and different voxels; the first in the reference's visit order must win, warm or cold"""
    g = np.arange(-3.0, 3.0, 0.25, dtype=np.float32)
    raw = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(2)
    raw = raw[rng.permutation(len(raw))]   # insertion order decides the ties
    gm, om = make_world(raw, 1.0, 80, cov=False)
    q = np.stack(np.meshgrid(np.arange(-2.0, 2.0, 0.125), np.arange(-2.0, 2.0, 0.125), np.arange(-1.0, 1.0, 0.125), indexing="ij"), -1)
    scan = q.reshape(-1, 3).astype(np.float32)
    I = np.eye(4)
    shifts = [np.array([0.125, 0.0, 0.0]), np.array([0.0, -0.125, 0.125]), np.array([0.25, 0.25, -0.125]), np.array([1.0, 0.0, 0.0])]
    poses = [I]
    for s in shifts:
        T = poses[-1].copy()
        T[:3, 3] += s
        poses.append(T)
    check_sequence(E.Registration(device=0), gm, om, scan, poses, E.P2P)


@pytest.mark.parametrize("method", [E.P2P, E.GICP])
def test_registration_with_and_without_warm_start_agrees_to_rounding(dense, method):
    """the warm start changes which bytes are read, not a single correspondence; the two paths add the per-point terms in a
    different order (the warm iteration sums inside its two kernels), so the outputs agree to fp64 rounding, not bit for bit"""
    T_true = synth.se3([3.0, 4.0, 2.5], [0.02, -0.01, 0.3])
    scan = synth.scan_m(dense["stored"], 6000, T_true, seed=21)
    T0 = T_true @ synth.canonical_offset()
    cfg = E.RegistrationConfig(icp_method=method, max_iteration=12, **synth.timing_knobs())
    reg = E.Registration(device=0)
    out = []
    for warm in (True, False):
        reg.set_warm_start(warm)
        out.append(reg.RunRegister(scan, dense["gm"], T0, cfg))
    assert np.abs(out[0][0] - out[1][0]).max() < 1e-11 and out[0][1] == out[1][1] and abs(out[0][2] - out[1][2]) < 1e-12
    assert np.abs(out[0][3] - out[1][3]).max() <= 1e-9 * np.abs(out[1][3]).max()
    o = O.Registration().RunRegister(scan, dense["om"], T0, O.make_config(icp_method=method, max_iteration=12, **synth.timing_knobs()))
    assert np.abs(out[0][0] - o["pose"]).max() / np.abs(o["pose"]).max() < 1e-4


def test_warm_state_does_not_leak_between_scans(dense):
    """a new RunRegister call starts cold: the previous scan's matches must not be used for another scan of the same size"""
    T_true = synth.se3([3.0, 4.0, 2.5], [0.02, -0.01, 0.3])
    T0 = T_true @ synth.canonical_offset()
    cfg = E.RegistrationConfig(icp_method=E.P2P, max_iteration=6, **synth.timing_knobs())
    reg = E.Registration(device=0)
    a = synth.scan_m(dense["stored"], 4096, T_true, seed=31)
    b = synth.scan_m(dense["stored"], 4096, T_true, seed=32)
    reg.RunRegister(a, dense["gm"], T0, cfg)
    got = reg.RunRegister(b, dense["gm"], T0, cfg)
    fresh = E.Registration(device=0).RunRegister(b, dense["gm"], T0, cfg)
    assert np.array_equal(got[0], fresh[0]) and got[2] == fresh[2]   # (same path, same order: bit for bit)


@pytest.mark.parametrize("method", [E.P2P, E.GICP])
def test_warm_search_across_voxel_borders(method):
    """Queries that move into another voxel keep their candidate list where both neighbourhoods lie among keys >= 1 (stored keys are
    floor keys there, so "within one voxel size" implies "inside the 27 voxels") and must refresh anywhere else (insert keys truncate
    toward zero, Q1: the neighbourhood of a negative key is shifted).  Maps entirely positive, entirely negative, and straddling the
    origin; many small steps so that a good share of the queries crosses a voxel face with a valid list; every prefix against the
    oracle, bit for bit — and the warm searches that had to refresh after the first one are far fewer on the positive map."""
    late = {}
    for origin in (3.0, -20.0, -1.5):
        raw = synth.map_u(90_000, 20.0, origin=origin)
        gm, om = make_world(raw)
        rng = np.random.default_rng(11)
        c = origin + 10.0
        T = synth.se3([c, c, c], [0.01, -0.02, 0.2])
        scan = synth.scan_m(gm.Pointcloud(), 20000, T, noise=0.03, seed=13)
        poses = pose_walk(T @ synth.canonical_offset(), 6, rng, 0.012, 0.0004)
        reg = E.Registration(device=0)
        check_sequence(reg, gm, om, scan, poses, method)
        reg.set_stats(True)
        reg.correspondences_sequence(scan, gm, poses, method, 5.0)
        s = reg.stats_raw()
        reg.set_stats(False)
        assert s[21] == 6 * len(scan)
        late[origin] = s[20] - len(scan)        # (the first warm search refreshes every query)
    assert late[-20.0] > 100, late              # voxel changes force a refresh among negative keys ...
    assert late[3.0] < 0.3 * late[-20.0], late  # ... and no longer among positive ones
