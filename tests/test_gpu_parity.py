"""GPU-vs-oracle parity of the registration hot path, called through the C ABI (via the ctypes mirror).

Tolerances (BASELINE.json north_star): 1e-5 relative on JtJ / Jtr entries, 1e-4 relative on the final SE(3) pose;
correspondences (integer/index work) bit-exact."""
import numpy as np
import pytest

import elimaloc_b200 as E
from elimaloc_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

METHODS = [E.P2P, E.GICP, E.VGICP, E.AVGICP]
NAMES = {0: "P2P", 1: "GICP", 2: "VGICP", 3: "AVGICP"}


def both_cfg(**kw):
    return E.RegistrationConfig(**kw), O.make_config(**kw)


@pytest.fixture(scope="module")
def world():
    """config 1 sizes: 100 k raw map points in a 21.5 m box straddling the origin (exercises Q1), 4096-point Scan-M."""
    raw = synth.map_u(100_000, 21.5, origin=-6.0)
    gm = E.VoxelHashMap(1.0, 30, device=0)
    gm.AddPoints(raw)
    gm.CalVoxelCovAll()
    gm.CalPointCovAll(0.4)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    om.CalVoxelCovAll()
    om.CalPointCovAll(0.4)
    stored = gm.Pointcloud()
    T_true = synth.se3([4.0, 5.0, 3.5], [0.02, -0.01, 0.3])
    scan = synth.scan_m(stored, 4096, T_true)
    T0 = T_true @ synth.canonical_offset()
    return dict(gm=gm, om=om, scan=scan, T_true=T_true, T0=T0, greg=E.Registration(device=0), oreg=O.Registration())


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_map_build_matches_oracle(world):
    ge, oe = world["gm"].export(True, True), world["om"].export()
    for k in ("keys", "counts", "pxyz"):
        assert np.array_equal(ge[k], oe[k]), k
    for k in ("vmean", "vcov", "pmean", "pcov"):
        assert np.abs(ge[k] - oe[k]).max() < 1e-9, k


@pytest.mark.parametrize("exhaustive", [False, True])
@pytest.mark.parametrize("method", METHODS)
def test_correspondences_bit_exact(world, method, exhaustive):
    """index-level parity of the search, both with the exact pruning and visiting all 27 voxels like the reference"""
    world["greg"].set_exhaustive(exhaustive)
    try:
        for T in (world["T0"], world["T_true"]):
            gc, gt = world["greg"].correspondences(world["scan"], world["gm"], T, method, 5.0)
            oc, ot = O.correspondences(world["om"], world["scan"], T, method, 5.0)
            assert np.array_equal(gc, oc), NAMES[method]
            assert np.array_equal(gt, ot), NAMES[method]
    finally:
        world["greg"].set_exhaustive(False)


@pytest.mark.parametrize("method", [E.P2P, E.GICP, E.VGICP])
def test_binning_does_not_change_results(world, method):
    """the search walks the scan in spatially binned order; match[] and every sum must be what the caller's order gives"""
    gcfg, _ = both_cfg(icp_method=method, **synth.timing_knobs())
    out = []
    for binning in (True, False):
        world["greg"].set_binning(binning)
        c, t = world["greg"].correspondences(world["scan"], world["gm"], world["T0"], method, 5.0)
        lin = world["greg"].linearize(world["scan"], world["gm"], world["T0"], gcfg)
        out.append((c, t, lin))
    world["greg"].set_binning(False)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2]["JTJ"], out[1][2]["JTJ"]) and np.array_equal(out[0][2]["JTr"], out[1][2]["JTr"])


def test_correspondences_random_scan_bit_exact(world):
    """uniformly random scan (most queries far from any map point, many in empty space): P2P search parity"""
    scan = synth.scan_u(8192, 14.0, seed=77)
    T = synth.se3([4.5, 4.5, 4.5], [0.1, 0.2, -0.4])
    for md in (1.0, 5.0):
        oc, ot = O.correspondences(world["om"], scan, T, E.P2P, md)
        for exhaustive in (False, True):
            world["greg"].set_exhaustive(exhaustive)
            gc, gt = world["greg"].correspondences(scan, world["gm"], T, E.P2P, md)
            world["greg"].set_exhaustive(False)
            assert np.array_equal(gc, oc) and np.array_equal(gt, ot)


@pytest.mark.parametrize("method", METHODS)
def test_single_iteration_linearization(world, method):
    gcfg, ocfg = both_cfg(icp_method=method, **synth.timing_knobs())
    g = world["greg"].linearize(world["scan"], world["gm"], world["T0"], gcfg)
    o = world["oreg"].linearize(world["scan"], world["om"], world["T0"], ocfg)
    assert g["n_corr"] == o["n_corr"]
    assert rel_err(g["JTJ"], o["JTJ"]) < 1e-5
    assert rel_err(g["JTr"], o["JTr"]) < 1e-5
    assert abs(g["residual_sum"] - o["residual_sum"]) <= 1e-9 * max(1.0, abs(o["residual_sum"]))


@pytest.mark.parametrize("method", METHODS)
def test_full_registration_config1(world, method):
    """BASELINE config 1 (all four methods): 10 forced iterations, final pose within 1e-4 relative of the oracle."""
    gcfg, ocfg = both_cfg(icp_method=method, max_iteration=10, **synth.timing_knobs())
    T, ok, fit, cov = world["greg"].RunRegister(world["scan"], world["gm"], world["T0"], gcfg)
    o = world["oreg"].RunRegister(world["scan"], world["om"], world["T0"], ocfg)
    assert ok == o["is_success"]
    assert rel_err(T, o["pose"]) < 1e-4
    assert abs(fit - o["fitness_score"]) <= 1e-6 * max(1.0, abs(o["fitness_score"]))
    assert rel_err(cov, o["local_cov"]) < 1e-5
    if method in (E.P2P, E.GICP):  # converges onto the true pose
        err = np.linalg.inv(world["T_true"]) @ T
        assert np.linalg.norm(err[:3, 3]) < 0.02


@pytest.mark.parametrize("method", METHODS)
def test_reference_golden_config1(world, method):
    """The CUDA path against tests/golden/config1_*.npz — outputs of the REFERENCE's own registration.cpp / voxel_hash_map.cpp
    (oracle/_ref, generated in the build container by tests/golden/make_golden.py; same seeded inputs as this module's world):
    correspondences of the first 256 scan points bit-equal, first-iteration pair count equal, pose 1e-4, JTJ 1e-5."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"config1_{NAMES[method].lower()}.npz"))
    gcfg, _ = both_cfg(icp_method=method, max_iteration=10, **synth.timing_knobs())
    T, ok, fit, cov = world["greg"].RunRegister(world["scan"], world["gm"], world["T0"], gcfg)
    assert ok == bool(g["is_success"])
    assert rel_err(T, g["pose"]) < 1e-4
    assert abs(fit - float(g["fitness"])) <= 1e-6 * max(1.0, abs(float(g["fitness"])))
    assert rel_err(cov, g["local_cov"]) < 1e-5
    gc, gt = world["greg"].correspondences(world["scan"][:256], world["gm"], world["T0"], method, 5.0)
    assert np.array_equal(gc, g["corr_count"]) and np.array_equal(gt, g["corr_target"])
    lin = world["greg"].linearize(world["scan"], world["gm"], world["T0"], gcfg)
    assert lin["n_corr"] == int(g["ncorr"][0])
    assert rel_err(lin["JTJ"], g["JTJ"][0]) < 1e-5 and rel_err(lin["JTr"], g["JTr"][0]) < 1e-5


@pytest.mark.parametrize("method", METHODS)
def test_default_ini_config(world, method):
    """The reference's own knobs (localization.ini): early termination + overlap + fitness gates."""
    gcfg, ocfg = both_cfg(icp_method=method)
    T, ok, fit, cov = world["greg"].RunRegister(world["scan"], world["gm"], world["T0"], gcfg, fitness_score=-1.0)
    o = world["oreg"].RunRegister(world["scan"], world["om"], world["T0"], ocfg, fitness_in=-1.0)
    assert ok == o["is_success"]
    assert rel_err(T, o["pose"]) < 1e-4
    assert abs(fit - o["fitness_score"]) <= 1e-6 * max(1.0, abs(o["fitness_score"]))


@pytest.mark.parametrize("method", METHODS)
def test_gpu_against_the_reference_sources_directly(world, method):
    """no oracle in between: the CUDA path against oracle/_ref/libref.so — the reference's own registration.cpp /
    voxel_hash_map.cpp compiled against stand-in third-party headers (built in the build container; the prebuilt file travels
    to the GPU box).  Correspondences bit-equal, RunRegister with the reference's .ini knobs within the north-star tolerances."""
    from oracle import reference_build as RB
    if not RB.available():
        pytest.skip("oracle/_ref/libref.so is not here")
    if "rm" not in world:
        rm = RB.VoxelHashMap(1.0, 30)
        rm.AddPoints(synth.map_u(100_000, 21.5, origin=-6.0))
        rm.CalVoxelCovAll()
        rm.CalPointCovAll(0.4)
        world["rm"] = rm
    rm = world["rm"]
    for T in (world["T0"], world["T_true"]):
        gc, gt = world["greg"].correspondences(world["scan"], world["gm"], T, method, 5.0)
        rc, rt = RB.correspondences(rm, world["scan"], T, method, 5.0)
        assert np.array_equal(gc, rc) and np.array_equal(gt, rt), NAMES[method]
    gcfg, rcfg = both_cfg(icp_method=method)
    T, ok, fit, cov = world["greg"].RunRegister(world["scan"], world["gm"], world["T0"], gcfg, fitness_score=-1.0)
    r = RB.Registration().RunRegister(world["scan"], rm, world["T0"], rcfg, fitness_in=-1.0)
    assert ok == r["is_success"]
    assert rel_err(T, r["pose"]) < 1e-4
    assert abs(fit - r["fitness_score"]) <= 1e-6 * max(1.0, abs(r["fitness_score"]))
    assert rel_err(cov, r["local_cov"]) < 1e-5
