"""The drop-in claim, demonstrated: the reference's own ROS node pcm_matching.cpp — UNMODIFIED, compiled from /root/reference —
built with shim/registration_shim.hpp standing where its registration.hpp / voxel_hash_map.hpp stood and linked against
elimaloc_b200/libelimaloc_b200.so (ROS / tf / PCL / Eigen are stand-ins, oracle/ref_build/).

CPU: it compiles and links; with the shim's host-only device (-1) the map facade the node uses at start-up (Init, AddPoints,
     CalVoxelCovAll, Pointcloud, Covariances, FindGroundHeight, VoxelDownsample) gives what the reference's own classes give,
     and a lidar message ends in the reference's soft failure ("ICP FAIL", nothing published) instead of a crash.
GPU: the same node on the CUDA library publishes the same poses as the node on the reference's CPU classes."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import elimaloc_b200 as E  # noqa: E402
from elimaloc_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import reference_build as R  # noqa: E402
import test_reference_build_node as N  # noqa: E402

pytestmark = pytest.mark.skipif(not (R.node_on_cuda_available() and R.node_available()),
                                reason="neither /root/reference nor the prebuilt oracle/_ref node libraries are here")


def sorted_rows(a, decimals=None):
    a = np.asarray(a, dtype=np.float64)
    key = np.round(a, decimals) if decimals is not None else a
    return a[np.lexsort(key.T[::-1])]


@pytest.fixture(scope="module")
def raw_map():
    return synth.map_s(60_000, 24.0)


def test_the_unmodified_node_builds_against_the_shim():
    so = R.build_node_on_cuda()
    assert os.path.isfile(so) and os.path.isfile(so.replace("libref_node_on_cuda.so", "libref_node_on_shim_host.so"))


@pytest.mark.parametrize("method", [O.GICP, O.VGICP])
def test_map_facade_under_the_real_node(raw_map, method):
    """Init() of the node on the shim (host-only map) vs Init() of the node on the reference's classes: same stored points,
    same covariance markers (VGICP: one per voxel with more than 2 points, scale from the eigenvalues of the voxel covariance)"""
    ref = R.PcmMatchingNode(raw_map, icp_method=method)
    ref_cloud, ref_markers = ref.init_publications()
    shim = R.PcmMatchingNode(raw_map, on_cuda="host", icp_method=method)
    shim_cloud, shim_markers = shim.init_publications()
    assert shim.map_points() == ref.map_points() == len(ref_cloud) == len(shim_cloud) > 10_000
    assert np.array_equal(sorted_rows(ref_cloud), sorted_rows(shim_cloud))           # hash order vs canonical order: same set
    if method == O.VGICP:
        assert len(ref_markers) == len(shim_markers) > 100
        a, b = sorted_rows(ref_markers, 6), sorted_rows(shim_markers, 6)
        assert np.abs(a - b).max() < 1e-6
    else:
        assert len(ref_markers) == len(shim_markers) == 0


def test_initial_pose_path_reaches_the_registration_and_fails_softly_without_a_device(raw_map):
    """CallbackInitialPose on the host-only shim: FindGroundHeight answers from the shim's map, VoxelDownsample runs, RunRegister
    has no device and degrades to the reference's soft failure — nothing is published, nothing throws"""
    shim = R.PcmMatchingNode(raw_map, on_cuda="host", icp_method=O.GICP)
    ref = R.PcmMatchingNode(raw_map, icp_method=O.GICP)
    N.feed(shim, (12.0, 12.0, 1.6))
    N.feed(ref, (12.0, 12.0, 1.6))
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw_map)
    T = synth.se3([12.0, 12.0, 1.6], [0.0, 0.0, 0.3])
    scan = synth.scan_m(om.export()["pxyz"], 2000, T, noise=0.01, seed=4)
    rel = np.linspace(0.0, 0.1, len(scan)).astype(np.float32)
    assert shim.cloud(N.T0 + 0.05, scan, rel) is None                                # "ICP FAIL": no pose published
    assert shim.initial_pose(12.0, 12.0, 0.3) is None                                # "ICP failed": no init pose published
    assert shim.initial_pose(500.0, 500.0, 0.0) is None                              # no ground there: returns before the ICP
    assert ref.cloud(N.T0 + 0.05, scan, rel) is not None                             # the same inputs localise on the reference
    got = ref.initial_pose(12.0, 12.0, 0.3)
    assert got is not None and np.linalg.norm(got["pos"][:2] - [12.0, 12.0]) < 0.5


@pytest.mark.skipif(E.device_count() > 0, reason="a CUDA device is present: the soft-failure path is not reachable")
def test_node_on_the_cuda_library_without_a_gpu_fails_softly(raw_map):
    node = R.PcmMatchingNode(raw_map[:5000], on_cuda=True, icp_method=O.P2P)
    assert node.map_points() == 0                                                    # elm_map_create had no device to go to
    N.feed(node, (12.0, 12.0, 1.6))
    scan = synth.scan_u(500, 5.0, seed=1)
    assert node.cloud(N.T0 + 0.05, scan, np.linspace(0, 0.1, 500).astype(np.float32)) is None


@pytest.mark.gpu
@pytest.mark.parametrize("method", [O.P2P, O.GICP, O.VGICP, O.AVGICP])
def test_unmodified_node_on_cuda_publishes_what_the_reference_node_publishes(raw_map, method):
    """one lidar message (and one rviz initial pose) through the unmodified node twice: on the reference's CPU classes and on the
    CUDA drop-in.  Same published pose within the north-star tolerance (1e-4 relative), same covariance (1e-5), same decision."""
    kw = dict(lidar_xyz=N.LIDAR_XYZ, lidar_rpy_deg=N.LIDAR_RPY_DEG, icp_method=method, input_max_dist=60.0, max_fitness_score=2.0)
    ref = R.PcmMatchingNode(raw_map, **kw)
    gpu = R.PcmMatchingNode(raw_map, on_cuda=True, **kw)
    assert gpu.map_points() == ref.map_points()
    centre = (12.0, 12.0, 1.6)
    N.feed(ref, centre)
    N.feed(gpu, centre)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw_map)
    stamp, n = N.T0 + 0.05, 4000
    rng = np.random.default_rng(7)
    rel = np.sort(rng.random(n).astype(np.float32) * np.float32(0.1))
    p_end, rpy_end = N.ego_pose(stamp + float(rel[-1]), centre)
    T = np.eye(4)
    T[:3, :3] = N.H.rpy_to_R(*rpy_end)
    T[:3, 3] = p_end
    tf = np.eye(4)
    tf[:3, :3] = N.H.rpy_to_R(*np.deg2rad(N.LIDAR_RPY_DEG))
    tf[:3, 3] = N.LIDAR_XYZ
    scan = synth.scan_m(om.export()["pxyz"], n, T @ tf, noise=0.02, seed=11)
    a, b = ref.cloud(stamp, scan, rel), gpu.cloud(stamp, scan, rel)
    assert (a is None) == (b is None)
    if a is not None:
        scale = np.abs(a["pos"]).max()
        assert np.abs(a["pos"] - b["pos"]).max() <= 1e-4 * scale
        assert min(np.abs(a["quat_wxyz"] - b["quat_wxyz"]).max(), np.abs(a["quat_wxyz"] + b["quat_wxyz"]).max()) <= 1e-4
        assert np.abs(a["cov"] - b["cov"]).max() <= 1e-4 * np.abs(a["cov"]).max()
        assert a["stamp"] == b["stamp"] and len(a["registered_world"]) == len(b["registered_world"])
    ia, ib = ref.initial_pose(p_end[0], p_end[1], rpy_end[2]), gpu.initial_pose(p_end[0], p_end[1], rpy_end[2])
    assert (ia is None) == (ib is None)
    if ia is not None:
        assert np.abs(ia["pos"] - ib["pos"]).max() <= 1e-4 * np.abs(ia["pos"]).max()


def test_shim_host_helpers(tmp_path):
    """VoxelDownsample and TransformPoints of the shim (host code the node calls around RunRegister): the first point of every
    floor-keyed voxel (the oracle's survivor set) emitted in the REFERENCE's order — the survivors' indices equal, position by
    position, what the reference's own VoxelDownsample returns (its hash-table iteration order: same container, hash values, reserve
    and insertion sequence on the same standard library); the transform moves `pose` and leaves `local` alone"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "shim_helpers_check")
    lib_dir = os.path.join(root, "elimaloc_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(root, "oracle", "ref_build", "stubs"), "-I" + os.path.join(root, "shim"),
                    "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "shim_helpers_check.cpp"), "-o", exe,
                    os.path.join(lib_dir, "libelimaloc_b200.so"), "-Wl,-rpath," + lib_dir], check=True, capture_output=True, text=True)
    rng = np.random.default_rng(12)
    xyz = ((rng.random((20_000, 3)) - 0.5) * 40).astype(np.float32)
    xyz.tofile(str(tmp_path / "pts.f32"))
    for voxel in (1.5, 0.5):
        out = subprocess.run([exe, str(tmp_path / "pts.f32"), str(len(xyz)), repr(voxel)], check=True, capture_output=True, text=True).stdout.split("\n")
        k = int(out[0])
        idx = np.array([int(v) for v in out[1:1 + k]])
        assert np.array_equal(np.sort(idx), O.scan_preprocess(xyz, 0.0, voxel))
        want = R.voxel_downsample(xyz, voxel)
        assert np.array_equal(idx, want)
        for line, i in zip(out[1 + k:1 + k + 5], want[:5]):
            v = [float(t) for t in line.split()]
            p = xyz[i].astype(np.float64)
            assert np.allclose(v[:3], [-p[1] + 1.5, p[0] - 2.5, p[2] + 0.25], rtol=0, atol=1e-12)   # pose moved
            assert v[3] == v[0] and v[4] == p[0] and v[5] == p[0] and int(v[6]) == i              # in-place variant; local, input untouched
