"""shim/registration_shim.hpp — the header a maintainer drops next to pcm_matching.cpp — compiled against
a minimal Eigen stand-in (tests/mock_eigen; the image has no Eigen) and driven through the node's call sequence
(pcm_matching.cpp:82-101, 280-282) by tests/shim_check.cpp.  CPU: it builds, the VoxelHashMap facade works on a host-only
map and a missing CUDA device degrades to the reference's soft failure.  GPU: RunRegister through the shim returns the
same bits as the ctypes mirror for every method (row-major ABI <-> column-major Eigen conversions included)."""
import os
import subprocess

import numpy as np
import pytest

import elimaloc_b200 as E
from elimaloc_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_shim_check(out_dir, device):
    exe = os.path.join(out_dir, f"shim_check_{device}".replace("-", "m"))
    lib_dir = os.path.join(ROOT, "elimaloc_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", f"-DELM_SHIM_DEVICE={device}", "-I" + os.path.join(ROOT, "tests", "mock_eigen"),
           "-I" + os.path.join(ROOT, "shim"), "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "shim_check.cpp"), "-o", exe,
           os.path.join(lib_dir, "libelimaloc_b200.so"), "-Wl,-rpath," + lib_dir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def run_shim_check(exe, tmp, raw, scan, T0, method):
    raw.astype(np.float32).tofile(os.path.join(tmp, "map.f32"))
    scan.astype(np.float32).tofile(os.path.join(tmp, "scan.f32"))
    out = os.path.join(tmp, f"out{method}.txt")
    args = [exe, os.path.join(tmp, "map.f32"), str(len(raw)), os.path.join(tmp, "scan.f32"), str(len(scan))]
    args += [repr(float(v)) for v in np.asarray(T0).T.reshape(-1)]  # Eigen storage order: column-major
    args += [str(int(method)), out]
    subprocess.run(args, check=True, capture_output=True, text=True, timeout=300)
    lines = open(out).read().split("\n")
    return dict(empty=int(lines[0]), T=np.array(lines[1].split(), float).reshape(4, 4).T, ok=int(lines[2]), fit=float(lines[3]),
                cov=np.array(lines[4].split(), float).reshape(6, 6).T, n_points=int(lines[5]))


def fixture_data():
    raw = synth.map_u(20_000, 12.0, origin=-3.0)
    T_true = synth.se3([2.0, 3.0, 2.5], [0.01, -0.02, 0.2])
    T0 = T_true @ synth.canonical_offset()
    return raw, T_true, T0


def test_shim_compiles_and_degrades_softly_without_a_gpu(tmp_path):
    raw, T_true, T0 = fixture_data()
    hm = E.VoxelHashMap(1.0, 30, device=-1)
    hm.AddPoints(raw)
    scan = synth.scan_m(hm.Pointcloud(), 500, T_true)
    exe = build_shim_check(str(tmp_path), -1)  # host-only map, no registration handle: every compute call must fail softly
    r = run_shim_check(exe, str(tmp_path), raw, scan, T0, E.GICP)
    assert r["empty"] == 0 and r["n_points"] == hm.num_points()
    assert r["ok"] == 0 and r["fit"] == -7.0                      # is_success = false, fitness_score untouched
    assert np.array_equal(r["T"], T0) and np.array_equal(r["cov"], np.eye(6))   # pose = initial guess, local_cov = I


@pytest.mark.gpu
@pytest.mark.parametrize("method", [E.P2P, E.GICP, E.VGICP, E.AVGICP])
def test_shim_run_register_equals_the_c_abi(tmp_path, method):
    raw, T_true, T0 = fixture_data()
    gm = E.VoxelHashMap(1.0, 30, device=0)
    gm.AddPoints(raw)
    gm.CalVoxelCovAll()
    gm.CalPointCovAll(0.4)
    scan = synth.scan_m(gm.Pointcloud(), 1500, T_true)
    exe = build_shim_check(str(tmp_path), 0)
    r = run_shim_check(exe, str(tmp_path), raw, scan, T0, method)
    cfg = E.RegistrationConfig(icp_method=method, max_iteration=6, max_search_dist=5.0, lm_lambda=0.5, icp_termination_threshold_m=0.0,
                               min_overlap_ratio=0.0, max_fitness_score=1e30)
    T, ok, fit, cov = E.Registration(device=0).RunRegister(scan, gm, T0, cfg)
    assert r["ok"] == int(ok) and r["fit"] == fit
    assert np.array_equal(r["T"], T) and np.array_equal(r["cov"], cov)
