"""GPU map build (elimaloc_b200/csrc/map_build.cu): AddPoints (stable sort by voxel + per-voxel replay of the spacing
test in arrival order), CalVoxelCovAll and CalPointCovAll on the device against the product's host builder — which the CPU suite
pins on the reference's own voxel_hash_map.cpp (tests/test_reference_build.py) — BIT for bit: keys, counts, stored points in
canonical order, voxel means / covariances, point means / covariances.  The bar is array_equal, not a tolerance: both builders share
one source for the covariance arithmetic (cov_math.hpp) and neither contracts into FMA."""
import time

import numpy as np
import pytest

import elimaloc_b200 as E
from elimaloc_b200 import synth

pytestmark = pytest.mark.gpu


def build(raw, device, vs=1.0, cap=30, gpu_build=True, point_cov=True, r=0.4):
    m = E.VoxelHashMap(vs, cap, device=device)
    if device >= 0:
        m.set_gpu_build(gpu_build)
    m.AddPoints(raw)
    m.CalVoxelCovAll()
    if point_cov:
        m.CalPointCovAll(r)
    return m


def assert_identical(g, h, point_cov=True):
    eg, eh = g.export(voxel_cov=True, point_cov=point_cov), h.export(voxel_cov=True, point_cov=point_cov)
    assert set(eg) == set(eh)
    for k in eg:
        assert eg[k].shape == eh[k].shape, k
        assert np.array_equal(eg[k], eh[k]), (k, float(np.abs(eg[k].astype(np.float64) - eh[k].astype(np.float64)).max()))
    assert g.directory_check()[2] == 0


def lattice_map():
    g = (np.arange(-6, 6, dtype=np.float32) * np.float32(0.5))
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + np.float32(0.25)
    rng = np.random.default_rng(3)
    lat = lat[rng.permutation(len(lat))]
    return np.vstack([lat, lat + np.float32(0.125), lat, lat[:200] + np.float32(1e-3)]).astype(np.float32)  # exact duplicates, near duplicates


@pytest.mark.parametrize("kind,vs,cap", [("u_negative", 1.0, 30), ("u_negative", 0.5, 8), ("surface", 1.0, 30), ("surface", 2.0, 100),
                                         ("lattice", 1.0, 30), ("lattice", 0.7, 3), ("dense_cap", 1.0, 5)])
def test_gpu_builder_is_bit_identical_to_the_host_builder(kind, vs, cap):
    if kind == "u_negative":
        raw = synth.map_u(300_000, 30.0, origin=-12.0)       # straddles the origin: truncating insert keys (Q1)
    elif kind == "surface":
        raw = synth.map_s(400_000, 40.0) - np.float32(7.0)     # planes: rank-deficient covariances, thousands of raw points per voxel
    elif kind == "lattice":
        raw = lattice_map()
    else:
        raw = synth.map_u(200_000, 6.0, origin=-3.0)           # ~900 raw points per voxel: the cap and the spacing test decide
    g = build(raw, 0, vs, cap, gpu_build=True)
    h = build(raw, -1, vs, cap)
    assert_identical(g, h)
    assert g.num_points() < len(raw)                           # the spacing filter really dropped points


def test_gpu_builder_then_incremental_add_on_the_host():
    """the first AddPoints of a map runs on the GPU, later ones merge on the host: same map as two host calls"""
    raw = synth.map_u(200_000, 20.0, origin=-5.0)
    g = E.VoxelHashMap(1.0, 30, device=0)
    h = E.VoxelHashMap(1.0, 30, device=-1)
    for m in (g, h):
        m.AddPoints(raw[:150_000])
        m.AddPoints(raw[150_000:])
        m.CalVoxelCovAll()
        m.CalPointCovAll(0.4)
    assert_identical(g, h)


def test_gpu_builder_rejects_what_the_host_builder_rejects():
    bad = np.array([[0, 0, 0], [3e6, 0, 0]], np.float32)       # beyond +-2^20 voxels
    with pytest.raises(E.ElmError) as e:
        E.VoxelHashMap(1.0, 30, device=0).AddPoints(bad)
    assert e.value.status == E._capi.ELM_ERR_RANGE
    nan = np.array([[0, 0, 0], [np.nan, 0, 0]], np.float32)
    with pytest.raises(E.ElmError):
        E.VoxelHashMap(1.0, 30, device=0).AddPoints(nan)


def test_registration_on_a_gpu_built_map_equals_a_host_built_map():
    raw = synth.map_u(300_000, 30.0, origin=-12.0)
    g = build(raw, 0, gpu_build=True)
    h = build(raw, 0, gpu_build=False)
    T_true = synth.se3([2.0, 3.0, 2.5], [0.01, -0.02, 0.2])
    scan = synth.scan_m(g.Pointcloud(), 4096, T_true)
    T0 = T_true @ synth.canonical_offset()
    reg = E.Registration(device=0)
    for method in (E.P2P, E.GICP, E.VGICP, E.AVGICP):
        cfg = E.RegistrationConfig(icp_method=method, max_iteration=5, **synth.timing_knobs())
        a, b = reg.RunRegister(scan, g, T0, cfg), reg.RunRegister(scan, h, T0, cfg)
        assert np.array_equal(a[0], b[0]) and a[2] == b[2] and np.array_equal(a[3], b[3])


@pytest.mark.parametrize("m_raw,box", [(10_000_000, 100.0)])
def test_full_size_build_identical_and_timed(m_raw, box, capsys):
    """BASELINE config-2 / 3 size: 10 M raw points, all three passes, GPU vs host builder bit for bit; prints the build times"""
    raw = synth.map_u(m_raw, box)
    t = time.time(); g = build(raw, 0, gpu_build=True); tg = time.time() - t
    gt = g.build_times()
    t = time.time(); h = build(raw, 0, gpu_build=False); th = time.time() - t
    ht = h.build_times()
    assert_identical(g, h)
    with capsys.disabled():
        print(f"\n[map build {m_raw} raw -> {g.num_points()} stored, {g.num_voxels()} voxels] builder proper, ms: "
              f"AddPoints GPU {gt['add_points_ms']:.0f} / host {ht['add_points_ms']:.0f}, CalVoxelCovAll GPU {gt['voxel_cov_ms']:.0f} / host {ht['voxel_cov_ms']:.0f}, "
              f"CalPointCovAll GPU {gt['point_cov_ms']:.0f} / host {ht['point_cov_ms']:.0f}; whole calls incl. derived tables + upload: GPU {tg:.2f} s / host {th:.2f} s")
