"""Host-side logic of the product, runnable without a GPU: the C-ABI library loads and exports every declared symbol,
the host map builder (sort-based AddPoints + covariance passes) matches the oracle, and compute entry points refuse
loudly when there is no CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import elimaloc_b200 as E
from elimaloc_b200 import _capi, synth
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "elimaloc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(elm_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 25
    lib = C.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/elimaloc_b200.h but not exported"
    assert set(names) == set(_capi.SIGNATURES), "ctypes binding and header disagree"


def test_config_struct_layout_matches_header():
    assert C.sizeof(_capi.RegConfig) == 6 * 4 + 8 * 8
    assert _capi.RegConfig.max_search_dist.offset == 24
    assert C.sizeof(O.RegConfig) == C.sizeof(_capi.RegConfig)


@pytest.mark.parametrize("origin", [0.0, -6.0])
def test_host_map_builder_matches_oracle(origin):
    raw = synth.map_u(60_000, 18.0, origin=origin)
    pm = E.VoxelHashMap(1.0, 30, device=-1)
    pm.AddPoints(raw)
    pm.CalVoxelCovAll()
    pm.CalPointCovAll(0.4)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw)
    om.CalVoxelCovAll()
    om.CalPointCovAll(0.4)
    pe, oe = pm.export(True, True), om.export()
    for k in ("keys", "counts", "pxyz"):  # integer / order-dependent work: bit exact
        assert np.array_equal(pe[k], oe[k]), k
    for k in ("vmean", "vcov", "pmean", "pcov"):
        assert np.abs(pe[k] - oe[k]).max() < 1e-9, k
    assert not pm.Empty() and pm.num_points() == om.num_points() and pm.num_voxels() == om.num_voxels()


def test_add_points_is_incremental_and_order_dependent():
    """AddPoints may be called repeatedly (voxel_hash_map.cpp:268-285): two calls == one call on the concatenation,
    and the spacing filter depends on arrival order."""
    raw = synth.map_u(30_000, 9.0, origin=-1.0)
    a = E.VoxelHashMap(0.8, 12, device=-1)
    a.AddPoints(raw[:10_000])
    a.AddPoints(raw[10_000:])
    b = E.VoxelHashMap(0.8, 12, device=-1)
    b.AddPoints(raw)
    om = O.VoxelHashMap(0.8, 12)
    om.AddPoints(raw)
    ea, eb, eo = a.export(), b.export(), om.export()
    for k in ("keys", "counts", "pxyz"):
        assert np.array_equal(ea[k], eb[k]) and np.array_equal(eb[k], eo[k]), k
    assert eb["counts"].max() <= 12
    c = E.VoxelHashMap(0.8, 12, device=-1)
    c.AddPoints(raw[::-1].copy())
    assert not np.array_equal(np.sort(c.export()["pxyz"], axis=0), np.sort(eb["pxyz"], axis=0))


def test_empty_and_degenerate_maps():
    m = E.VoxelHashMap(1.0, 30, device=-1)
    assert m.Empty() and m.num_points() == 0
    m.AddPoints(np.zeros((0, 3), np.float32))
    assert m.Empty()
    m.AddPoints(np.array([[0.1, 0.2, 0.3]] * 5, np.float32))  # duplicates: only the first survives the spacing rule
    assert m.num_points() == 1 and m.num_voxels() == 1
    m.CalVoxelCovAll()
    e = m.export(voxel_cov=True)
    assert np.array_equal(e["vcov"][0], np.eye(3)) and np.allclose(e["vmean"][0], [0.1, 0.2, 0.3], atol=1e-7)


def test_error_statuses_without_fallback():
    m = E.VoxelHashMap(1.0, 30, device=-1)
    with pytest.raises(E.ElmError) as ei:  # 2^20 voxels per axis is the key range
        m.AddPoints(np.array([[3.0e6, 0.0, 0.0]], np.float32))
    assert ei.value.status == _capi.ELM_ERR_RANGE
    with pytest.raises(E.ElmError) as ei:
        m.export(voxel_cov=True)
    assert ei.value.status == _capi.ELM_ERR_STATE
    with pytest.raises(E.ElmError):
        E.VoxelHashMap(-1.0, 30, device=-1)
    if E.device_count() == 0:  # no CUDA device: the product refuses, it never computes on the CPU
        with pytest.raises(E.ElmError) as ei:
            E.Registration(device=0)
        assert ei.value.status == _capi.ELM_ERR_CUDA
        with pytest.raises(E.ElmError) as ei:
            E.VoxelHashMap(1.0, 30, device=0)
        assert ei.value.status == _capi.ELM_ERR_CUDA


def test_product_does_not_import_the_oracle():
    """only tests/, smoke() and bench.py's CPU legs may touch oracle/"""
    for pkg in (os.path.join(ROOT, "elimaloc_b200"), os.path.join(ROOT, "shim"), os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(pkg):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                    assert "oracle" not in open(os.path.join(dirpath, f), errors="replace").read().lower(), os.path.join(dirpath, f)


@pytest.mark.parametrize("kind", ["dense_negative", "sparse_surface", "single_point", "incremental", "long_line", "two_far_clusters"])
def test_neighbourhood_directory_is_consistent(kind):
    """The 2-choice directory the P2P/GICP search reads: every centre key whose 27 voxels (GetAdjacentVoxels range 2,
    voxel_hash_map.cpp:232-241) hold a point is found, column descriptors equal the canonical arrays, others miss."""
    pm = E.VoxelHashMap(1.0, 30, device=-1)
    if kind == "dense_negative":
        pm.AddPoints(synth.map_u(80_000, 16.0, origin=-7.0))
    elif kind == "sparse_surface":
        pm.AddPoints(synth.map_s(40_000, 60.0))
    elif kind == "single_point":
        pm.AddPoints(np.array([[-0.5, 0.25, 3.5]], np.float32))
    elif kind == "long_line":          # 60 000 voxels in a row: keys differ in one field only (a bad case for a weak hash)
        x = np.arange(-30_000, 30_000, dtype=np.float64) + 0.5
        pm.AddPoints(np.stack([x, np.full_like(x, 0.5), np.full_like(x, -0.5)], 1).astype(np.float32))
    elif kind == "two_far_clusters":   # near the ends of the key range, 2 million voxels apart
        a = synth.map_u(3_000, 6.0, origin=-1_000_000.0)
        b = synth.map_u(3_000, 6.0, origin=999_990.0, seed=5)
        pm.AddPoints(np.vstack([a, b]))
    else:
        raw = synth.map_u(30_000, 12.0, origin=-2.0)
        pm.AddPoints(raw[:10_000])
        pm.AddPoints(raw[10_000:])
    entries, slots, bad = pm.directory_check()
    assert bad == 0
    pm.CalVoxelCovAll()  # adds the VGICP / AVGICP candidate lists: checked as well from now on
    assert pm.directory_check()[2] == 0
    assert entries >= pm.num_voxels() and slots >= entries and slots <= 8 * max(entries, 2)
    if kind == "single_point":
        assert entries == 27


def test_map_file_round_trip(tmp_path):
    """elm_map_save / elm_map_load: the restored map is identical (canonical arrays, covariances, directory) and corrupt
    or foreign files are refused with ELM_ERR_IO."""
    pm = E.VoxelHashMap(1.0, 30, device=-1)
    pm.AddPoints(synth.map_u(40_000, 14.0, origin=-5.0))
    pm.CalVoxelCovAll()
    pm.CalPointCovAll(0.4)
    path = tmp_path / "map.elm"
    pm.Save(path)
    lm = E.VoxelHashMap.Load(path, device=-1)
    a, b = pm.export(True, True), lm.export(True, True)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert lm.directory_check() == pm.directory_check() and lm.directory_check()[2] == 0
    assert not lm.Empty() and lm.num_points() == pm.num_points()
    # AddPoints keeps working on a restored map (incremental build, voxel_hash_map.cpp:268-285)
    extra = synth.map_u(5_000, 14.0, origin=-5.0, seed=99)
    pm.AddPoints(extra)
    lm.AddPoints(extra)
    assert np.array_equal(pm.export()["pxyz"], lm.export()["pxyz"])
    raw = open(path, "rb").read()
    for bad in (raw[: len(raw) // 2], b"not a map" + raw[9:], raw[:12] + bytes([raw[12] ^ 0xFF]) + raw[13:]):
        p2 = tmp_path / "bad.elm"
        p2.write_bytes(bad)
        with pytest.raises(E.ElmError) as ei:
            E.VoxelHashMap.Load(p2, device=-1)
        assert ei.value.status == _capi.ELM_ERR_IO
    with pytest.raises(E.ElmError):
        E.VoxelHashMap.Load(tmp_path / "missing.elm", device=-1)


def test_result_shaping_matches_oracle_and_closed_forms():
    """elm_shape_pcm_covariance vs the oracle restatement of PublishPcmOdom (pcm_matching.cpp:1082-1098) on random SPD
    inputs, plus the branches of NormalizeCovariance (pcm_matching.hpp:247-273): cap at 5, the 1e9 rescue of tiny diagonals,
    the 0.25 m floor of the std, untouched off-diagonal blocks."""
    rng = np.random.default_rng(3)
    for k in range(50):
        A = rng.normal(size=(6, 6))
        cov = A @ A.T * 10.0 ** rng.integers(-14, 2)
        Rz = synth.se3([0, 0, 0], rng.normal(size=3))[:3, :3]
        std = float(rng.choice([0.01, 0.25, 0.4, 3.0]))
        seed36 = rng.normal(size=(6, 6))
        g = E.shape_pcm_covariance(Rz, cov, std, seed36)
        o = O.shape_pcm_covariance(Rz, cov, std, seed36)
        assert np.allclose(g, o, rtol=1e-12, atol=1e-300), k
        assert np.array_equal(g[:3, 3:], seed36[:3, 3:]) and np.array_equal(g[3:, :3], seed36[3:, :3])
    # closed forms: identity local_cov (every method but GICP, registration.cpp:280) -> std^2 I and (std pi/180)^2 I
    g = E.shape_pcm_covariance(np.eye(3), np.eye(6), 0.1)
    assert np.allclose(g[:3, :3], 0.25 ** 2 * np.eye(3)) and np.allclose(g[3:, 3:], (0.25 * np.pi / 180) ** 2 * np.eye(3))
    g = E.shape_pcm_covariance(np.eye(3), np.diag([1.0, 2.0, 100.0, 1e-12, 2e-12, 1.0]), 1.0)
    assert np.allclose(np.diag(g)[:3], [1.0, 2.0, 5.0])                                   # normalised by the minimum, capped at 5
    assert np.allclose(np.diag(g)[3:], np.array([1.0, 2.0, 5.0]) * (np.pi / 180) ** 2)     # x1e9 rescue: (1e-3, 2e-3, 1e9) / 1e-3, capped
    g = E.shape_pcm_covariance(np.eye(3), np.diag([0.0, 1e-30, 1e-20, 1.0, 1.0, 1.0]), 1.0)
    assert np.allclose(np.diag(g)[:3], [0.0, 1e-12, 0.01])                                # still below the threshold: divided by 1e-9


def test_new_entry_points_validate_their_arguments():
    """argument checks of the entry points added for SURVEY 8f and the peer exchange (no GPU needed: they fail before any CUDA call)"""
    L = _capi.lib()
    cov, R = np.eye(6), np.eye(3)
    dp = C.POINTER(C.c_double)
    assert L.elm_shape_pcm_covariance(None, cov.ctypes.data_as(dp), 1.0, cov.ctypes.data_as(dp)) == _capi.ELM_ERR_INVALID
    assert L.elm_shape_pcm_covariance(R.ctypes.data_as(dp), cov.ctypes.data_as(dp), 1.0, None) == _capi.ELM_ERR_INVALID
    assert L.elm_map_save(None, b"/tmp/x") == _capi.ELM_ERR_INVALID
    h = C.c_void_p()
    assert L.elm_map_load(C.byref(h), None, -1) == _capi.ELM_ERR_INVALID
    assert L.elm_map_load(C.byref(h), b"/nonexistent/dir/map.elm", -1) == _capi.ELM_ERR_IO
    assert b"cannot open" in L.elm_last_error()
    n = C.c_size_t(7)
    assert L.elm_scan_preprocess(None, None, None, 0, 0.0, 0.0, None, None, None, C.byref(n)) == _capi.ELM_ERR_INVALID
    assert L.elm_registration_peer_export(None, None) == _capi.ELM_ERR_INVALID
    assert L.elm_registration_peer_attach(None, None, 0, 1) == _capi.ELM_ERR_INVALID
    u = C.c_uint64(0)
    assert L.elm_map_directory_check(None, C.byref(u), C.byref(u), C.byref(u)) == _capi.ELM_ERR_INVALID
    pm = E.VoxelHashMap(1.0, 30, device=-1)
    with pytest.raises(E.ElmError) as ei:   # more points per voxel than a column descriptor can count
        E.VoxelHashMap(1.0, 2000, device=-1).AddPoints(np.zeros((1, 3), np.float32))
    assert ei.value.status == _capi.ELM_ERR_RANGE
    assert pm.directory_check() == (0, 0, 0)   # empty map: empty directory


@pytest.mark.parametrize("kind", ["surface_half_metre", "forty_km_wide", "every_host_thread"])
def test_parallel_builder_equals_sequential_insert_on_awkward_maps(kind):
    """The slab-partitioned parallel AddPoints (x-slabs, stable counting sort, per-voxel sequential spacing filter) must equal
    the oracle's one-point-at-a-time insert bit for bit: sparse surface map with 0.5 m voxels; an extent of 40 000 voxels
    along x (more slabs than buckets: several voxel columns per slab) straddling the origin; both fed in three calls."""
    if kind == "surface_half_metre":
        raw, vs = synth.map_s(120_000, 150.0), 0.5
    elif kind == "every_host_thread":   # enough points for one partition chunk per hardware thread
        raw, vs = synth.map_u(600_000, 38.0, origin=-9.0), 1.0
    else:
        rng = np.random.default_rng(1)
        raw, vs = (rng.random((80_000, 3)) * np.array([40000.0, 30.0, 10.0]) - np.array([20000.0, 15.0, 5.0])).astype(np.float32), 1.0
    pm, om = E.VoxelHashMap(vs, 30, device=-1), O.VoxelHashMap(vs, 30)
    for part in np.array_split(raw, 3):
        pm.AddPoints(part)
        om.AddPoints(part)
    pe, oe = pm.export(), om.export()
    for k in ("keys", "counts", "pxyz"):
        assert np.array_equal(pe[k], oe[k]), k
    assert pm.directory_check()[2] == 0


def test_find_ground_height_matches_oracle_and_numpy():
    """VoxelHashMap::FindGroundHeight (voxel_hash_map.hpp:285-322): product (column walk over the sorted map), oracle (full scan
    like the reference) and a numpy one-liner agree; three or fewer points in range -> not found, output untouched."""
    raw = synth.map_s(60_000, 60.0)
    raw[:, 2] += np.float32(-3.0)                       # ground below zero, walls across it
    raw[:, :2] -= np.float32(25.0)                      # positive and negative x / y
    pm, om = E.VoxelHashMap(1.0, 30, device=-1), O.VoxelHashMap(1.0, 30)
    pm.AddPoints(raw)
    om.AddPoints(raw)
    stored = pm.export()["pxyz"].astype(np.float64)
    rng = np.random.default_rng(0)
    for q in np.vstack([rng.uniform(-30, 40, (40, 2)), [[-25.0, -25.0], [1000.0, 0.0], [34.99, 34.99]]]):
        d2 = (stored[:, 0] - q[0]) ** 2 + (stored[:, 1] - q[1]) ** 2
        zs = np.sort(stored[d2 <= 25.0, 2])
        gf, gz = pm.FindGroundHeight(q)
        of, oz = om.FindGroundHeight(q)
        assert gf == of == (len(zs) > 3)
        if gf:
            want = zs[:5].sum() / min(5, len(zs))
            assert abs(gz - want) < 1e-12 and abs(oz - want) < 1e-12
        else:
            assert gz == 0.0
