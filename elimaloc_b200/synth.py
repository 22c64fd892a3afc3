"""Synthetic maps, scans and poses of BASELINE.md section 4 (numpy.random.default_rng; seeds: map 1234, scan 5678,
pose 91011).  Shared by the tests and bench.py so both sides of every comparison see identical inputs."""
import numpy as np

SEED_MAP, SEED_SCAN, SEED_POSE = 1234, 5678, 91011


def exp_so3(w):
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w)
    if th < 1e-15:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def se3(t, w):
    T = np.eye(4)
    T[:3, :3] = exp_so3(w)
    T[:3, 3] = t
    return T


def map_u(m_raw, box, seed=SEED_MAP, origin=0.0):
    """Map-U: m_raw float32 points ~ U[origin, origin + box)^3 (about 10 raw points per 1 m voxel at the canonical
    sizes: 100 k -> 21.5 m, 10 M -> 100 m, 50 M -> 171 m)."""
    rng = np.random.default_rng(seed)
    out = np.empty((m_raw, 3), dtype=np.float32)
    step = 1 << 22
    for i in range(0, m_raw, step):
        n = min(step, m_raw - i)
        out[i:i + n] = (rng.random((n, 3), dtype=np.float32) * np.float32(box) + np.float32(origin))
    return out


def map_s(m_raw, box, seed=SEED_MAP, ground_z=0.0, n_walls=24, wall_height=6.0):
    """Map-S: surface-like map (optional): a ground plane z ~ N(ground_z, 0.02^2) plus random
    vertical walls with 2 cm thickness noise, inside [0, box)^2.  Voxel-Gaussian methods (VGICP / AVGICP) need surfaces;
    on the uniform Map-U their voxel means carry almost no geometry."""
    rng = np.random.default_rng(seed)
    n_ground = m_raw // 2
    g = np.empty((n_ground, 3))
    g[:, :2] = rng.random((n_ground, 2)) * box
    g[:, 2] = ground_z + rng.normal(0.0, 0.02, n_ground)
    n_wall = m_raw - n_ground
    per = n_wall // n_walls
    walls = []
    for w in range(n_walls):
        n = per if w < n_walls - 1 else n_wall - per * (n_walls - 1)
        c = rng.random(2) * box
        a = rng.random() * np.pi
        d = np.array([np.cos(a), np.sin(a)])
        length = box * (0.2 + 0.5 * rng.random())
        u = (rng.random(n) - 0.5) * length
        xy = c[None, :] + u[:, None] * d[None, :] + rng.normal(0.0, 0.02, (n, 1)) * np.array([-d[1], d[0]])[None, :]
        z = ground_z + rng.random(n) * wall_height
        walls.append(np.column_stack([xy, z]))
    pts = np.vstack([g] + walls)
    keep = (pts[:, 0] >= 0) & (pts[:, 0] < box) & (pts[:, 1] >= 0) & (pts[:, 1] < box)
    pts = pts[keep]
    return pts[rng.permutation(len(pts))].astype(np.float32)


def scan_u(n, half_width, seed=SEED_SCAN):
    """Scan-U (throughput): n float32 points uniform in a cube of the given half width around the sensor."""
    rng = np.random.default_rng(seed)
    return ((rng.random((n, 3), dtype=np.float32) * 2 - 1) * np.float32(half_width)).astype(np.float32)


def scan_m(stored_xyz, n, T_true, noise=0.02, seed=SEED_SCAN):
    """Scan-M (parity): n points drawn from the stored map points + N(0, noise^2), seen from pose T_true."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, stored_xyz.shape[0], n)
    pts = stored_xyz[idx].astype(np.float64) + rng.normal(0.0, noise, (n, 3))
    Ti = np.linalg.inv(T_true)
    return (pts @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)


def canonical_offset():
    """initial_guess = T_true * Exp([0.30, -0.20, 0.10 m ; 0.5, -0.5, 1.0 deg])"""
    return se3([0.30, -0.20, 0.10], np.deg2rad([0.5, -0.5, 1.0]))


def timing_knobs():
    """Knobs that force every iteration to run."""
    return dict(icp_termination_threshold_m=0.0, min_overlap_ratio=0.0, max_fitness_score=1e30, lm_lambda=0.5,
                max_search_dist=5.0, use_radar_cov=0, debug_print=0)


class ScanWorld:
    """BASELINE config 5 stream: constant-twist arc through the map; IMU in the body frame; raw scans with per-point motion
    distortion (shared by tests/pipeline_harness.py and bench.py --config 5)"""

    def __init__(self, box, n_points, seed=7, radius=8.0, omega=0.25, t0=100.0, height=1.6):
        self.c = np.array([box / 2 - 4.0, box / 2 - 6.0, height])  # sensor `height` above the ground plane of Map-S
        self.r, self.w, self.t0, self.n = radius, omega, t0, n_points
        self.rng = np.random.default_rng(seed)

    def pose(self, t):
        a = self.w * (t - self.t0)
        T = np.eye(4)
        T[:3, :3] = exp_so3([0, 0, a])
        T[:3, 3] = self.c + self.r * np.array([np.sin(a), 1 - np.cos(a), 0.0])
        return T

    def imu(self, t):
        v = self.r * self.w
        gyro = np.array([0.0, 0.0, self.w]) + self.rng.normal(0, 5e-4, 3)
        acc = np.array([0.0, v * self.w, 9.81]) + self.rng.normal(0, 5e-3, 3)
        return gyro, acc

    def scan(self, stored, t_end, span=0.1):
        Te = self.pose(t_end)
        near = stored[np.linalg.norm(stored - Te[:3, 3].astype(np.float32), axis=1) < 14.0]
        idx = self.rng.integers(0, len(near), self.n)
        tt = np.sort(self.rng.random(self.n)) * span
        pts = near[idx].astype(np.float64) + self.rng.normal(0, 0.01, (self.n, 3))
        a = self.w * (t_end - span + tt - self.t0)
        ca, sa = np.cos(a), np.sin(a)
        pos = self.c[None, :] + self.r * np.stack([sa, 1 - ca, np.zeros_like(a)], axis=1)
        d = pts - pos
        local = np.stack([ca * d[:, 0] + sa * d[:, 1], -sa * d[:, 0] + ca * d[:, 1], d[:, 2]], axis=1)  # R(t)^T (p - pos(t))
        return local.astype(np.float32), tt.astype(np.float32)
