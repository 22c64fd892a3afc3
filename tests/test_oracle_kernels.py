"""The third-party arithmetic kernels the oracle restates (oracle/smallmat.hpp, shared with the stand-in Eigen of oracle/_ref) —
Matrix{3,4,6}d::inverse, LDLT solve, SelfAdjointEigenSolver<Matrix3d>, the JacobiSVD-based plane regularisation, AngleAxisd —
each on its own against LAPACK / closed forms through numpy, on random well- and badly-conditioned inputs.  Runs without a GPU.
These are the only parts of the oracle that the pin on the reference's own sources (tests/test_reference_build*.py) does not cover."""
import ctypes as C

import numpy as np
import pytest

from elimaloc_b200 import synth
from oracle import oracle as O

DP = C.POINTER(C.c_double)


def d(a):
    return a.ctypes.data_as(DP)


@pytest.fixture(scope="module")
def L():
    lib = O.lib()
    lib.orc_k_inverse.argtypes = [C.c_int, DP, DP]
    lib.orc_k_ldlt_solve.argtypes = [DP, DP, DP]
    lib.orc_k_sym_eig3.argtypes = [DP, DP, DP]
    lib.orc_k_plane_regularize.argtypes = [DP, DP, DP]
    lib.orc_k_angle_axis_to_rot.argtypes = [C.c_double, DP, DP]
    lib.orc_k_rot_angle.restype = C.c_double
    lib.orc_k_rot_angle.argtypes = [DP]
    return lib


@pytest.mark.parametrize("n", [3, 4, 6])
def test_inverse(L, n):
    rng = np.random.default_rng(n)
    for trial in range(200):
        A = rng.normal(size=(n, n)) * 10.0 ** rng.integers(-3, 4)
        if n == 4 and trial % 2:  # what the path inverts: rigid transforms
            A = synth.se3(rng.normal(0, 50, 3), rng.normal(0, 1, 3))
        if n == 6 and trial % 2:  # ... and regularised normal equations
            J = rng.normal(size=(40, 6)) * np.array([1, 1, 1, 30, 30, 30])
            A = J.T @ J
            A = A + 0.5 * np.diag(np.diag(A))
        out = np.zeros((n, n))
        L.orc_k_inverse(n, d(np.ascontiguousarray(A)), d(out))
        want = np.linalg.inv(A)
        assert np.abs(out - want).max() <= 1e-10 * np.linalg.cond(A) * np.abs(want).max() / 1e3 + 1e-13 * np.abs(want).max()


def test_ldlt_solve(L):
    rng = np.random.default_rng(0)
    for trial in range(300):
        J = rng.normal(size=(30, 6)) * np.array([1, 1, 1, 20, 20, 20]) * 10.0 ** rng.integers(-2, 3)
        A = J.T @ J
        A = A + [0.0, 0.1, 0.5][trial % 3] * np.diag(np.diag(A))
        b = rng.normal(size=6) * 100
        x = np.zeros(6)
        L.orc_k_ldlt_solve(d(np.ascontiguousarray(A)), d(b), d(x))
        want = np.linalg.solve(A, b)
        assert np.abs(x - want).max() <= 1e-9 * np.abs(want).max()
    # zero matrix: every pivot is below the threshold, Eigen's solve returns zero (the empty-scan path of RunRegister)
    x = np.ones(6)
    L.orc_k_ldlt_solve(d(np.zeros((6, 6))), d(np.ones(6)), d(x))
    assert np.array_equal(x, np.zeros(6))
    # a zero row / column (a direction without information): that component is zero, the rest is the reduced solve
    A = np.diag([4.0, 9.0, 0.0, 1.0, 2.0, 3.0])
    A[0, 1] = A[1, 0] = 1.0
    b = np.arange(1.0, 7.0)
    L.orc_k_ldlt_solve(d(A), d(b), d(x))
    keep = [0, 1, 3, 4, 5]
    assert x[2] == 0.0 and np.allclose(x[keep], np.linalg.solve(A[np.ix_(keep, keep)], b[keep]), rtol=1e-13)


def test_symmetric_eigen_solver(L):
    rng = np.random.default_rng(1)
    for trial in range(300):
        P = rng.normal(size=(8, 3)) * np.array([1.0, 10.0 ** rng.integers(-3, 1), 10.0 ** rng.integers(-3, 1)])
        A = np.cov((P @ synth.exp_so3(rng.normal(0, 1, 3))).T)
        w, V = np.zeros(3), np.zeros((3, 3))
        L.orc_k_sym_eig3(d(np.ascontiguousarray(A)), d(w), d(V))
        ww, VV = np.linalg.eigh(A)
        assert np.all(np.diff(w) >= 0) and np.abs(w - ww).max() <= 1e-12 * max(ww.max(), 1e-300)
        assert np.abs(V.T @ V - np.eye(3)).max() < 1e-12
        assert np.abs(A @ V - V * w).max() <= 1e-12 * ww.max()          # columns are eigenvectors
    w, V = np.zeros(3), np.ones((3, 3))
    L.orc_k_sym_eig3(d(np.diag([3.0, 1.0, 2.0])), d(w), d(V))            # diagonal input: a permutation, no rotation
    assert np.array_equal(w, [1.0, 2.0, 3.0]) and np.array_equal(np.abs(V), np.eye(3)[:, [1, 2, 0]])


def test_plane_regularisation_equals_the_svd_product(L):
    """U diag(1, 1, 1e-3) V^T of JacobiSVD(cov) (voxel_hash_map.hpp:141-144) for full-rank sample covariances"""
    rng = np.random.default_rng(2)
    for trial in range(300):
        P = rng.normal(size=(int(rng.integers(4, 30)), 3)) * np.array([1.0, 0.5, 10.0 ** rng.integers(-3, 0)])
        A = np.cov((P @ synth.exp_so3(rng.normal(0, 1, 3))).T)
        out, n = np.zeros((3, 3)), np.zeros(3)
        L.orc_k_plane_regularize(d(np.ascontiguousarray(A)), d(out), d(n))
        U, s, Vt = np.linalg.svd(A)
        assert np.abs(out - U @ np.diag([1.0, 1.0, 1e-3]) @ Vt).max() < 1e-9
        assert abs(abs(n @ U[:, 2]) - 1.0) < 1e-9
    # documented conventions for rank-deficient input
    out, n = np.zeros((3, 3)), np.zeros(3)
    L.orc_k_plane_regularize(d(np.zeros((3, 3))), d(out), d(n))
    assert np.array_equal(n, [0.0, 0.0, 1.0]) and np.allclose(out, np.diag([1.0, 1.0, 1e-3]))
    u = np.array([1.0, 2.0, 2.0]) / 3.0
    L.orc_k_plane_regularize(d(np.ascontiguousarray(4.0 * np.outer(u, u))), d(out), d(n))  # rank 1: two points
    assert abs(n @ u) < 1e-12 and abs(np.linalg.norm(n) - 1.0) < 1e-12 and np.allclose(out, np.eye(3) - 0.999 * np.outer(n, n))


def test_angle_axis(L):
    rng = np.random.default_rng(3)
    for trial in range(300):
        w = rng.normal(0, 1, 3) * 10.0 ** rng.integers(-8, 1)
        ang = np.linalg.norm(w)
        R = np.zeros((3, 3))
        L.orc_k_angle_axis_to_rot(ang, d(w / ang), d(R))
        assert np.abs(R - synth.exp_so3(w)).max() < 1e-14
        got = L.orc_k_rot_angle(d(np.ascontiguousarray(R)))
        assert abs(got - (ang if ang <= np.pi else 2 * np.pi - ang)) <= 5e-8  # absolute: a tiny angle loses its low bits in R itself
    R = np.zeros((3, 3))
    L.orc_k_angle_axis_to_rot(0.0, d(np.zeros(3)), d(R))  # zero step: normalized() of the zero vector stays zero
    assert np.array_equal(R, np.eye(3)) and L.orc_k_rot_angle(d(R)) == 0.0
    for ang in (np.pi - 1e-6, np.pi, 3.0):                # near a half turn: the quaternion route stays accurate
        L.orc_k_angle_axis_to_rot(ang, d(np.array([0.0, 0.6, 0.8])), d(R))
        assert abs(L.orc_k_rot_angle(d(np.ascontiguousarray(R))) - ang) < 1e-7
