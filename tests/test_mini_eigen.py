"""The stand-in Eigen that lets the reference's sources compile here (oracle/ref_build/stubs/mini_eigen.hpp, test infrastructure)
is itself checked: tests/mini_eigen_check.cpp runs its operations on fixed inputs, this test recomputes each result with numpy.
Expression semantics (products, blocks, comma initialiser, transposes, diagonal products, 3 x Dynamic statistics), the
decompositions it delegates to oracle/smallmat.hpp, the quaternion / angle-axis / affine algebra and the truncating int cast."""
import os
import subprocess

import numpy as np

from elimaloc_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def qmul(a, b):
    return np.array([a[0] * b[0] - a[1:] @ b[1:], *(a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:]))])


def test_stand_in_eigen_against_numpy(tmp_path):
    exe = str(tmp_path / "mini_eigen_check")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-I" + os.path.join(ROOT, "oracle", "ref_build", "stubs"),
                    os.path.join(ROOT, "tests", "mini_eigen_check.cpp"), "-o", exe], check=True, capture_output=True, text=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    r = {ln.split()[0]: np.array([float(x) for x in ln.split()[1:]]) for ln in out.strip().split("\n")}
    A = r["A"].reshape(3, 3)
    T = r["T"].reshape(4, 4)
    v, w = np.array([0.3, -1.2, 2.5]), np.array([1.0, 0.5, -0.25])
    close = lambda name, want, tol=1e-13: np.testing.assert_allclose(r[name].reshape(np.shape(want)), want, rtol=tol, atol=tol, err_msg=name)  # noqa: E731
    assert np.array_equal(A, [[2.0, -1.0, 0.5], [0.25, 3.0, -0.75], [1.5, 0.125, 4.0]])                 # comma initialiser: row by row
    Rz, Ry, Rx = synth.exp_so3([0, 0, 0.3]), synth.exp_so3([0, -0.2, 0]), synth.exp_so3([0.1, 0, 0])
    close("T", np.block([[Rz @ Ry @ Rx, np.array([[1.5], [-2.5], [0.75]])], [np.zeros((1, 3)), np.ones((1, 1))]]))
    close("A_times_At", A @ A.T)
    close("A_inverse", np.linalg.inv(A))
    close("T_inverse", np.linalg.inv(T))
    close("T_times_Tinv", np.eye(4))
    close("scaled_sum", 2.5 * A + A * 0.5 - A / 4.0)
    close("A_v", A @ v)
    close("vT_A", v @ A)
    close("cross", np.cross(v, w))
    close("dot_norm", [v @ w, np.linalg.norm(v), v @ v, np.trace(A)])
    close("normalized", v / np.linalg.norm(v))
    close("diag_product", A @ np.diag(v))
    S = r["S"].reshape(6, 6)
    want_S = np.block([[A @ A.T, 0.1 * A], [0.1 * A.T, A.T @ A + np.eye(3)]])
    close("S", want_S)                                                                                  # block assignment
    b = np.array([1.0, -2.0, 3.0, 0.5, 0.25, -1.5])
    close("S_inverse", np.linalg.inv(S), 1e-12)
    close("ldlt_solve", np.linalg.solve(S, b), 1e-12)
    close("tail_head", b[3:] - b[:3])
    N = np.stack([v, w, v + w, 2 * w - v], axis=1)
    close("mean", N.mean(axis=1))
    close("sample_cov", np.cov(N, ddof=1))
    close("eigenvalues", np.linalg.eigvalsh(A @ A.T), 1e-12)
    close("eig_residual", np.zeros((3, 3)), 1e-11)
    U, s, Vt = np.linalg.svd(A @ A.T)
    close("plane", U @ np.diag([1, 1, 1e-3]) @ Vt, 1e-11)
    axis = np.array([0.0, 0.6, 0.8])
    q1 = np.array([np.cos(0.35), *(np.sin(0.35) * axis)])
    close("q_from_angle_axis", q1)
    q2 = r["q_from_matrix"]
    np.testing.assert_allclose(quat_to_R(q2), T[:3, :3], atol=1e-14)                                    # matrix -> quaternion
    p = qmul(q1, q2)
    close("q_product_normalized", p / np.linalg.norm(p))
    close("q_inverse", np.array([q2[0], -q2[1], -q2[2], -q2[3]]) / (q2 @ q2))
    close("q_rotate", quat_to_R(q1) @ v)
    close("q_to_matrix", quat_to_R(q1))
    close("angle_of_matrix", [np.arccos((np.trace(T[:3, :3]) - 1) / 2)], 1e-12)
    qf = np.array([0.9, 0.1, -0.3, 0.2], np.float32)
    qf = qf / np.linalg.norm(qf)
    th = np.arccos(qf[0])                                                                               # angle to the identity quaternion
    want = (np.sin(0.7 * th) * np.array([1, 0, 0, 0]) + np.sin(0.3 * th) * qf) / np.sin(th)
    close("slerp", want, 1e-6)
    a1 = np.eye(4)
    a1[:3, :3], a1[:3, 3] = quat_to_R(qf.astype(np.float64)), [1, 2, 3]
    a2 = np.eye(4)
    a2[:3, :3], a2[:3, 3] = synth.exp_so3([0, 0, 0.4]), [0.5, -0.5, 0.25]
    close("affine_between", np.linalg.inv(a1) @ a2, 1e-6)
    assert list(r["cast_int"]) == [0, 1, -2]                                                            # truncation toward zero, not floor
