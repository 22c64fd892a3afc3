// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, load or call anything under oracle/.  The product (elimaloc_b200/) never does.
//
// Parity status (registration + voxel map): PINNED on the reference's own sources.  The reference (jaeyoungjo99/ELiMaLoc
// @ 8254ee7e) ships no tests or golden vectors and its third-party dependencies (Eigen3, oneTBB, PCL, ROS) are absent from
// this image, but registration.cpp + voxel_hash_map.{hpp,cpp} compile UNMODIFIED against the stand-in headers of
// oracle/ref_build/stubs into oracle/_ref/libref.so; tests/test_reference_build.py runs this restatement against that
// library (index-level results identical, floating point to rounding) and tests/golden/*.npz are its outputs.
// What stays restated and unpinned: Eigen's own arithmetic kernels — the routines of THIS file, which the stand-in Eigen
// reuses — and, for a rank-deficient covariance, Eigen's implementation-defined null-space basis (plane_regularize below).
// The EKF (ekf.hpp) is pinned the same way on oracle/_ref/libref_ekf.so (the reference's ekf_algorithm.cpp), and the deskew,
// distance filter, pose interpolation and covariance shaping on oracle/_ref/libref_node.so (the ROS node pcm_matching.cpp
// itself, against stand-in ROS / tf / PCL headers).
//
// Tiny fixed-size fp64 linear algebra that stands in for the Eigen3 calls the reference makes
// (Eigen is a third-party dependency that is absent from /root/reference; apt libeigen3-dev,
// unpinned, Ubuntu 20.04 => 3.3.7).  Every routine names the Eigen call it replaces.
// All matrices are ROW-major here (Eigen is column-major; only the C-ABI cares).
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>

namespace orc {

struct V3 {
    double x = 0, y = 0, z = 0;
    V3() {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator*(double s, const V3& a) { return V3(s * a.x, s * a.y, s * a.z); }
inline double dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// Eigen Vector3d::squaredNorm(): (x^2 + y^2) + z^2 (packet-of-2 redux then the odd tail).
inline double sqnorm(const V3& a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
inline double norm(const V3& a) { return std::sqrt(sqnorm(a)); }

struct M3 {
    double m[9];
    M3() { std::memset(m, 0, sizeof m); }
    static M3 Identity() { M3 r; r.m[0] = r.m[4] = r.m[8] = 1.0; return r; }
    double& operator()(int r, int c) { return m[r * 3 + c]; }
    double operator()(int r, int c) const { return m[r * 3 + c]; }
};
struct M4 {
    double m[16];
    M4() { std::memset(m, 0, sizeof m); }
    static M4 Identity() { M4 r; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0; return r; }
    double& operator()(int r, int c) { return m[r * 4 + c]; }
    double operator()(int r, int c) const { return m[r * 4 + c]; }
};
struct M6 {
    double m[36];
    M6() { std::memset(m, 0, sizeof m); }
    static M6 Identity() { M6 r; for (int i = 0; i < 6; ++i) r.m[i * 7] = 1.0; return r; }
    double& operator()(int r, int c) { return m[r * 6 + c]; }
    double operator()(int r, int c) const { return m[r * 6 + c]; }
};
struct V6 { double v[6] = {0, 0, 0, 0, 0, 0}; };

inline M3 mul(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
    return r;
}
inline M3 transpose(const M3& a) {
    M3 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = a(j, i);
    return r;
}
inline V3 mul(const M3& a, const V3& v) {
    return V3((a(0, 0) * v.x + a(0, 1) * v.y) + a(0, 2) * v.z, (a(1, 0) * v.x + a(1, 1) * v.y) + a(1, 2) * v.z,
              (a(2, 0) * v.x + a(2, 1) * v.y) + a(2, 2) * v.z);
}
inline M4 mul(const M4& a, const M4& b) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r(i, j) = ((a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j)) + a(i, 3) * b(3, j);
    return r;
}
// Matrix4d * Vector4d(p,1) .head<3>()  (reg.hpp:142-145, reg.cpp:29-33): the small fixed-size product
// evaluates column by column, i.e. ((T0*x + T1*y) + T2*z) + T3*1, no FMA (reference builds without -march).
inline V3 apply(const M4& T, const V3& p) {
    return V3(((T(0, 0) * p.x + T(0, 1) * p.y) + T(0, 2) * p.z) + T(0, 3) * 1.0,
              ((T(1, 0) * p.x + T(1, 1) * p.y) + T(1, 2) * p.z) + T(1, 3) * 1.0,
              ((T(2, 0) * p.x + T(2, 1) * p.y) + T(2, 2) * p.z) + T(2, 3) * 1.0);
}
inline M3 rot_of(const M4& T) {
    M3 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = T(i, j);
    return r;
}

// Matrix3d::inverse()  (reg.cpp:79,113,191): cofactor / determinant, as Eigen's compute_inverse<3>.
inline M3 inverse(const M3& a) {
    M3 c;
    c(0, 0) = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
    c(0, 1) = a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2);
    c(0, 2) = a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1);
    c(1, 0) = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2);
    c(1, 1) = a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0);
    c(1, 2) = a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2);
    c(2, 0) = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
    c(2, 1) = a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1);
    c(2, 2) = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
    const double det = (a(0, 0) * c(0, 0) + a(0, 1) * c(1, 0)) + a(0, 2) * c(2, 0);
    const double inv = 1.0 / det;
    M3 r;
    for (int i = 0; i < 9; ++i) r.m[i] = c.m[i] * inv;
    return r;
}

// General NxN inverse by Gauss-Jordan with partial pivoting.  Stands in for Matrix4d::inverse()
// (reg.cpp:24,81,167 — Eigen uses a cofactor kernel for 4x4) and Matrix6d::inverse() (reg.cpp:141 —
// Eigen uses PartialPivLU).  Results agree to rounding for the well-conditioned inputs of this path.
template <int N>
inline void inverse_n(const double* a, double* out) {
    double w[N][2 * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) { w[i][j] = a[i * N + j]; w[i][N + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < N; ++c) {
        int p = c;
        for (int r = c + 1; r < N; ++r) if (std::fabs(w[r][c]) > std::fabs(w[p][c])) p = r;
        if (p != c) for (int j = 0; j < 2 * N; ++j) std::swap(w[p][j], w[c][j]);
        const double d = 1.0 / w[c][c];
        for (int j = 0; j < 2 * N; ++j) w[c][j] *= d;
        for (int r = 0; r < N; ++r) {
            if (r == c) continue;
            const double f = w[r][c];
            if (f == 0.0) continue;
            for (int j = 0; j < 2 * N; ++j) w[r][j] -= f * w[c][j];
        }
    }
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) out[i * N + j] = w[i][N + j];
}
inline M4 inverse(const M4& a) { M4 r; inverse_n<4>(a.m, r.m); return r; }
inline M6 inverse(const M6& a) { M6 r; inverse_n<6>(a.m, r.m); return r; }

// Matrix6d::ldlt().solve(b)  (reg.cpp:56,138,214).  Eigen's LDLT is the robust Cholesky with
// diagonal pivoting (largest remaining |diagonal| first) and, in solve(), a pseudo-inverse of D:
// pivots with |d| <= 1/highest() give a zero component.  Restated here for a symmetric 6x6.
inline V6 ldlt_solve(const M6& A_in, const V6& b) {
    const int N = 6;
    double A[6][6];
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) A[i][j] = A_in(i, j);
    int perm[6];
    for (int i = 0; i < N; ++i) perm[i] = i;
    // in-place LDL^T on the lower triangle with symmetric pivoting
    for (int k = 0; k < N; ++k) {
        int p = k;
        for (int i = k + 1; i < N; ++i) if (std::fabs(A[i][i]) > std::fabs(A[p][p])) p = i;
        if (p != k) {
            for (int j = 0; j < N; ++j) std::swap(A[k][j], A[p][j]);
            for (int i = 0; i < N; ++i) std::swap(A[i][k], A[i][p]);
            std::swap(perm[k], perm[p]);
        }
        // A[k][k] -= sum_{j<k} L[k][j]^2 d_j  (already folded in by the trailing update below)
        const double d = A[k][k];
        if (d == 0.0) { for (int i = k + 1; i < N; ++i) A[i][k] = 0.0; continue; }
        for (int i = k + 1; i < N; ++i) A[i][k] /= d;
        for (int i = k + 1; i < N; ++i)
            for (int j = k + 1; j <= i; ++j) { A[i][j] -= A[i][k] * d * A[j][k]; A[j][i] = A[i][j]; }
    }
    double y[6];
    for (int i = 0; i < N; ++i) y[i] = b.v[perm[i]];
    for (int i = 0; i < N; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
    const double tol = 1.0 / 1.7976931348623157e308;
    for (int i = 0; i < N; ++i) y[i] = (std::fabs(A[i][i]) > tol) ? y[i] / A[i][i] : 0.0;
    for (int i = N - 1; i >= 0; --i) for (int j = i + 1; j < N; ++j) y[i] -= A[j][i] * y[j];
    V6 x;
    for (int i = 0; i < N; ++i) x.v[perm[i]] = y[i];
    return x;
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi, sweep order (0,1),(0,2),(1,2).
// Returns eigenvalues ASCENDING (selection sort that keeps the first of equal values, as Eigen's
// SelfAdjointEigenSolver does) and eigenvectors as COLUMNS of V.  Stands in for
// SelfAdjointEigenSolver<Matrix3d> (reg.cpp:89-91) and, via plane_regularize(), JacobiSVD<Matrix3d>
// (vhm.hpp:141-144, 241-244).  An exactly diagonal input yields V = I untouched.
inline void sym_eig3(const M3& A_in, double w[3], M3& V) {
    double A[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = 0.5 * (A_in(i, j) + A_in(j, i));
    V = M3::Identity();
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        if (off == 0.0) break;
        const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
        for (int e = 0; e < 3; ++e) {
            const int p = P[e], q = Q[e];
            const double apq = A[p][q];
            if (apq == 0.0) continue;
            const double app = A[p][p], aqq = A[q][q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; ++k) {  // A <- A * J
                const double akp = A[k][p], akq = A[k][q];
                A[k][p] = c * akp - s * akq;
                A[k][q] = s * akp + c * akq;
            }
            for (int k = 0; k < 3; ++k) {  // A <- J^T * A
                const double apk = A[p][k], aqk = A[q][k];
                A[p][k] = c * apk - s * aqk;
                A[q][k] = s * apk + c * aqk;
            }
            A[p][q] = A[q][p] = 0.0;
            for (int k = 0; k < 3; ++k) {  // V <- V * J
                const double vkp = V(k, p), vkq = V(k, q);
                V(k, p) = c * vkp - s * vkq;
                V(k, q) = s * vkp + c * vkq;
            }
        }
    }
    w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
    for (int i = 0; i < 2; ++i) {
        int k = i;
        for (int j = i + 1; j < 3; ++j) if (w[j] < w[k]) k = j;
        if (k != i) {
            std::swap(w[i], w[k]);
            for (int r = 0; r < 3; ++r) std::swap(V(r, i), V(r, k));
        }
    }
}

// "Plane regularisation"  cov <- U * diag(1,1,1e-3) * V^T  from JacobiSVD(cov)  (vhm.hpp:141-144).
// For a symmetric PSD input U == V and the product equals  I - (1 - 1e-3) * n n^T  with n the unit
// singular vector of the SMALLEST singular value; the two leading vectors cancel out of the result.
// n is unique (up to sign, which also cancels) unless the two smallest singular values coincide
// (rank-deficient sample covariance: 2 points, {self,self}, collinear sets).  There Eigen's result is
// implementation-defined; ORACLE CONVENTION (shared with the product, documented in DESIGN.md):
//   * all three equal within tol (incl. the zero matrix)      -> n = e_z     (Eigen gives U = I here)
//   * two smallest equal within tol, dominant direction u1    -> n = normalise(e_k - (e_k.u1) u1),
//     k = axis with the smallest |u1_k| (lowest k on ties; components within 1e-9 of each other are tied)
// tol = 1e-9 * max(lambda_max, 1e-300).
inline M3 plane_regularize(const M3& cov, V3* normal_out = nullptr) {
    double w[3];
    M3 V;
    sym_eig3(cov, w, V);  // ascending
    const double lmax = std::max(std::fabs(w[2]), std::fabs(w[0]));
    const double tol = 1e-9 * std::max(lmax, 1e-300);
    V3 n;
    if (std::fabs(w[2] - w[0]) <= tol) {
        n = V3(0, 0, 1);
    } else if (std::fabs(w[1] - w[0]) <= tol) {
        const V3 u1(V(0, 2), V(1, 2), V(2, 2));
        const double kAxisTie = 1e-9;  // |u1_k| within 1e-9 of each other are tied: the lowest axis wins (robust to the last bits of u1)
        int k = 0;
        double best = std::fabs(u1.x);
        if (std::fabs(u1.y) < best - kAxisTie) { best = std::fabs(u1.y); k = 1; }
        if (std::fabs(u1.z) < best - kAxisTie) { best = std::fabs(u1.z); k = 2; }
        V3 e(k == 0, k == 1, k == 2);
        const double d = dot(e, u1);
        V3 v = e - d * u1;
        const double l = norm(v);
        n = V3(v.x / l, v.y / l, v.z / l);
    } else {
        n = V3(V(0, 0), V(1, 0), V(2, 0));
    }
    if (normal_out) *normal_out = n;
    M3 r = M3::Identity();
    const double k = 1.0 - 1e-3;
    const double nv[3] = {n.x, n.y, n.z};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) -= k * nv[i] * nv[j];
    return r;
}

// AngleAxisd(angle, axis).toRotationMatrix()  (reg.cpp:60-61), Eigen's formula verbatim in meaning.
inline M3 angle_axis_to_rot(double angle, const V3& axis) {
    M3 res;
    const double s = std::sin(angle), c = std::cos(angle);
    const V3 sin_axis = s * axis;
    const V3 cos1_axis = (1.0 - c) * axis;
    double tmp;
    tmp = cos1_axis.x * axis.y; res(0, 1) = tmp - sin_axis.z; res(1, 0) = tmp + sin_axis.z;
    tmp = cos1_axis.x * axis.z; res(0, 2) = tmp + sin_axis.y; res(2, 0) = tmp - sin_axis.y;
    tmp = cos1_axis.y * axis.z; res(1, 2) = tmp - sin_axis.x; res(2, 1) = tmp + sin_axis.x;
    res(0, 0) = cos1_axis.x * axis.x + c;
    res(1, 1) = cos1_axis.y * axis.y + c;
    res(2, 2) = cos1_axis.z * axis.z + c;
    return res;
}
// Vector3d::normalized(): zero vector stays zero (Eigen 3.3: divides only when the norm is > 0).
inline V3 normalized(const V3& v) {
    const double z = sqnorm(v);
    if (z > 0.0) { const double l = std::sqrt(z); return V3(v.x / l, v.y / l, v.z / l); }
    return v;
}
// AngleAxisd(Matrix3d).angle()  (reg.cpp:381-382): matrix -> quaternion (Shepperd) -> angle = 2 atan2(|vec|, |w|).
inline double rot_angle(const M3& m) {
    double qw, qx, qy, qz;
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        qw = 0.5 * t;
        t = 0.5 / t;
        qx = (m(2, 1) - m(1, 2)) * t;
        qy = (m(0, 2) - m(2, 0)) * t;
        qz = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        double q[3];
        q[i] = 0.5 * t;
        t = 0.5 / t;
        qw = (m(k, j) - m(j, k)) * t;
        q[j] = (m(j, i) + m(i, j)) * t;
        q[k] = (m(k, i) + m(i, k)) * t;
        qx = q[0]; qy = q[1]; qz = q[2];
    }
    const double n = std::sqrt((qx * qx + qy * qy) + qz * qz);
    if (n != 0.0) return 2.0 * std::atan2(n, std::fabs(qw));
    return 0.0;
}

}  // namespace orc
