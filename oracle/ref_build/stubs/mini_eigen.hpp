// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// Minimal stand-in for the slice of Eigen3 that the reference's registration path uses
//   pcm_matching/include/registration.hpp, voxel_hash_map.hpp, pcm_matching/src/registration.cpp, voxel_hash_map.cpp
// so that those four files compile UNMODIFIED, from where they lie under /root/reference, into oracle/_ref/libref.so
// (Eigen3 itself is a third-party dependency that is absent from the reference tree and from this image).
//
// What this gives: the reference's own control flow — insertion keys, spacing test, visit order, tie-breaks, the three
// correspondence searches, the weights and early-outs of the AlignClouds* loops, the LM step, the RunRegister loop — runs as
// written.  What it does NOT give: Eigen's own floating-point kernels.  Expressions are evaluated eagerly, left to right,
// with plain ascending-k dot products, and the decompositions (inverse, LDLT, SelfAdjointEigenSolver, JacobiSVD, AngleAxis)
// are the restatements of oracle/smallmat.hpp.  So a comparison of the oracle with oracle/_ref pins the oracle's
// ALGORITHM to the reference sources; third-party arithmetic stays restated (rounding-level differences only).
//
// Storage is column-major like Eigen's default.  Only what the four files need exists; anything else fails to compile.
#pragma once
#include <cmath>
#include <cstddef>
#include <limits>
#include <ostream>
#include <type_traits>
#include <vector>

#include "../../smallmat.hpp"

namespace Eigen {

constexpr int Dynamic = -1;
enum DecompositionOptions { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };

template <typename T, int R, int C> class Matrix;
template <typename T, int N> struct DiagonalWrapper { Matrix<T, N, 1> d; };
template <typename T, int BR, int BC, int R, int C> class BlockRef;
template <typename T, int N> class LDLT;
template <typename T> class Quaternion;
struct ScaledIdentityXd { int n; double s; };  // MatrixXd::Identity(n, n) * s  (ekf_algorithm.cpp:41)

// Records every ldlt().solve(A, b) when enabled: lets the test driver read the normal equations (JTJ + lambda diag, JTr)
// that the reference's AlignClouds* functions keep in locals.
struct LdltTap {
    bool enabled = false;
    std::vector<std::vector<double>> A, b;  // row-major A
    static LdltTap& get() { static thread_local LdltTap t; return t; }
};

template <typename M> class CommaInit {
public:
    CommaInit(M& m, typename M::Scalar v) : m_(m), k_(0) { put(v); }
    template <typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
    CommaInit& operator,(S v) { put(static_cast<typename M::Scalar>(v)); return *this; }
private:
    void put(typename M::Scalar v) { m_(k_ / M::Cols, k_ % M::Cols) = v; ++k_; }  // row by row, as Eigen fills
    M& m_;
    int k_;
};

template <typename T, int R, int C>
class Matrix {
    static_assert(R > 0 && C > 0, "fixed sizes only (3 x Dynamic has its own specialisation)");
public:
    using Scalar = T;
    static constexpr int Rows = R, Cols = C, Size = R * C;

    Matrix() { for (int i = 0; i < Size; ++i) m_[i] = T(0); }
    template <typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
    Matrix(A a, B b) { static_assert(Size == 2, "2-vector"); m_[0] = T(a); m_[1] = T(b); }
    template <typename A, typename B, typename D>
    Matrix(A a, B b, D c) { static_assert(Size == 3 && (R == 1 || C == 1), "3-vector"); m_[0] = T(a); m_[1] = T(b); m_[2] = T(c); }
    template <typename A, typename B, typename D, typename E>
    Matrix(A a, B b, D c, E d) { static_assert(Size == 4 && (R == 1 || C == 1), "4-vector"); m_[0] = T(a); m_[1] = T(b); m_[2] = T(c); m_[3] = T(d); }
    Matrix(const DiagonalWrapper<T, R>& w) {
        static_assert(R == C, "square");
        for (int i = 0; i < Size; ++i) m_[i] = T(0);
        for (int i = 0; i < R; ++i) (*this)(i, i) = w.d(i);
    }

    Matrix(const Quaternion<T>& q);  // 3x3 only: rotation matrix of a quaternion (Eigen: RotationBase assignment)
    Matrix(const ScaledIdentityXd& w) {
        static_assert(R == C, "square");
        for (int i = 0; i < Size; ++i) m_[i] = T(0);
        for (int i = 0; i < R; ++i) (*this)(i, i) = static_cast<T>(w.s);
    }

    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() { Matrix r; for (int i = 0; i < (R < C ? R : C); ++i) r(i, i) = T(1); return r; }
    static Matrix UnitX() { Matrix r; r.m_[0] = T(1); return r; }
    static Matrix UnitY() { Matrix r; r.m_[1] = T(1); return r; }
    static Matrix UnitZ() { Matrix r; r.m_[2] = T(1); return r; }

    T& operator()(int r, int c) { return m_[c * R + r]; }
    const T& operator()(int r, int c) const { return m_[c * R + r]; }
    T& operator()(int i) { static_assert(R == 1 || C == 1, "vector"); return m_[i]; }
    const T& operator()(int i) const { static_assert(R == 1 || C == 1, "vector"); return m_[i]; }
    T& operator[](int i) { return (*this)(i); }
    const T& operator[](int i) const { return (*this)(i); }
    T& x() { return m_[0]; }
    T& y() { return m_[1]; }
    T& z() { return m_[2]; }
    T& w() { return m_[3]; }
    const T& x() const { return m_[0]; }
    const T& y() const { return m_[1]; }
    const T& z() const { return m_[2]; }
    const T& w() const { return m_[3]; }
    T* data() { return m_; }
    const T* data() const { return m_; }

    void setZero() { for (int i = 0; i < Size; ++i) m_[i] = T(0); }
    void setIdentity() { *this = Identity(); }
    T determinant() const {
        static_assert(R == 3 && C == 3, "3x3");
        const Matrix& a = *this;
        return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) + a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
    }
    T minCoeff() const { T v = m_[0]; for (int i = 1; i < Size; ++i) if (m_[i] < v) v = m_[i]; return v; }
    Matrix cwiseMin(T s) const { Matrix r; for (int i = 0; i < Size; ++i) r.m_[i] = m_[i] < s ? m_[i] : s; return r; }
    Matrix cwiseMax(T s) const { Matrix r; for (int i = 0; i < Size; ++i) r.m_[i] = m_[i] > s ? m_[i] : s; return r; }
    template <typename F> Matrix unaryExpr(F f) const { Matrix r; for (int i = 0; i < Size; ++i) r.m_[i] = f(m_[i]); return r; }
    template <typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>> Matrix& operator*=(S s) { for (int i = 0; i < Size; ++i) m_[i] = m_[i] * static_cast<T>(s); return *this; }
    template <typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>> Matrix& operator/=(S s) { for (int i = 0; i < Size; ++i) m_[i] = m_[i] / static_cast<T>(s); return *this; }
    BlockRef<T, R, 1, R, C> col(int j) { return BlockRef<T, R, 1, R, C>(*this, 0, j); }
    T trace() const { T s = (*this)(0, 0); for (int i = 1; i < (R < C ? R : C); ++i) s = s + (*this)(i, i); return s; }
    Matrix cross(const Matrix& o) const {
        static_assert(Size == 3, "3-vector");
        return Matrix(m_[1] * o.m_[2] - m_[2] * o.m_[1], m_[2] * o.m_[0] - m_[0] * o.m_[2], m_[0] * o.m_[1] - m_[1] * o.m_[0]);
    }
    Matrix& noalias() { return *this; }
    const Matrix& matrix() const { return *this; }

    CommaInit<Matrix> operator<<(T v) { return CommaInit<Matrix>(*this, v); }

    template <typename U> Matrix<U, R, C> cast() const {
        Matrix<U, R, C> r;
        for (int i = 0; i < Size; ++i) r.data()[i] = static_cast<U>(m_[i]);  // double -> int truncates toward zero
        return r;
    }

    T squaredNorm() const { T s = m_[0] * m_[0]; for (int i = 1; i < Size; ++i) s = s + m_[i] * m_[i]; return s; }
    T norm() const { return std::sqrt(squaredNorm()); }
    Matrix normalized() const {  // Eigen 3.3: divide only when the squared norm is > 0
        const T z = squaredNorm();
        if (z > T(0)) return *this / std::sqrt(z);
        return *this;
    }
    void normalize() { *this = normalized(); }
    T dot(const Matrix& o) const { T s = m_[0] * o.m_[0]; for (int i = 1; i < Size; ++i) s = s + m_[i] * o.m_[i]; return s; }

    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> r;
        for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) r(j, i) = (*this)(i, j);
        return r;
    }
    Matrix<T, (R < C ? R : C), 1> diagonal() const {
        Matrix<T, (R < C ? R : C), 1> d;
        for (int i = 0; i < (R < C ? R : C); ++i) d(i) = (*this)(i, i);
        return d;
    }
    DiagonalWrapper<T, Size> asDiagonal() const { static_assert(R == 1 || C == 1, "vector"); DiagonalWrapper<T, Size> w; for (int i = 0; i < Size; ++i) w.d(i) = m_[i]; return w; }

    template <int N> Matrix<T, N, 1> head() const { static_assert(C == 1, "column vector"); Matrix<T, N, 1> r; for (int i = 0; i < N; ++i) r(i) = m_[i]; return r; }
    template <int N> Matrix<T, N, 1> tail() const { static_assert(C == 1, "column vector"); Matrix<T, N, 1> r; for (int i = 0; i < N; ++i) r(i) = m_[R - N + i]; return r; }
    template <int N> Matrix<T, N, 1> segment(int i) const { static_assert(C == 1, "column vector"); Matrix<T, N, 1> r; for (int k = 0; k < N; ++k) r(k) = m_[i + k]; return r; }
    template <int N> BlockRef<T, N, 1, R, C> head() { static_assert(C == 1, "column vector"); return BlockRef<T, N, 1, R, C>(*this, 0, 0); }
    template <int N> BlockRef<T, N, 1, R, C> tail() { static_assert(C == 1, "column vector"); return BlockRef<T, N, 1, R, C>(*this, R - N, 0); }
    template <int N> BlockRef<T, N, 1, R, C> segment(int i) { static_assert(C == 1, "column vector"); return BlockRef<T, N, 1, R, C>(*this, i, 0); }
    Matrix<T, R, 1> col(int j) const { Matrix<T, R, 1> r; for (int i = 0; i < R; ++i) r(i) = (*this)(i, j); return r; }

    template <int BR, int BC> BlockRef<T, BR, BC, R, C> block(int i, int j) { return BlockRef<T, BR, BC, R, C>(*this, i, j); }
    template <int BR, int BC> Matrix<T, BR, BC> block(int i, int j) const {
        Matrix<T, BR, BC> r;
        for (int a = 0; a < BR; ++a) for (int b = 0; b < BC; ++b) r(a, b) = (*this)(i + a, j + b);
        return r;
    }

    Matrix inverse() const;  // 3x3 / 4x4 / 6x6, defined below
    LDLT<T, R> ldlt() const { static_assert(R == C, "square"); return LDLT<T, R>(*this); }

    Matrix& operator+=(const Matrix& o) { for (int i = 0; i < Size; ++i) m_[i] = m_[i] + o.m_[i]; return *this; }
    Matrix& operator-=(const Matrix& o) { for (int i = 0; i < Size; ++i) m_[i] = m_[i] - o.m_[i]; return *this; }
    Matrix operator-() const { Matrix r; for (int i = 0; i < Size; ++i) r.m_[i] = -m_[i]; return r; }
    bool operator==(const Matrix& o) const { for (int i = 0; i < Size; ++i) if (!(m_[i] == o.m_[i])) return false; return true; }
    bool operator!=(const Matrix& o) const { return !(*this == o); }

private:
    T m_[Size];
};

template <typename T, int R, int C> Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> r = a; r += b; return r; }
template <typename T, int R, int C> Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> r = a; r -= b; return r; }
template <typename T, int R, int C, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
Matrix<T, R, C> operator*(S s, const Matrix<T, R, C>& a) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.data()[i] = static_cast<T>(s) * a.data()[i]; return r; }
template <typename T, int R, int C, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, S s) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.data()[i] = a.data()[i] * static_cast<T>(s); return r; }
template <typename T, int R, int C, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, S s) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.data()[i] = a.data()[i] / static_cast<T>(s); return r; }
// coefficient (i, j) = ((a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j) + ...   — no FMA (build with -ffp-contract=off)
template <typename T, int R, int K, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const Matrix<T, K, C>& b) {
    Matrix<T, R, C> r;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) {
            T s = a(i, 0) * b(0, j);
            for (int k = 1; k < K; ++k) s = s + a(i, k) * b(k, j);
            r(i, j) = s;
        }
    return r;
}
template <typename T, int R, int N>
Matrix<T, R, N> operator*(const Matrix<T, R, N>& a, const DiagonalWrapper<T, N>& w) {
    Matrix<T, R, N> r;
    for (int i = 0; i < R; ++i) for (int j = 0; j < N; ++j) r(i, j) = a(i, j) * w.d(j);
    return r;
}

template <typename T, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& m) {  // debug prints only
    for (int i = 0; i < R; ++i) { for (int j = 0; j < C; ++j) os << (j ? " " : "") << m(i, j); if (i + 1 < R) os << "\n"; }
    return os;
}

// Writable view of a fixed block of a fixed matrix; reads convert to a Matrix value.
template <typename T, int BR, int BC, int R, int C>
class BlockRef {
public:
    using Scalar = T;
    static constexpr int Rows = BR, Cols = BC;
    BlockRef(Matrix<T, R, C>& m, int i, int j) : m_(m), i_(i), j_(j) {}
    T& operator()(int a, int b) { return m_(i_ + a, j_ + b); }
    CommaInit<BlockRef> operator<<(T v) { return CommaInit<BlockRef>(*this, v); }
    template <int N> Matrix<T, N, 1> head() const { static_assert(BC == 1, "column vector"); Matrix<T, N, 1> r; for (int k = 0; k < N; ++k) r(k) = m_(i_ + k, j_); return r; }
    Matrix<T, BR, BC> operator-() const { return -eval(); }
    // same shape, or — as Eigen allows for vectors — a column vector into a row-vector block (reg.cpp:252 assigns a
    // Vector3d to a 1x3 block)
    template <int VR, int VC>
    BlockRef& operator=(const Matrix<T, VR, VC>& v) {
        static_assert((VR == BR && VC == BC) || ((BR == 1 || BC == 1) && VR == BC && VC == BR), "block assignment: shape mismatch");
        if (VR == BR && VC == BC) {
            for (int a = 0; a < BR; ++a) for (int b = 0; b < BC; ++b) m_(i_ + a, j_ + b) = v.data()[b * VR + a];
        } else {
            for (int k = 0; k < BR * BC; ++k) m_(i_ + (BR == 1 ? 0 : k), j_ + (BR == 1 ? k : 0)) = v.data()[k];
        }
        return *this;
    }
    operator Matrix<T, BR, BC>() const { return eval(); }
    Matrix<T, BR, BC> eval() const {
        Matrix<T, BR, BC> r;
        for (int a = 0; a < BR; ++a) for (int b = 0; b < BC; ++b) r(a, b) = m_(i_ + a, j_ + b);
        return r;
    }
    T norm() const { return eval().norm(); }
    T squaredNorm() const { return eval().squaredNorm(); }
    Matrix<T, BC, BR> transpose() const { return eval().transpose(); }
    BlockRef& operator=(const BlockRef& o) { return *this = o.eval(); }
    template <int R2, int C2> BlockRef& operator=(const BlockRef<T, BR, BC, R2, C2>& o) { return *this = o.eval(); }
    BlockRef& operator+=(const Matrix<T, BR, BC>& v) { return *this = eval() + v; }
    BlockRef& operator-=(const Matrix<T, BR, BC>& v) { return *this = eval() - v; }
    template <int K> friend Matrix<T, K, BC> operator*(const Matrix<T, K, BR>& a, const BlockRef& b) { return a * b.eval(); }
    template <int K> friend Matrix<T, BR, K> operator*(const BlockRef& a, const Matrix<T, BC, K>& b) { return a.eval() * b; }
    friend Matrix<T, BR, BC> operator+(const BlockRef& a, const Matrix<T, BR, BC>& b) { return a.eval() + b; }
    friend Matrix<T, BR, BC> operator+(const Matrix<T, BR, BC>& a, const BlockRef& b) { return a + b.eval(); }
    friend Matrix<T, BR, BC> operator-(const BlockRef& a, const Matrix<T, BR, BC>& b) { return a.eval() - b; }
    friend Matrix<T, BR, BC> operator-(const Matrix<T, BR, BC>& a, const BlockRef& b) { return a - b.eval(); }
    template <int R2, int C2> Matrix<T, BR, BC> operator+(const BlockRef<T, BR, BC, R2, C2>& b) const { return eval() + b.eval(); }
    template <int R2, int C2> Matrix<T, BR, BC> operator-(const BlockRef<T, BR, BC, R2, C2>& b) const { return eval() - b.eval(); }
private:
    Matrix<T, R, C>& m_;
    int i_, j_;
};

// ---- 3 x Dynamic (the neighbour matrices of CalVoxelCov / ProcessVoxelBlock / FindGroundHeight) ----------------------
template <typename T>
class Matrix<T, 3, Dynamic> {
public:
    using Scalar = T;
    Matrix(int rows, std::size_t cols) : cols_(cols), d_(3 * cols, T(0)) { (void)rows; }
    std::size_t cols() const { return cols_; }
    T& operator()(int r, std::size_t c) { return d_[3 * c + r]; }
    const T& operator()(int r, std::size_t c) const { return d_[3 * c + r]; }

    struct ColRef {
        Matrix& m; std::size_t j;
        ColRef& operator=(const Matrix<T, 3, 1>& v) { for (int r = 0; r < 3; ++r) m(r, j) = v(r); return *this; }
        operator Matrix<T, 3, 1>() const { return Matrix<T, 3, 1>(m(0, j), m(1, j), m(2, j)); }
    };
    ColRef col(std::size_t j) { return ColRef{*this, j}; }

    struct Rowwise {
        const Matrix& m;
        Matrix<T, 3, 1> sum() const {  // columns added in ascending order
            Matrix<T, 3, 1> s;
            for (std::size_t j = 0; j < m.cols(); ++j) for (int r = 0; r < 3; ++r) s(r) = s(r) + m(r, j);
            return s;
        }
        Matrix<T, 3, 1> mean() const { return sum() / static_cast<T>(m.cols()); }
    };
    Rowwise rowwise() const { return Rowwise{*this}; }

    struct Colwise {
        Matrix& m;
        void operator-=(const Matrix<T, 3, 1>& v) { for (std::size_t j = 0; j < m.cols(); ++j) for (int r = 0; r < 3; ++r) m(r, j) = m(r, j) - v(r); }
    };
    Colwise colwise() { return Colwise{*this}; }

    struct Transposed { const Matrix& m; };
    Transposed transpose() const { return Transposed{*this}; }
    friend Matrix<T, 3, 3> operator*(const Matrix& a, const Transposed& bt) {  // sum over the columns in ascending order
        Matrix<T, 3, 3> r;
        for (std::size_t j = 0; j < a.cols(); ++j)
            for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) r(p, q) = r(p, q) + a(p, j) * bt.m(q, j);
        return r;
    }
private:
    std::size_t cols_;
    std::vector<T> d_;
};

// ---- conversions to / from the row-major helpers of oracle/smallmat.hpp ------------------------------------------------
namespace stub {
inline orc::M3 to_orc(const Matrix<double, 3, 3>& a) { orc::M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = a(i, j); return r; }
inline orc::M4 to_orc(const Matrix<double, 4, 4>& a) { orc::M4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r(i, j) = a(i, j); return r; }
inline orc::M6 to_orc(const Matrix<double, 6, 6>& a) { orc::M6 r; for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) r(i, j) = a(i, j); return r; }
inline Matrix<double, 3, 3> from_orc(const orc::M3& a) { Matrix<double, 3, 3> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = a(i, j); return r; }
inline Matrix<double, 4, 4> from_orc(const orc::M4& a) { Matrix<double, 4, 4> r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r(i, j) = a(i, j); return r; }
inline Matrix<double, 6, 6> from_orc(const orc::M6& a) { Matrix<double, 6, 6> r; for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) r(i, j) = a(i, j); return r; }
inline orc::V3 to_orc(const Matrix<double, 3, 1>& v) { return orc::V3(v(0), v(1), v(2)); }
}  // namespace stub

template <typename T, int R, int C>
Matrix<T, R, C> Matrix<T, R, C>::inverse() const {
    static_assert(R == C, "inverse(): square matrices only");
    if constexpr (R == 3 && std::is_same<T, double>::value) {
        return stub::from_orc(orc::inverse(stub::to_orc(*this)));  // cofactor formula
    } else if constexpr (R == 3) {  // float: the same cofactor formula
        const Matrix& a = *this;
        Matrix c;
        c(0, 0) = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1); c(0, 1) = a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2); c(0, 2) = a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1);
        c(1, 0) = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2); c(1, 1) = a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0); c(1, 2) = a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2);
        c(2, 0) = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0); c(2, 1) = a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1); c(2, 2) = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
        const T det = (a(0, 0) * c(0, 0) + a(0, 1) * c(1, 0)) + a(0, 2) * c(2, 0);
        return c * (T(1) / det);
    } else {
        static_assert(std::is_same<T, double>::value, "N x N inverse: double only");  // Gauss-Jordan with partial pivoting (oracle/smallmat.hpp inverse_n)
        double a[R * R], out[R * R];
        for (int i = 0; i < R; ++i) for (int j = 0; j < R; ++j) a[i * R + j] = (*this)(i, j);
        orc::inverse_n<R>(a, out);
        Matrix r;
        for (int i = 0; i < R; ++i) for (int j = 0; j < R; ++j) r(i, j) = out[i * R + j];
        return r;
    }
}

template <typename T, int N>
class LDLT {
public:
    explicit LDLT(const Matrix<T, N, N>& a) : a_(a) {}
    Matrix<T, N, 1> solve(const Matrix<T, N, 1>& b) const {
        static_assert(std::is_same<T, double>::value && N == 6, "ldlt().solve(): double 6x6 only");
        orc::V6 rhs;
        for (int i = 0; i < 6; ++i) rhs.v[i] = b(i);
        LdltTap& tap = LdltTap::get();
        if (tap.enabled) {
            std::vector<double> A(36), B(6);
            for (int i = 0; i < 6; ++i) { B[i] = b(i); for (int j = 0; j < 6; ++j) A[6 * i + j] = a_(i, j); }
            tap.A.push_back(A);
            tap.b.push_back(B);
        }
        const orc::V6 x = orc::ldlt_solve(stub::to_orc(a_), rhs);
        Matrix<T, N, 1> r;
        for (int i = 0; i < 6; ++i) r(i) = x.v[i];
        return r;
    }
private:
    Matrix<T, N, N> a_;
};

template <typename M> class SelfAdjointEigenSolver;
template <> class SelfAdjointEigenSolver<Matrix<double, 3, 3>> {
public:
    explicit SelfAdjointEigenSolver(const Matrix<double, 3, 3>& a) {
        double w[3];
        orc::M3 V;
        orc::sym_eig3(stub::to_orc(a), w, V);  // ascending, eigenvectors in columns
        vec_ = stub::from_orc(V);
        val_ = Matrix<double, 3, 1>(w[0], w[1], w[2]);
    }
    const Matrix<double, 3, 3>& eigenvectors() const { return vec_; }
    const Matrix<double, 3, 1>& eigenvalues() const { return val_; }
private:
    Matrix<double, 3, 3> vec_;
    Matrix<double, 3, 1> val_;
};

// The reference only ever takes JacobiSVD of a symmetric PSD sample covariance and forms U diag(1,1,1e-3) V^T from it
// (vhm.hpp:141-144, 241-244).  For such an input U = V = eigenvectors by descending eigenvalue; the third column is the
// plane normal, chosen by the oracle's documented convention when the two smallest singular values coincide
// (orc::plane_regularize), and the first two columns are an orthonormal completion (their choice cancels in the product).
template <typename M> class JacobiSVD;
template <> class JacobiSVD<Matrix<double, 3, 3>> {
public:
    JacobiSVD(const Matrix<double, 3, 3>& a, unsigned int /*options*/ = 0) {
        orc::V3 n;
        (void)orc::plane_regularize(stub::to_orc(a), &n);
        // orthonormal completion: e_k with the smallest |n_k|, Gram-Schmidt, then the cross product
        int k = 0;
        if (std::fabs(n.y) < std::fabs(n[k])) k = 1;
        if (std::fabs(n.z) < std::fabs(n[k])) k = 2;
        const orc::V3 e(k == 0, k == 1, k == 2);
        const orc::V3 t = e - orc::dot(e, n) * n;
        const double l = orc::norm(t);
        const orc::V3 u1(t.x / l, t.y / l, t.z / l);
        const orc::V3 u2(n.y * u1.z - n.z * u1.y, n.z * u1.x - n.x * u1.z, n.x * u1.y - n.y * u1.x);
        const orc::V3 cols[3] = {u1, u2, n};
        for (int c = 0; c < 3; ++c) { u_(0, c) = cols[c].x; u_(1, c) = cols[c].y; u_(2, c) = cols[c].z; }
    }
    const Matrix<double, 3, 3>& matrixU() const { return u_; }
    const Matrix<double, 3, 3>& matrixV() const { return u_; }
private:
    Matrix<double, 3, 3> u_;
};

template <typename T>
class AngleAxis {
public:
    AngleAxis() : angle_(0), axis_(1, 0, 0) {}
    AngleAxis(T angle, const Matrix<T, 3, 1>& axis) : angle_(angle), axis_(axis) {}
    // Eigen goes matrix -> quaternion -> angle/axis; the angle is orc::rot_angle (same route), the axis is the
    // normalised antisymmetric part (only CalculateVelocity, off the registration path, reads it).
    AngleAxis(const Matrix<T, 3, 3>& m) {
        angle_ = orc::rot_angle(stub::to_orc(m));
        Matrix<T, 3, 1> v(m(2, 1) - m(1, 2), m(0, 2) - m(2, 0), m(1, 0) - m(0, 1));
        axis_ = v.squaredNorm() > T(0) ? v.normalized() : Matrix<T, 3, 1>(1, 0, 0);
    }
    T angle() const { return angle_; }
    const Matrix<T, 3, 1>& axis() const { return axis_; }
    Matrix<T, 3, 3> toRotationMatrix() const { return stub::from_orc(orc::angle_axis_to_rot(angle_, stub::to_orc(axis_))); }
    // AngleAxis * AngleAxis is a quaternion product in Eigen (defined after Quaternion below)
    friend Quaternion<T> operator*(const AngleAxis& a, const AngleAxis& b) { return Quaternion<T>(a) * Quaternion<T>(b); }
private:
    T angle_;
    Matrix<T, 3, 1> axis_;
};

// Eigen::Quaternion restated from its published formulas (coefficient order w, x, y, z in the constructor).
template <typename T>
class Quaternion {
public:
    Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
    Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
    Quaternion(const AngleAxis<T>& aa) {  // w = cos(a/2), vec = sin(a/2) * axis
        const T h = T(0.5) * aa.angle(), sn = std::sin(h);
        w_ = std::cos(h); x_ = sn * aa.axis()(0); y_ = sn * aa.axis()(1); z_ = sn * aa.axis()(2);
    }
    explicit Quaternion(const Matrix<T, 3, 3>& m) {  // Shepperd's method, as Eigen's quaternionbase_assign_impl
        T t = m(0, 0) + m(1, 1) + m(2, 2);
        if (t > T(0)) {
            t = std::sqrt(t + T(1));
            w_ = T(0.5) * t; t = T(0.5) / t;
            x_ = (m(2, 1) - m(1, 2)) * t; y_ = (m(0, 2) - m(2, 0)) * t; z_ = (m(1, 0) - m(0, 1)) * t;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + T(1));
            T v[3];
            v[i] = T(0.5) * t; t = T(0.5) / t;
            w_ = (m(k, j) - m(j, k)) * t;
            v[j] = (m(j, i) + m(i, j)) * t;
            v[k] = (m(k, i) + m(i, k)) * t;
            x_ = v[0]; y_ = v[1]; z_ = v[2];
        }
    }
    static Quaternion Identity() { return Quaternion(T(1), T(0), T(0), T(0)); }
    T& w() { return w_; } T& x() { return x_; } T& y() { return y_; } T& z() { return z_; }
    const T& w() const { return w_; } const T& x() const { return x_; } const T& y() const { return y_; } const T& z() const { return z_; }
    Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(x_, y_, z_); }
    T squaredNorm() const { return w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_; }
    T norm() const { return std::sqrt(squaredNorm()); }
    Quaternion normalized() const { const T n = norm(); return Quaternion(w_ / n, x_ / n, y_ / n, z_ / n); }
    void normalize() { *this = normalized(); }
    Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
    Quaternion inverse() const { const T n2 = squaredNorm(); return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2); }
    Quaternion operator*(const Quaternion& b) const {
        const Quaternion& a = *this;
        return Quaternion(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_, a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                          a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_, a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
    }
    Quaternion operator*(const AngleAxis<T>& b) const { return *this * Quaternion(b); }
    Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& v) const {  // _transformVector: v + w * uv + vec x uv, uv = 2 vec x v
        const Matrix<T, 3, 1> u = vec();
        Matrix<T, 3, 1> uv = u.cross(v);
        uv += uv;
        return v + w_ * uv + u.cross(uv);
    }
    Matrix<T, 3, 3> toRotationMatrix() const {
        const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
        const T twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
        Matrix<T, 3, 3> r;
        r(0, 0) = T(1) - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
        r(1, 0) = txy + twz; r(1, 1) = T(1) - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = T(1) - (txx + tyy);
        return r;
    }
    Quaternion slerp(T t, const Quaternion& o) const {  // Eigen's slerp
        const T one = T(1) - std::numeric_limits<T>::epsilon();
        const T d = w_ * o.w_ + x_ * o.x_ + y_ * o.y_ + z_ * o.z_, ad = std::fabs(d);
        T s0, s1;
        if (ad >= one) { s0 = T(1) - t; s1 = t; }
        else { const T th = std::acos(ad), st = std::sin(th); s0 = std::sin((T(1) - t) * th) / st; s1 = std::sin(t * th) / st; }
        if (d < T(0)) s1 = -s1;
        return Quaternion(s0 * w_ + s1 * o.w_, s0 * x_ + s1 * o.x_, s0 * y_ + s1 * o.y_, s0 * z_ + s1 * o.z_);
    }
private:
    T w_, x_, y_, z_;
};
template <typename T, int R, int C>
Matrix<T, R, C>::Matrix(const Quaternion<T>& q) { static_assert(R == 3 && C == 3, "rotation matrix"); *this = q.toRotationMatrix(); }

struct MatrixXd { static ScaledIdentityXd Identity(int rows, int cols) { (void)cols; return ScaledIdentityXd{rows, 1.0}; } };
inline ScaledIdentityXd operator*(const ScaledIdentityXd& a, double s) { return ScaledIdentityXd{a.n, a.s * s}; }

// Transform<T, 3, Affine> as far as the node and InterpolateTfWithTime use it: a 4 x 4 matrix whose last row stays (0 0 0 1).
template <typename T>
class AffineStub {
public:
    AffineStub() : m_(Matrix<T, 4, 4>::Identity()) {}
    static AffineStub Identity() { return AffineStub(); }
    T& operator()(int i, int j) { return m_(i, j); }
    const T& operator()(int i, int j) const { return m_(i, j); }
    const Matrix<T, 4, 4>& matrix() const { return m_; }
    BlockRef<T, 3, 1, 4, 4> translation() { return BlockRef<T, 3, 1, 4, 4>(m_, 0, 3); }
    Matrix<T, 3, 1> translation() const { return m_.template block<3, 1>(0, 3); }
    Matrix<T, 3, 3> linear() const { return m_.template block<3, 3>(0, 0); }
    Matrix<T, 3, 3> rotation() const { return linear(); }  // the node only stores rigid transforms
    AffineStub& translate(const Matrix<T, 3, 1>& v) { m_.template block<3, 1>(0, 3) = translation_value() + linear() * v; return *this; }
    AffineStub& rotate(const Quaternion<T>& q) { m_.template block<3, 3>(0, 0) = linear() * q.toRotationMatrix(); return *this; }
    AffineStub inverse() const {  // Affine mode: linear^-1 and -linear^-1 * translation
        AffineStub r;
        const Matrix<T, 3, 3> li = linear().inverse();
        r.m_.template block<3, 3>(0, 0) = li;
        r.m_.template block<3, 1>(0, 3) = -(li * translation_value());
        return r;
    }
    AffineStub operator*(const AffineStub& o) const {
        AffineStub r;
        r.m_.template block<3, 3>(0, 0) = linear() * o.linear();
        r.m_.template block<3, 1>(0, 3) = linear() * o.translation_value() + translation_value();
        return r;
    }
private:
    Matrix<T, 3, 1> translation_value() const { return m_.template block<3, 1>(0, 3); }
    Matrix<T, 4, 4> m_;
};

template <typename M> class Map;
template <typename T, int R, int C>
class Map<const Matrix<T, R, C>> : public Matrix<T, R, C> {
public:
    explicit Map(const T* p) { for (int i = 0; i < R * C; ++i) this->data()[i] = p[i]; }
};

using Matrix2d = Matrix<double, 2, 2>;
using Matrix3f = Matrix<float, 3, 3>;
using Quaterniond = Quaternion<double>;
using Quaternionf = Quaternion<float>;
using Affine3f = AffineStub<float>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;
using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector4d = Matrix<double, 4, 1>;
using Vector3i = Matrix<int, 3, 1>;
using Vector3f = Matrix<float, 3, 1>;
using AngleAxisd = AngleAxis<double>;

}  // namespace Eigen
