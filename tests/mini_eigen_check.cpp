// Exercises the stand-in Eigen of oracle/ref_build/stubs (test infrastructure) on fixed inputs and prints the results, one
// labelled row of numbers per line; tests/test_mini_eigen.py recomputes every row with numpy.  Row-major printing.
#include <cstdio>
#include <Eigen/Dense>

template <typename M> void row(const char* name, const M& m, int r, int c) {
    std::printf("%s", name);
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) std::printf(" %.17g", static_cast<double>(m(i, j)));
    std::printf("\n");
}
void rowq(const char* name, const Eigen::Quaterniond& q) { std::printf("%s %.17g %.17g %.17g %.17g\n", name, q.w(), q.x(), q.y(), q.z()); }

int main() {
    using namespace Eigen;
    Matrix3d A;
    A << 2.0, -1.0, 0.5, 0.25, 3.0, -0.75, 1.5, 0.125, 4.0;
    Matrix4d T = Matrix4d::Identity();
    T.block<3, 3>(0, 0) = (AngleAxisd(0.3, Vector3d::UnitZ()) * AngleAxisd(-0.2, Vector3d::UnitY()) * AngleAxisd(0.1, Vector3d::UnitX())).toRotationMatrix();
    T.block<3, 1>(0, 3) << 1.5, -2.5, 0.75;
    Vector3d v(0.3, -1.2, 2.5), w(1.0, 0.5, -0.25);
    row("A", A, 3, 3);
    row("T", T, 4, 4);
    row("A_times_At", A * A.transpose(), 3, 3);
    row("A_inverse", A.inverse(), 3, 3);
    row("T_inverse", T.inverse(), 4, 4);
    row("T_times_Tinv", T * T.inverse(), 4, 4);
    row("scaled_sum", 2.5 * A + A * 0.5 - A / 4.0, 3, 3);
    row("A_v", A * v, 3, 1);
    row("vT_A", v.transpose() * A, 1, 3);
    row("cross", v.cross(w), 3, 1);
    std::printf("dot_norm %.17g %.17g %.17g %.17g\n", v.dot(w), v.norm(), v.squaredNorm(), A.trace());
    row("normalized", v.normalized(), 3, 1);
    row("diag_product", A * v.asDiagonal(), 3, 3);
    Matrix<double, 6, 6> S = Matrix<double, 6, 6>::Zero();
    S.block<3, 3>(0, 0) = A * A.transpose();
    S.block<3, 3>(3, 3) = A.transpose() * A + Matrix3d::Identity();
    S.block<3, 3>(0, 3) = 0.1 * A;
    S.block<3, 3>(3, 0) = 0.1 * A.transpose();
    Matrix<double, 6, 1> b;
    b << 1.0, -2.0, 3.0, 0.5, 0.25, -1.5;
    row("S", S, 6, 6);
    row("S_inverse", S.inverse(), 6, 6);
    row("ldlt_solve", S.ldlt().solve(b), 6, 1);
    row("tail_head", b.tail<3>() - b.head<3>(), 3, 1);
    Matrix<double, 3, Dynamic> N(3, 4);
    N.col(0) = v; N.col(1) = w; N.col(2) = v + w; N.col(3) = 2.0 * w - v;
    Vector3d mean = N.rowwise().mean();
    N.colwise() -= mean;
    row("mean", mean, 3, 1);
    row("sample_cov", (N * N.transpose()) / 3, 3, 3);
    SelfAdjointEigenSolver<Matrix3d> es(A * A.transpose());
    row("eigenvalues", es.eigenvalues(), 3, 1);
    row("eig_residual", (A * A.transpose()) * es.eigenvectors() - es.eigenvectors() * Matrix3d(es.eigenvalues().asDiagonal()), 3, 3);
    JacobiSVD<Matrix3d> svd(A * A.transpose(), ComputeFullU | ComputeFullV);
    row("plane", svd.matrixU() * Vector3d(1, 1, 1e-3).asDiagonal() * svd.matrixV().transpose(), 3, 3);
    Quaterniond q1(AngleAxisd(0.7, Vector3d(0.0, 0.6, 0.8))), q2(T.block<3, 3>(0, 0));
    rowq("q_from_angle_axis", q1);
    rowq("q_from_matrix", q2);
    rowq("q_product_normalized", (q1 * q2).normalized());
    rowq("q_inverse", q2.inverse());
    row("q_rotate", q1 * v, 3, 1);
    row("q_to_matrix", q1.toRotationMatrix(), 3, 3);
    AngleAxisd aa(T.block<3, 3>(0, 0));
    std::printf("angle_of_matrix %.17g\n", aa.angle());
    Quaternionf qf(0.9f, 0.1f, -0.3f, 0.2f);
    qf.normalize();
    Quaternionf sl = Quaternionf::Identity().slerp(0.3f, qf);
    std::printf("slerp %.9g %.9g %.9g %.9g\n", sl.w(), sl.x(), sl.y(), sl.z());
    Affine3f a1 = Affine3f::Identity(), a2 = Affine3f::Identity();
    a1.translation() << 1.0f, 2.0f, 3.0f;
    a1.rotate(qf);
    a2.translate(Vector3f(0.5f, -0.5f, 0.25f));
    a2.rotate(Quaternionf(AngleAxis<float>(0.4f, Vector3f::UnitZ())));
    row("affine_between", (a1.inverse() * a2).matrix(), 4, 4);
    Matrix<int, 3, 1> k = (Vector3d(-0.5, 1.7, -2.2) / 1.0).cast<int>();
    std::printf("cast_int %d %d %d\n", k.x(), k.y(), k.z());
    return 0;
}
