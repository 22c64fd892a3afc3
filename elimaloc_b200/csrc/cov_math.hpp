// Covariance arithmetic of the map build shared by the host builder (host_map.cpp) and the GPU builder (map_build.cu) — ONE
// source, compiled without FMA contraction on both sides (-ffp-contract=off / -fmad=false), so the two builders agree bit for bit.
//   sample covariance + U diag(1, 1, 1e-3) V^T regularisation   pcm_matching/include/voxel_hash_map.hpp:114-148, 222-250
#pragma once
#include <cmath>

#include "voxel_key.hpp"

namespace elm {

ELM_HD double cm_max(double a, double b) { return a > b ? a : b; }

// Jacobi rotations on a symmetric 3x3 held as a[6] = {xx, xy, xz, yy, yz, zz}; v = eigenvectors (columns).
ELM_HD void eig_sym3(const double a_in[6], double w[3], double v[3][3]) {
    double a[3][3] = {{a_in[0], a_in[1], a_in[2]}, {a_in[1], a_in[3], a_in[4]}, {a_in[2], a_in[4], a_in[5]}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
    const int pairs[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    for (int sweep = 0; sweep < 64; ++sweep) {
        if (a[0][1] == 0.0 && a[0][2] == 0.0 && a[1][2] == 0.0) break;
        for (int pi = 0; pi < 3; ++pi) {
            const int p = pairs[pi][0], q = pairs[pi][1];
            const double apq = a[p][q];
            if (apq == 0.0) continue;
            const double tau = (a[q][q] - a[p][p]) / (2.0 * apq);
            const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
            const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
            for (int k = 0; k < 3; ++k) {
                const double x = a[k][p], y = a[k][q];
                a[k][p] = c * x - s * y;
                a[k][q] = s * x + c * y;
            }
            for (int k = 0; k < 3; ++k) {
                const double x = a[p][k], y = a[q][k];
                a[p][k] = c * x - s * y;
                a[q][k] = s * x + c * y;
            }
            a[p][q] = a[q][p] = 0.0;
            for (int k = 0; k < 3; ++k) {
                const double x = v[k][p], y = v[k][q];
                v[k][p] = c * x - s * y;
                v[k][q] = s * x + c * y;
            }
        }
    }
    w[0] = a[0][0]; w[1] = a[1][1]; w[2] = a[2][2];
    // ascending, first-of-equals kept in place
    for (int i = 0; i < 2; ++i) {
        int k = i;
        for (int j = i + 1; j < 3; ++j) if (w[j] < w[k]) k = j;
        if (k != i) {
            { const double t_ = w[i]; w[i] = w[k]; w[k] = t_; }
            for (int r = 0; r < 3; ++r) { const double t_ = v[r][i]; v[r][i] = v[r][k]; v[r][k] = t_; }
        }
    }
}


// Symmetric 3x3 "plane regularisation" (voxel_hash_map.hpp:141-144, 241-244): I - (1 - 1e-3) n n^T with n the unit eigenvector of
// the smallest eigenvalue; degenerate smallest pair: convention documented in DESIGN.md.
ELM_HD void plane_regularize_hd(const double cov[9], double out[9], double normal[3]) {
    const double a[6] = {cov[0], 0.5 * (cov[1] + cov[3]), 0.5 * (cov[2] + cov[6]), cov[4], 0.5 * (cov[5] + cov[7]), cov[8]};
    double w[3], v[3][3];
    eig_sym3(a, w, v);
    const double lmax = cm_max(fabs(w[2]), fabs(w[0]));
    const double tol = 1e-9 * cm_max(lmax, 1e-300);
    double n[3];
    if (fabs(w[2] - w[0]) <= tol) {  // isotropic (or zero): Eigen's JacobiSVD leaves U = I
        n[0] = 0; n[1] = 0; n[2] = 1;
    } else if (fabs(w[1] - w[0]) <= tol) {  // rank-1-like: null space is a plane -> fixed completion
        const double u[3] = {v[0][2], v[1][2], v[2][2]};
        // axis least aligned with u; components within 1e-9 of each other count as tied (lowest axis wins), so that the choice
        // does not hinge on the last bits of u when the dominant direction is a lattice diagonal
        constexpr double kAxisTie = 1e-9;
        int k = 0;
        double best = fabs(u[0]);
        if (fabs(u[1]) < best - kAxisTie) { best = fabs(u[1]); k = 1; }
        if (fabs(u[2]) < best - kAxisTie) { best = fabs(u[2]); k = 2; }
        double e[3] = {0, 0, 0};
        e[k] = 1.0;
        const double d = u[k];
        double t[3] = {e[0] - d * u[0], e[1] - d * u[1], e[2] - d * u[2]};
        const double l = sqrt((t[0] * t[0] + t[1] * t[1]) + t[2] * t[2]);
        n[0] = t[0] / l; n[1] = t[1] / l; n[2] = t[2] / l;
    } else {
        n[0] = v[0][0]; n[1] = v[1][0]; n[2] = v[2][0];
    }
    const double k = 1.0 - 1e-3;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out[i * 3 + j] = ((i == j) ? 1.0 : 0.0) - k * n[i] * n[j];
    if (normal) { normal[0] = n[0]; normal[1] = n[1]; normal[2] = n[2]; }
}

// mean + sample covariance /(n - 1) of positions delivered by `next(i, p)` (i = 0 .. n - 1, called twice per i, in order), then the
// regularisation.  The summation order is the delivery order: both builders deliver the reference's order.
template <class Next>
ELM_HD void mean_cov_regularized_seq(size_t n, Next&& next, double mean[3], double cov_out[9], double normal[3]) {
    double s[3] = {0, 0, 0};
    for (size_t i = 0; i < n; ++i) { double p[3]; next(i, p); s[0] += p[0]; s[1] += p[1]; s[2] += p[2]; }
    const double dn = static_cast<double>(n);
    mean[0] = s[0] / dn; mean[1] = s[1] / dn; mean[2] = s[2] / dn;
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; ++i) {
        double p[3];
        next(i, p);
        const double d[3] = {p[0] - mean[0], p[1] - mean[1], p[2] - mean[2]};
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) c[a * 3 + b] += d[a] * d[b];
    }
    for (int i = 0; i < 9; ++i) c[i] /= (dn - 1.0);
    plane_regularize_hd(c, cov_out, normal);
}

}  // namespace elm
