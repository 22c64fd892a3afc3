// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.hpp header).
// Parity status: PINNED on oracle/_ref — the reference's own sources compiled against stand-in Eigen/oneTBB headers
// (tests/test_reference_build.py); Eigen's arithmetic kernels themselves stay restated (smallmat.hpp).
// Restates /root/reference/src/app/localization/pcm_matching/src/registration.cpp; every function
// cites the lines it follows.  The radar-covariance branch (use_radar_cov, reg.cpp:109-111,188-190,
// 302-305; quirk Q14) is out of scope and not restated.
#include "registration.hpp"

#include <cstdio>

namespace orc {

namespace {
inline double square(double x) { return x * x; }  // reg.hpp:219

// J_g = [ I(3x3) , -skew(local) ]   (reg.cpp:36-41, reg.hpp:221-225)
inline void jacobian(const V3& s, double J[3][6]) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 6; ++c) J[r][c] = 0.0;
    J[0][0] = J[1][1] = J[2][2] = 1.0;
    // skew(s) = [0 -z y; z 0 -x; -y x 0]; block = -1.0 * skew
    J[0][3] = -1.0 * 0.0;   J[0][4] = -1.0 * -s.z;  J[0][5] = -1.0 * s.y;
    J[1][3] = -1.0 * s.z;   J[1][4] = -1.0 * 0.0;   J[1][5] = -1.0 * -s.x;
    J[2][3] = -1.0 * -s.y;  J[2][4] = -1.0 * s.x;   J[2][5] = -1.0 * 0.0;
}

// JTJ += w * J^T * M * J ;  JTr += w * J^T * M * r
inline void accumulate(const double J[3][6], const M3& M, const V3& r, double w, M6& JTJ, V6& JTr) {
    double MJ[3][6];
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 6; ++c) MJ[i][c] = (M(i, 0) * J[0][c] + M(i, 1) * J[1][c]) + M(i, 2) * J[2][c];
    const V3 Mr = mul(M, r);
    const double mr[3] = {Mr.x, Mr.y, Mr.z};
    for (int a = 0; a < 6; ++a) {
        for (int b = 0; b < 6; ++b)
            JTJ(a, b) += w * ((J[0][a] * MJ[0][b] + J[1][a] * MJ[1][b]) + J[2][a] * MJ[2][b]);
        JTr.v[a] += w * ((J[0][a] * mr[0] + J[1][a] * mr[1]) + J[2][a] * mr[2]);
    }
}

// The accumulation loop of the three AlignClouds*.  threads <= 1: one serial loop, as in the reference.  Otherwise (the
// best-effort timing variant, not in the reference): static chunks, per-chunk sums joined in chunk order.
struct PairSums { M6 JTJ; V6 JTr; double res = 0.0; };
template <class Body>
inline PairSums accumulate_pairs(size_t n, int threads, Body body) {
    if (threads <= 1) {
        PairSums s;
        for (size_t i = 0; i < n; ++i) body(i, s);
        return s;
    }
    std::vector<PairSums> parts(static_cast<size_t>(threads));
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; ++t) {
        PairSums& s = parts[static_cast<size_t>(t)];
        for (size_t i = n * static_cast<size_t>(t) / threads; i < n * static_cast<size_t>(t + 1) / threads; ++i) body(i, s);
    }
    PairSums tot = parts[0];
    for (int t = 1; t < threads; ++t) {
        for (int a = 0; a < 6; ++a) { for (int b = 0; b < 6; ++b) tot.JTJ(a, b) += parts[t].JTJ(a, b); tot.JTr.v[a] += parts[t].JTr.v[a]; }
        tot.res += parts[t].res;
    }
    return tot;
}

// Tail shared by the three AlignClouds*: LM damping on diag(JTJ), LDLT solve, AngleAxis -> 4x4
// (reg.cpp:55-65, 136-151, 213-224).  x = [translation ; rotation vector]  (Q9, Q10).
inline M4 solve_update(const M6& JTJ, const V6& JTr, double lm_lambda, M6* regularized_out) {
    M6 A = JTJ;
    for (int i = 0; i < 6; ++i) A(i, i) = JTJ(i, i) + lm_lambda * JTJ(i, i);
    if (regularized_out) *regularized_out = A;
    const V6 x = ldlt_solve(A, JTr);
    const V3 rotv(x.v[3], x.v[4], x.v[5]);
    const M3 R = angle_axis_to_rot(norm(rotv), normalized(rotv));
    M4 T = M4::Identity();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T(i, j) = R(i, j);
    T(0, 3) = x.v[0]; T(1, 3) = x.v[1]; T(2, 3) = x.v[2];
    return T;
}
}  // namespace

// threads of TransformPoints: 1 (serial, as in the reference) unless the best-effort timing variant is on
static int g_transform_threads = 1;

// reg.hpp:136-148 — pose <- T * pose, every other field (incl. `local`) copied.
void Registration::TransformPoints(const M4& T, const std::vector<PointStruct>& points,
                                   std::vector<PointStruct>& o_points) {
    o_points.resize(points.size());
    const long long n = static_cast<long long>(points.size());
#pragma omp parallel for num_threads(g_transform_threads) schedule(static) if (g_transform_threads > 1)
    for (long long i = 0; i < n; ++i) {
        PointStruct p = points[i];
        p.pose = apply(T, points[i].pose);
        o_points[i] = p;
    }
}

// reg.cpp:15-66 (P2P).  Serial, as in the reference.
M4 Registration::AlignCloudsLocal(const std::vector<PointStruct>& source_global,
                                  const std::vector<PointStruct>& target_global, const M4& last_icp_pose,
                                  double trans_th, const RegistrationConfig& cfg, Linearization* lin) {
    const M4 last_icp_pose_inv = inverse(last_icp_pose);  // reg.cpp:24
    const M3 I3 = M3::Identity();
    const PairSums sums = accumulate_pairs(source_global.size(), cfg.parallel_accumulate ? cfg.i_max_thread : 1, [&](size_t i, PairSums& s) {
        const V3 target_local = apply(last_icp_pose_inv, target_global[i].pose);  // reg.cpp:29-33
        const V3 residual_local = target_local - source_global[i].local;          // reg.cpp:34
        double J[3][6];
        jacobian(source_global[i].local, J);
        const double weight_g = square(trans_th) / square(trans_th + sqnorm(residual_local));  // reg.cpp:44
        accumulate(J, I3, residual_local, weight_g, s.JTJ, s.JTr);                             // reg.cpp:47-48
        s.res += norm(residual_local);                                                         // reg.cpp:50
    });
    const M6& JTJ = sums.JTJ;
    const V6& JTr = sums.JTr;
    const double d_residual_sum = sums.res;
    d_fitness_score_ = d_residual_sum / source_global.size();  // reg.cpp:53
    if (lin) { lin->JTJ = JTJ; lin->JTr = JTr; lin->residual_sum = d_residual_sum; lin->n_corr = (long long)source_global.size(); }
    return solve_update(JTJ, JTr, cfg.lm_lambda, nullptr);
}

// reg.cpp:68-152 (GICP).  Target position is covariance.mean (Q4); fitness term is point-to-plane (Q12).
M4 Registration::AlignCloudsLocalPointCov(const std::vector<PointStruct>& source_global,
                                          const std::vector<PointStruct>& target_global, M6& local_cov,
                                          const M4& last_icp_pose, double trans_th, const RegistrationConfig& cfg,
                                          Linearization* lin) {
    const M3 sensor_rot = rot_of(last_icp_pose);
    const M3 sensor_rot_inv = inverse(sensor_rot);         // reg.cpp:79
    const M3 sensor_rot_inv_t = transpose(sensor_rot_inv);
    const M4 last_icp_pose_inv = inverse(last_icp_pose);   // reg.cpp:81
    const PairSums sums = accumulate_pairs(source_global.size(), cfg.parallel_accumulate ? cfg.i_max_thread : 1, [&](size_t i, PairSums& s) {
        const CovStruct& target_cov = target_global[i].covariance;
        double w[3];
        M3 V;
        sym_eig3(target_cov.cov, w, V);                     // reg.cpp:89
        const V3 vec_normal_global(V(0, 0), V(1, 0), V(2, 0));  // col(0): smallest eigenvalue (reg.cpp:91)
        const V3 vec_normal_local = normalized(mul(sensor_rot_inv, vec_normal_global));  // reg.cpp:94-95
        const V3 target_local = apply(last_icp_pose_inv, target_cov.mean);               // reg.cpp:97-100
        const V3 residual_local = target_local - source_global[i].local;                 // reg.cpp:101
        const M3 RCR = mul(mul(sensor_rot_inv, target_cov.cov), sensor_rot_inv_t);       // reg.cpp:107
        const M3 mahalanobis_local = inverse(RCR);                                       // reg.cpp:113
        double J[3][6];
        jacobian(source_global[i].local, J);
        const double weight_g = square(trans_th) / square(trans_th + sqnorm(residual_local)) * 0.8 + 0.2;  // reg.cpp:121
        accumulate(J, mahalanobis_local, residual_local, weight_g, s.JTJ, s.JTr);                        // reg.cpp:124-125
        s.res += std::fabs(dot(residual_local, vec_normal_local));                                       // reg.cpp:128-131
    });
    const M6& JTJ = sums.JTJ;
    const V6& JTr = sums.JTr;
    const double d_residual_sum = sums.res;
    d_fitness_score_ = d_residual_sum / source_global.size();  // reg.cpp:134
    if (lin) { lin->JTJ = JTJ; lin->JTr = JTr; lin->residual_sum = d_residual_sum; lin->n_corr = (long long)source_global.size(); }
    M6 regularized;
    const M4 T = solve_update(JTJ, JTr, cfg.lm_lambda, &regularized);
    local_cov = inverse(regularized);  // reg.cpp:141-142 (Q11)
    return T;
}

// reg.cpp:154-225 (VGICP and AVGICP).  weight < 0.01 skips both sums but not the denominator (Q7).
M4 Registration::AlignCloudsLocalVoxelCov(const std::vector<PointStruct>& source_global,
                                          const std::vector<CovStruct>& target_cov_global, const M4& last_icp_pose,
                                          double trans_th, const RegistrationConfig& cfg, Linearization* lin) {
    const M3 sensor_rot = rot_of(last_icp_pose);
    const M3 sensor_rot_inv = inverse(sensor_rot);        // reg.cpp:165
    const M3 sensor_rot_inv_t = transpose(sensor_rot_inv);
    const M4 last_icp_pose_inv = inverse(last_icp_pose);  // reg.cpp:167
    const PairSums sums = accumulate_pairs(source_global.size(), cfg.parallel_accumulate ? cfg.i_max_thread : 1, [&](size_t i, PairSums& s) {
        const CovStruct& target_cov = target_cov_global[i];
        const V3 target_local = apply(last_icp_pose_inv, target_cov.mean);          // reg.cpp:176-180
        const V3 residual_local = target_local - source_global[i].local;            // reg.cpp:181
        const M3 RCR = mul(mul(sensor_rot_inv, target_cov.cov), sensor_rot_inv_t);  // reg.cpp:187
        const M3 mahalanobis_local = inverse(RCR);                                  // reg.cpp:191
        double J[3][6];
        jacobian(source_global[i].local, J);
        const double weight_g = square(trans_th) / square(trans_th + sqnorm(residual_local));  // reg.cpp:199
        if (weight_g < 0.01) return;                                                           // reg.cpp:201 (continue)
        accumulate(J, mahalanobis_local, residual_local, weight_g, s.JTJ, s.JTr);              // reg.cpp:204-205
        s.res += norm(residual_local);                                                         // reg.cpp:207
    });
    const M6& JTJ = sums.JTJ;
    const V6& JTr = sums.JTr;
    const double d_residual_sum = sums.res;
    d_fitness_score_ = d_residual_sum / source_global.size();  // reg.cpp:210
    if (lin) { lin->JTJ = JTJ; lin->JTr = JTr; lin->residual_sum = d_residual_sum; lin->n_corr = (long long)source_global.size(); }
    return solve_update(JTJ, JTr, cfg.lm_lambda, nullptr);
}

Linearization Registration::LinearizeOnce(const std::vector<PointStruct>& source_local, const VoxelHashMap& voxel_map,
                                          const M4& pose, const RegistrationConfig& cfg) {
    std::vector<PointStruct> source_global, sc, tc;
    std::vector<CovStruct> tcc;
    g_transform_threads = cfg.parallel_accumulate ? cfg.i_max_thread : 1;
    TransformPoints(pose, source_local, source_global);
    Linearization lin;
    M6 cov_unused;
    const double saved = d_fitness_score_;
    switch (cfg.icp_method) {
        case P2P:
            std::tie(sc, tc) = voxel_map.GetCorrespondencePoints(source_global, cfg.max_search_dist, cfg.i_max_thread);
            AlignCloudsLocal(sc, tc, pose, cfg.max_search_dist, cfg, &lin);
            break;
        case GICP:
            std::tie(sc, tc) = voxel_map.GetCorrespondencePoints(source_global, cfg.max_search_dist, cfg.i_max_thread);
            AlignCloudsLocalPointCov(sc, tc, cov_unused, pose, cfg.max_search_dist, cfg, &lin);
            break;
        case VGICP:
            std::tie(sc, tcc) = voxel_map.GetCorrespondencesCov(source_global, cfg.max_search_dist, cfg.i_max_thread);
            AlignCloudsLocalVoxelCov(sc, tcc, pose, cfg.max_search_dist, cfg, &lin);
            break;
        default:
            std::tie(sc, tcc) = voxel_map.GetCorrespondencesAllCov(source_global, cfg.max_search_dist, cfg.i_max_thread);
            AlignCloudsLocalVoxelCov(sc, tcc, pose, cfg.max_search_dist, cfg, &lin);
            break;
    }
    d_fitness_score_ = saved;
    return lin;
}

// reg.cpp:274-418
M4 Registration::RunRegister(const std::vector<PointStruct>& source_local, const VoxelHashMap& voxel_map,
                             const M4& initial_guess, const RegistrationConfig& cfg, bool& is_success,
                             double& fitness_score, M6& local_cov, std::vector<IterTrace>* trace) {
    std::vector<PointStruct> source_c_global, target_c_global;
    std::vector<CovStruct> target_cov_c_global;
    local_cov = M6::Identity();  // reg.cpp:280

    const int i_source_total_num = static_cast<int>(source_local.size());
    int i_source_corr_num = 0;
    double corres_ratio = 0.0;

    std::vector<PointStruct> source_global;
    g_transform_threads = cfg.parallel_accumulate ? cfg.i_max_thread : 1;
    TransformPoints(initial_guess, source_local, source_global);  // reg.cpp:289

    if (voxel_map.Empty()) {  // reg.cpp:291-295
        is_success = false;
        return initial_guess;
    }

    M4 last_icp_pose = initial_guess;
    M4 estimation_local = M4::Identity();

    for (int j = 0; j < cfg.max_iteration; ++j) {  // reg.cpp:310
        switch (cfg.icp_method) {                  // reg.cpp:318-335
            case P2P:
            case GICP:
                std::tie(source_c_global, target_c_global) =
                    voxel_map.GetCorrespondencePoints(source_global, cfg.max_search_dist, cfg.i_max_thread);
                break;
            case VGICP:
                std::tie(source_c_global, target_cov_c_global) =
                    voxel_map.GetCorrespondencesCov(source_global, cfg.max_search_dist, cfg.i_max_thread);
                break;
            default:
                std::tie(source_c_global, target_cov_c_global) =
                    voxel_map.GetCorrespondencesAllCov(source_global, cfg.max_search_dist, cfg.i_max_thread);
                break;
        }
        i_source_corr_num = static_cast<int>(source_c_global.size());  // reg.cpp:349

        corres_ratio = (float)i_source_corr_num / i_source_total_num;  // reg.cpp:351 — float division (Q6)
        if (corres_ratio < cfg.min_overlap_ratio) {                    // reg.cpp:352-356
            is_success = false;
            return last_icp_pose;
        }

        IterTrace tr;
        tr.pose_in = last_icp_pose;
        switch (cfg.icp_method) {  // reg.cpp:358-375
            case P2P:
                estimation_local = AlignCloudsLocal(source_c_global, target_c_global, last_icp_pose,
                                                    cfg.max_search_dist, cfg, &tr.lin);
                break;
            case GICP:
                estimation_local = AlignCloudsLocalPointCov(source_c_global, target_c_global, local_cov, last_icp_pose,
                                                            cfg.max_search_dist, cfg, &tr.lin);
                break;
            default:
                estimation_local = AlignCloudsLocalVoxelCov(source_c_global, target_cov_c_global, last_icp_pose,
                                                            cfg.max_search_dist, cfg, &tr.lin);
                break;
        }

        last_icp_pose = mul(last_icp_pose, estimation_local);  // reg.cpp:378 — right multiplication (Q10)
        tr.pose_out = last_icp_pose;
        if (trace) trace->push_back(tr);

        const double rot_norm = rot_angle(rot_of(estimation_local));  // reg.cpp:381-382
        const V3 dt(estimation_local(0, 3), estimation_local(1, 3), estimation_local(2, 3));
        const double transform_norm = rot_norm + norm(dt);            // reg.cpp:384
        if (transform_norm < cfg.icp_termination_threshold_m) break;  // reg.cpp:385-387 (Q13)

        TransformPoints(last_icp_pose, source_local, source_global);  // reg.cpp:390
    }

    if (d_fitness_score_ > cfg.max_fitness_score) {  // reg.cpp:405-409
        is_success = false;
        return last_icp_pose;
    }
    fitness_score = d_fitness_score_;  // reg.cpp:415 — written only on success (Q11)
    is_success = true;
    return last_icp_pose;
}

}  // namespace orc
