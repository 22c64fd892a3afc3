// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// C entry points over the REFERENCE's own VoxelHashMap / Registration (pcm_matching/include/*.hpp, src/*.cpp compiled
// unmodified from /root/reference against the stand-in headers in stubs/), with the same signatures and dump layout as the
// oracle's orc_* functions (oracle/capi.cpp) so that tests/test_reference_build.py can drive both through one wrapper.
// Built by oracle/Makefile into oracle/_ref/libref.so (git-ignored; travels to the GPU box as a built file).
#include <algorithm>
#include <cstdint>
#include <chrono>
#include <cstring>

#include "registration.hpp"  // the reference's header (found through -I <reference>/pcm_matching/include)

namespace {

struct ref_reg_config {  // field order of orc_reg_config / elm_reg_config
    int32_t icp_method, max_iteration, max_thread, use_radar_cov, debug_print, reserved0;
    double max_search_dist, lm_lambda, icp_termination_threshold_m, min_overlap_ratio, max_fitness_score;
    double range_variance_m, azimuth_variance_deg, elevation_variance_deg;
};

RegistrationConfig to_cfg(const ref_reg_config* c) {
    RegistrationConfig r{};
    r.i_max_thread = c->max_thread > 0 ? c->max_thread : 1;
    r.icp_method = static_cast<IcpMethod>(c->icp_method);
    r.voxel_search_method = 2;
    r.gicp_cov_search_dist = 0.4;
    r.use_radar_cov = c->use_radar_cov != 0;
    r.max_iteration = c->max_iteration;
    r.max_search_dist = c->max_search_dist;
    r.lm_lambda = c->lm_lambda;
    r.icp_termination_threshold_m = c->icp_termination_threshold_m;
    r.min_overlap_ratio = c->min_overlap_ratio;
    r.max_fitness_score = c->max_fitness_score;
    r.doppler_trans_lambda = 0.0;
    r.range_variance_m = c->range_variance_m;
    r.azimuth_variance_deg = c->azimuth_variance_deg;
    r.elevation_variance_deg = c->elevation_variance_deg;
    r.b_debug_print = c->debug_print != 0;
    return r;
}

// What the node does with a pcl::PointXYZINormal (Pcl2PointStruct, pcm_matching.hpp:205-220): float fields widened to
// double, local = pose.
std::vector<PointStruct> to_points(const float* xyz, size_t n) {
    std::vector<PointStruct> v(n);
    for (size_t i = 0; i < n; ++i) {
        v[i].pose = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        v[i].local = v[i].pose;
    }
    return v;
}
Eigen::Matrix4d to_m4(const double* t) {  // row-major in
    Eigen::Matrix4d m;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m(i, j) = t[4 * i + j];
    return m;
}
template <int N>
void from_mat(const Eigen::Matrix<double, N, N>& m, double* out) {  // row-major out
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) out[N * i + j] = m(i, j);
}

}  // namespace

extern "C" {

// threads of the oneTBB stand-in (1 = inline).  Results do not depend on it: chunks are joined in order.
void ref_set_threads(int n) { tbb::stub_threads() = n > 0 ? n : 1; }
int ref_get_threads() { return tbb::stub_threads(); }

void* ref_map_create(double voxel_size, int max_pts) {
    auto* m = new VoxelHashMap();
    m->Init(voxel_size, max_pts);
    return m;
}
void ref_map_destroy(void* m) { delete static_cast<VoxelHashMap*>(m); }
void ref_map_add_points(void* m, const float* xyz, size_t n) { static_cast<VoxelHashMap*>(m)->AddPoints(to_points(xyz, n)); }
void ref_map_cal_voxel_cov(void* m) { static_cast<VoxelHashMap*>(m)->CalVoxelCovAll(); }
void ref_map_cal_point_cov(void* m, double d) { static_cast<VoxelHashMap*>(m)->CalPointCovAll(d); }
size_t ref_map_num_voxels(void* m) { return static_cast<VoxelHashMap*>(m)->map_.size(); }
size_t ref_map_num_points(void* m) { return static_cast<VoxelHashMap*>(m)->Pointcloud().size(); }
int ref_map_empty(void* m) { return static_cast<VoxelHashMap*>(m)->Empty() ? 1 : 0; }

// Canonical dump (voxels sorted by key x, y, z; points in insertion order inside a voxel); any pointer may be NULL.
void ref_map_export(void* mp, int32_t* keys, int32_t* counts, double* vmean, double* vcov, float* pxyz, double* pmean, double* pcov) {
    auto* m = static_cast<VoxelHashMap*>(mp);
    using Entry = std::pair<const VoxelHashMap::Voxel, VoxelHashMap::VoxelBlock>;
    std::vector<const Entry*> v;
    v.reserve(m->map_.size());
    for (const auto& kv : m->map_) v.push_back(&kv);
    std::sort(v.begin(), v.end(), [](const Entry* a, const Entry* b) {
        for (int k = 0; k < 3; ++k) if (a->first(k) != b->first(k)) return a->first(k) < b->first(k);
        return false;
    });
    size_t p = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        const auto& vb = v[i]->second;
        if (keys) for (int k = 0; k < 3; ++k) keys[3 * i + k] = v[i]->first(k);
        if (counts) counts[i] = static_cast<int32_t>(vb.points.size());
        if (vmean) for (int k = 0; k < 3; ++k) vmean[3 * i + k] = vb.covariance.mean(k);
        if (vcov) from_mat<3>(vb.covariance.cov, vcov + 9 * i);
        for (const auto& pt : vb.points) {
            if (pxyz) for (int k = 0; k < 3; ++k) pxyz[3 * p + k] = static_cast<float>(pt.pose(k));
            if (pmean) for (int k = 0; k < 3; ++k) pmean[3 * p + k] = pt.covariance.mean(k);
            if (pcov) from_mat<3>(pt.covariance.cov, pcov + 9 * p);
            ++p;
        }
    }
}

void* ref_reg_create() { return new Registration(); }
void ref_reg_destroy(void* r) { delete static_cast<Registration*>(r); }

// Registration::RunRegister (registration.cpp:274-418).  The reference keeps its normal equations in locals; the ldlt tap of
// the stand-in records them: tr_A[i] = JTJ + lm_lambda * diag(JTJ) and tr_b[i] = JTr of iteration i (up to max_trace).
void ref_run_register(void* reg, void* map, const float* src, size_t n, const double* T_init, const ref_reg_config* c, double* T_out,
                      int32_t* is_success, double* fitness, double* local_cov, int32_t max_trace, int32_t* n_iter, double* tr_A, double* tr_b) {
    auto* R = static_cast<Registration*>(reg);
    auto* M = static_cast<VoxelHashMap*>(map);
    const std::vector<PointStruct> pts = to_points(src, n);
    bool ok = (*is_success != 0);
    Eigen::Matrix6d cov;
    Eigen::LdltTap& tap = Eigen::LdltTap::get();
    tap.A.clear();
    tap.b.clear();
    tap.enabled = true;
    const Eigen::Matrix4d T = R->RunRegister(pts, *M, to_m4(T_init), to_cfg(c), ok, *fitness, cov);
    tap.enabled = false;
    from_mat<4>(T, T_out);
    *is_success = ok ? 1 : 0;
    from_mat<6>(cov, local_cov);
    if (n_iter) *n_iter = static_cast<int32_t>(tap.A.size());
    for (int i = 0; i < static_cast<int>(tap.A.size()) && i < max_trace; ++i) {
        if (tr_A) std::memcpy(tr_A + 36 * i, tap.A[i].data(), 36 * sizeof(double));
        if (tr_b) std::memcpy(tr_b + 6 * i, tap.b[i].data(), 6 * sizeof(double));
    }
}
// Wall seconds of one RunRegister call (the reference arm of bench.py); iterations counted through the ldlt tap.
double ref_time_register(void* reg, void* map, const float* src, size_t n, const double* T_init, const ref_reg_config* c, int32_t* iters_done) {
    auto* R = static_cast<Registration*>(reg);
    auto* M = static_cast<VoxelHashMap*>(map);
    const std::vector<PointStruct> pts = to_points(src, n);
    bool ok = false;
    double fit = 0.0;
    Eigen::Matrix6d cov;
    Eigen::LdltTap& tap = Eigen::LdltTap::get();
    tap.A.clear();
    tap.b.clear();
    tap.enabled = true;
    const auto t0 = std::chrono::steady_clock::now();
    R->RunRegister(pts, *M, to_m4(T_init), to_cfg(c), ok, fit, cov);
    const auto t1 = std::chrono::steady_clock::now();
    tap.enabled = false;
    if (iters_done) *iters_done = static_cast<int32_t>(tap.A.size());
    return std::chrono::duration<double>(t1 - t0).count();
}
double ref_reg_fitness(void* reg) { return static_cast<Registration*>(reg)->d_fitness_score_; }

// One search + one AlignClouds* call at a fixed pose with lm_lambda = 0, so the tapped system IS (JTJ, JTr).
// residual_sum = d_fitness_score_ * number of pairs (registration.cpp:53, 134, 211).
void ref_linearize(void* reg, void* map, const float* src, size_t n, const double* Tp, const ref_reg_config* c, double* JTJ, double* JTr,
                   double* res_sum, long long* n_corr) {
    auto* R = static_cast<Registration*>(reg);
    auto* M = static_cast<VoxelHashMap*>(map);
    RegistrationConfig cfg = to_cfg(c);
    cfg.lm_lambda = 0.0;
    Eigen::Matrix4d T = to_m4(Tp);
    std::vector<PointStruct> local = to_points(src, n), global;
    R->TransformPoints(T, local, global);
    std::vector<PointStruct> sc, tc;
    std::vector<CovStruct> tcc;
    Eigen::Matrix6d cov;
    Eigen::LdltTap& tap = Eigen::LdltTap::get();
    tap.A.clear();
    tap.b.clear();
    tap.enabled = true;
    switch (cfg.icp_method) {
        case P2P:
            std::tie(sc, tc) = M->GetCorrespondencePoints(global, cfg.max_search_dist);
            if (!sc.empty()) R->AlignCloudsLocal(sc, tc, T, cfg.max_search_dist, cfg);
            break;
        case GICP:
            std::tie(sc, tc) = M->GetCorrespondencePoints(global, cfg.max_search_dist);
            if (!sc.empty()) R->AlignCloudsLocalPointCov(sc, tc, cov, T, cfg.max_search_dist, cfg);
            break;
        case VGICP:
            std::tie(sc, tcc) = M->GetCorrespondencesCov(global, cfg.max_search_dist);
            if (!sc.empty()) R->AlignCloudsLocalVoxelCov(sc, tcc, T, cfg.max_search_dist, cfg);
            break;
        case AVGICP:
            std::tie(sc, tcc) = M->GetCorrespondencesAllCov(global, cfg.max_search_dist);
            if (!sc.empty()) R->AlignCloudsLocalVoxelCov(sc, tcc, T, cfg.max_search_dist, cfg);
            break;
    }
    tap.enabled = false;
    std::memset(JTJ, 0, 36 * sizeof(double));
    std::memset(JTr, 0, 6 * sizeof(double));
    *res_sum = 0.0;
    *n_corr = static_cast<long long>(sc.size());
    if (!tap.A.empty()) {
        std::memcpy(JTJ, tap.A[0].data(), 36 * sizeof(double));
        std::memcpy(JTr, tap.b[0].data(), 6 * sizeof(double));
        *res_sum = R->d_fitness_score_ * static_cast<double>(sc.size());
    }
}

// Correspondence dump, same layout as orc_correspondences: K = 1 (P2P / GICP / VGICP) or 7 (AVGICP) targets per scan point.
void ref_correspondences(void* map, const float* src, size_t n, const double* Tp, int method, double max_dist, int32_t* count, double* target) {
    auto* M = static_cast<VoxelHashMap*>(map);
    Registration R;
    const Eigen::Matrix4d T = to_m4(Tp);
    const int K = (method == AVGICP) ? 7 : 1;
    std::vector<PointStruct> one, moved, sc, tc;
    std::vector<CovStruct> tcc;
    for (size_t i = 0; i < n; ++i) {
        one = to_points(src + 3 * i, 1);
        R.TransformPoints(T, one, moved);
        for (int j = 0; j < K * 3; ++j) target[i * K * 3 + j] = 0.0;
        if (method == P2P || method == GICP) {
            std::tie(sc, tc) = M->GetCorrespondencePoints(moved, max_dist);
            count[i] = static_cast<int32_t>(sc.size());
            if (!tc.empty()) for (int k = 0; k < 3; ++k) target[i * 3 + k] = tc[0].pose(k);
        } else {
            if (method == VGICP) std::tie(sc, tcc) = M->GetCorrespondencesCov(moved, max_dist);
            else std::tie(sc, tcc) = M->GetCorrespondencesAllCov(moved, max_dist);
            count[i] = static_cast<int32_t>(sc.size());
            for (size_t j = 0; j < tcc.size(); ++j) for (int k = 0; k < 3; ++k) target[(i * K + j) * 3 + k] = tcc[j].mean(k);
        }
    }
}

// The whole-scan search in one call (the order of the emitted pairs is part of what is compared): for every emitted pair
// the index of its scan point is recovered from the `intensity` field, which the search copies through untouched.
size_t ref_search_pairs(void* map, const float* src, size_t n, const double* Tp, int method, double max_dist, int32_t* src_index, double* target,
                        size_t capacity) {
    auto* M = static_cast<VoxelHashMap*>(map);
    Registration R;
    std::vector<PointStruct> local = to_points(src, n), global, sc, tc;
    for (size_t i = 0; i < n; ++i) local[i].intensity = static_cast<double>(i);
    R.TransformPoints(to_m4(Tp), local, global);
    std::vector<CovStruct> tcc;
    if (method == P2P || method == GICP) std::tie(sc, tc) = M->GetCorrespondencePoints(global, max_dist);
    else if (method == VGICP) std::tie(sc, tcc) = M->GetCorrespondencesCov(global, max_dist);
    else std::tie(sc, tcc) = M->GetCorrespondencesAllCov(global, max_dist);
    for (size_t i = 0; i < sc.size() && i < capacity; ++i) {
        src_index[i] = static_cast<int32_t>(sc[i].intensity);
        for (int k = 0; k < 3; ++k) target[3 * i + k] = (method == P2P || method == GICP) ? tc[i].pose(k) : tcc[i].mean(k);
    }
    return sc.size();
}

int ref_find_ground_height(void* mp, double x, double y, double* ground_z) {
    return static_cast<VoxelHashMap*>(mp)->FindGroundHeight(Eigen::Vector2d(x, y), *ground_z) ? 1 : 0;
}

// VoxelHashMap::VoxelDownsample (voxel_hash_map.hpp:260-283): survivors' input indices (recovered through `intensity`),
// in the reference's own output order (unordered_map iteration order).
size_t ref_voxel_downsample(const float* xyz, size_t n, double voxel_size, int32_t* index_out) {
    VoxelHashMap m;
    m.Init(voxel_size, 1);
    std::vector<PointStruct> pts = to_points(xyz, n);
    for (size_t i = 0; i < n; ++i) pts[i].intensity = static_cast<double>(i);
    const std::vector<PointStruct> out = m.VoxelDownsample(pts, voxel_size);
    for (size_t i = 0; i < out.size(); ++i) index_out[i] = static_cast<int32_t>(out[i].intensity);
    return out.size();
}

}  // extern "C"
