// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.hpp header).
// Parity status: PINNED on oracle/_ref — the reference's own sources compiled against stand-in Eigen/oneTBB headers
// (tests/test_reference_build.py); Eigen's arithmetic kernels themselves stay restated (smallmat.hpp).
// Plain-C entry points so that tests/, smoke() and bench.py's cpu_baseline / --impl reference legs can
// drive the CPU restatement through ctypes (oracle/oracle.py).  Never linked into the product.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "registration.hpp"
#include "deskew.hpp"
#include "ekf.hpp"

using namespace orc;

extern "C" {

// Same field order as elm_reg_config in include/elimaloc_b200.h (kept independent on purpose).
struct orc_reg_config {
    int32_t icp_method, max_iteration, max_thread, use_radar_cov, debug_print, reserved0;
    double max_search_dist, lm_lambda, icp_termination_threshold_m, min_overlap_ratio, max_fitness_score;
    double range_variance_m, azimuth_variance_deg, elevation_variance_deg;
};

static RegistrationConfig to_cfg(const orc_reg_config* c) {
    RegistrationConfig r;
    r.icp_method = c->icp_method;
    r.max_iteration = c->max_iteration;
    r.i_max_thread = c->max_thread > 0 ? c->max_thread : 1;
    r.use_radar_cov = c->use_radar_cov != 0;
    r.b_debug_print = c->debug_print != 0;
    r.parallel_accumulate = (c->reserved0 & 1) != 0;
    r.max_search_dist = c->max_search_dist;
    r.lm_lambda = c->lm_lambda;
    r.icp_termination_threshold_m = c->icp_termination_threshold_m;
    r.min_overlap_ratio = c->min_overlap_ratio;
    r.max_fitness_score = c->max_fitness_score;
    r.range_variance_m = c->range_variance_m;
    r.azimuth_variance_deg = c->azimuth_variance_deg;
    r.elevation_variance_deg = c->elevation_variance_deg;
    return r;
}

// pcm.hpp:205-220 Pcl2PointStruct: float fields widened to double, local = pose.
static std::vector<PointStruct> to_points(const float* xyz, size_t n) {
    std::vector<PointStruct> v(n);
    for (size_t i = 0; i < n; ++i) {
        v[i].pose = V3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        v[i].local = v[i].pose;
    }
    return v;
}
static M4 to_m4(const double* t) { M4 m; std::memcpy(m.m, t, sizeof m.m); return m; }

void* orc_map_create(double voxel_size, int max_pts) {
    auto* m = new VoxelHashMap();
    m->Init(voxel_size, max_pts);
    return m;
}
void orc_map_destroy(void* m) { delete static_cast<VoxelHashMap*>(m); }
void orc_map_add_points(void* m, const float* xyz, size_t n) { static_cast<VoxelHashMap*>(m)->AddPoints(to_points(xyz, n)); }
void orc_map_cal_voxel_cov(void* m) { static_cast<VoxelHashMap*>(m)->CalVoxelCovAll(); }
void orc_map_cal_point_cov(void* m, double d) { static_cast<VoxelHashMap*>(m)->CalPointCovAll(d); }
size_t orc_map_num_voxels(void* m) { return static_cast<VoxelHashMap*>(m)->map_.size(); }
size_t orc_map_num_points(void* m) { return static_cast<VoxelHashMap*>(m)->NumPoints(); }

// Canonical dump: voxels sorted by key (x, then y, then z), points in insertion order inside a voxel.
// Any pointer may be NULL.
void orc_map_export(void* mp, int32_t* keys, int32_t* counts, double* vmean, double* vcov, float* pxyz, double* pmean,
                    double* pcov) {
    auto* m = static_cast<VoxelHashMap*>(mp);
    std::vector<const std::pair<const Voxel, VoxelHashMap::VoxelBlock>*> v;
    v.reserve(m->map_.size());
    for (const auto& kv : m->map_) v.push_back(&kv);
    std::sort(v.begin(), v.end(), [](auto a, auto b) {
        if (a->first.x != b->first.x) return a->first.x < b->first.x;
        if (a->first.y != b->first.y) return a->first.y < b->first.y;
        return a->first.z < b->first.z;
    });
    size_t p = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        const auto& vb = v[i]->second;
        if (keys) { keys[3 * i] = v[i]->first.x; keys[3 * i + 1] = v[i]->first.y; keys[3 * i + 2] = v[i]->first.z; }
        if (counts) counts[i] = static_cast<int32_t>(vb.points.size());
        if (vmean) { vmean[3 * i] = vb.covariance.mean.x; vmean[3 * i + 1] = vb.covariance.mean.y; vmean[3 * i + 2] = vb.covariance.mean.z; }
        if (vcov) std::memcpy(vcov + 9 * i, vb.covariance.cov.m, 9 * sizeof(double));
        for (const auto& pt : vb.points) {
            if (pxyz) { pxyz[3 * p] = (float)pt.pose.x; pxyz[3 * p + 1] = (float)pt.pose.y; pxyz[3 * p + 2] = (float)pt.pose.z; }
            if (pmean) { pmean[3 * p] = pt.covariance.mean.x; pmean[3 * p + 1] = pt.covariance.mean.y; pmean[3 * p + 2] = pt.covariance.mean.z; }
            if (pcov) std::memcpy(pcov + 9 * p, pt.covariance.cov.m, 9 * sizeof(double));
            ++p;
        }
    }
}

void* orc_reg_create() { return new Registration(); }
void orc_reg_destroy(void* r) { delete static_cast<Registration*>(r); }

// RunRegister (reg.cpp:274-418).  Trace arrays (may be NULL) hold up to max_trace iterations.
void orc_run_register(void* reg, void* map, const float* src, size_t n, const double* T_init, const orc_reg_config* c,
                      double* T_out, int32_t* is_success, double* fitness, double* local_cov, int32_t max_trace,
                      int32_t* n_iter, double* tr_pose_in, double* tr_JTJ, double* tr_JTr, double* tr_res,
                      double* tr_ncorr, double* tr_pose_out) {
    auto* R = static_cast<Registration*>(reg);
    auto* M = static_cast<VoxelHashMap*>(map);
    const auto pts = to_points(src, n);
    bool ok = (*is_success != 0);
    M6 cov;
    std::vector<IterTrace> trace;
    const M4 T = R->RunRegister(pts, *M, to_m4(T_init), to_cfg(c), ok, *fitness, cov, &trace);
    std::memcpy(T_out, T.m, sizeof T.m);
    *is_success = ok ? 1 : 0;
    std::memcpy(local_cov, cov.m, sizeof cov.m);
    if (n_iter) *n_iter = static_cast<int32_t>(trace.size());
    for (int i = 0; i < (int)trace.size() && i < max_trace; ++i) {
        if (tr_pose_in) std::memcpy(tr_pose_in + 16 * i, trace[i].pose_in.m, 16 * sizeof(double));
        if (tr_JTJ) std::memcpy(tr_JTJ + 36 * i, trace[i].lin.JTJ.m, 36 * sizeof(double));
        if (tr_JTr) std::memcpy(tr_JTr + 6 * i, trace[i].lin.JTr.v, 6 * sizeof(double));
        if (tr_res) tr_res[i] = trace[i].lin.residual_sum;
        if (tr_ncorr) tr_ncorr[i] = (double)trace[i].lin.n_corr;
        if (tr_pose_out) std::memcpy(tr_pose_out + 16 * i, trace[i].pose_out.m, 16 * sizeof(double));
    }
}

void orc_linearize(void* reg, void* map, const float* src, size_t n, const double* T, const orc_reg_config* c,
                   double* JTJ, double* JTr, double* res_sum, long long* n_corr) {
    auto* R = static_cast<Registration*>(reg);
    auto* M = static_cast<VoxelHashMap*>(map);
    const auto pts = to_points(src, n);
    const Linearization lin = R->LinearizeOnce(pts, *M, to_m4(T), to_cfg(c));
    std::memcpy(JTJ, lin.JTJ.m, 36 * sizeof(double));
    std::memcpy(JTr, lin.JTr.v, 6 * sizeof(double));
    *res_sum = lin.residual_sum;
    *n_corr = lin.n_corr;
}

// Correspondence dump for index-level parity.  K = 1 (P2P/GICP/VGICP) or 7 (AVGICP).
// count[i] = pairs emitted for scan point i; target[(i*K + j)*3 ..] = matched position
// (P2P: matched point pose; GICP: same search, pose reported; VGICP/AVGICP: voxel mean).
void orc_correspondences(void* map, const float* src, size_t n, const double* Tp, int method, double max_dist,
                         int32_t* count, double* target) {
    auto* M = static_cast<VoxelHashMap*>(map);
    const M4 T = to_m4(Tp);
    const int K = (method == AVGICP) ? 7 : 1;
    std::vector<PointStruct> one(1), sc, tc;
    std::vector<CovStruct> tcc;
    for (size_t i = 0; i < n; ++i) {
        one[0].local = V3(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
        one[0].pose = apply(T, one[0].local);
        for (int j = 0; j < K * 3; ++j) target[i * K * 3 + j] = 0.0;
        if (method == P2P || method == GICP) {
            std::tie(sc, tc) = M->GetCorrespondencePoints(one, max_dist, 1);
            count[i] = (int32_t)sc.size();
            if (!tc.empty()) { target[i * 3] = tc[0].pose.x; target[i * 3 + 1] = tc[0].pose.y; target[i * 3 + 2] = tc[0].pose.z; }
        } else {
            if (method == VGICP) std::tie(sc, tcc) = M->GetCorrespondencesCov(one, max_dist, 1);
            else std::tie(sc, tcc) = M->GetCorrespondencesAllCov(one, max_dist, 1);
            count[i] = (int32_t)sc.size();
            for (size_t j = 0; j < tcc.size(); ++j) {
                target[(i * K + j) * 3] = tcc[j].mean.x;
                target[(i * K + j) * 3 + 1] = tcc[j].mean.y;
                target[(i * K + j) * 3 + 2] = tcc[j].mean.z;
            }
        }
    }
}

// Timed CPU baseline: `iters` full ICP iterations (search parallel over `threads`, accumulate + transform serial,
// exactly the reference's structure).  Returns wall seconds of the RunRegister call.
double orc_time_register(void* reg, void* map, const float* src, size_t n, const double* T_init,
                         const orc_reg_config* c, int32_t* iters_done) {
    auto* R = static_cast<Registration*>(reg);
    auto* M = static_cast<VoxelHashMap*>(map);
    const auto pts = to_points(src, n);
    bool ok = false;
    double fit = 0.0;
    M6 cov;
    std::vector<IterTrace> trace;
    const auto t0 = std::chrono::steady_clock::now();
    R->RunRegister(pts, *M, to_m4(T_init), to_cfg(c), ok, fit, cov, &trace);
    const auto t1 = std::chrono::steady_clock::now();
    if (iters_done) *iters_done = (int32_t)trace.size();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- deskew (pcm_matching.cpp:467-824) -------------------------------------------------------------------------------
// tables: out arrays of kImuQueueLength doubles each; meta_i = {imu_pointer_cur, imu_available, odom_available};
// meta_f = {odom_incre_x, y, z}
void orc_deskew_tables(const double* imu_stamp, const double* gyro_xyz, int n_imu, const double* start_pose, double start_stamp,
                       const double* end_pose, double end_stamp, double time_scan_cur, double time_scan_end, double* imu_time,
                       double* rot_x, double* rot_y, double* rot_z, int32_t* meta_i, float* meta_f) {
    DeskewTables t;
    t.time_scan_cur = time_scan_cur;
    t.time_scan_end = time_scan_end;
    ImuDeskewInfo(imu_stamp, gyro_xyz, n_imu, t);
    if (start_pose && end_pose) OdomDeskewInfo(start_pose, start_stamp, end_pose, end_stamp, t);
    std::memcpy(imu_time, t.imu_time.data(), kImuQueueLength * sizeof(double));
    std::memcpy(rot_x, t.imu_rot_x.data(), kImuQueueLength * sizeof(double));
    std::memcpy(rot_y, t.imu_rot_y.data(), kImuQueueLength * sizeof(double));
    std::memcpy(rot_z, t.imu_rot_z.data(), kImuQueueLength * sizeof(double));
    meta_i[0] = t.imu_pointer_cur; meta_i[1] = t.imu_available; meta_i[2] = t.odom_available;
    meta_f[0] = t.odom_incre_x; meta_f[1] = t.odom_incre_y; meta_f[2] = t.odom_incre_z;
}

void orc_deskew_points(const double* imu_time, const double* rot_x, const double* rot_y, const double* rot_z, const int32_t* meta_i,
                       const float* meta_f, double time_scan_cur, double time_scan_end, const float* xyz, const float* rel_time, size_t n,
                       float* out) {
    DeskewTables t;
    const int e = meta_i[0] + 1;
    t.imu_time.assign(imu_time, imu_time + e); t.imu_rot_x.assign(rot_x, rot_x + e);
    t.imu_rot_y.assign(rot_y, rot_y + e); t.imu_rot_z.assign(rot_z, rot_z + e);
    t.imu_pointer_cur = meta_i[0]; t.imu_available = meta_i[1] != 0; t.odom_available = meta_i[2] != 0;
    t.odom_incre_x = meta_f[0]; t.odom_incre_y = meta_f[1]; t.odom_incre_z = meta_f[2];
    t.time_scan_cur = time_scan_cur; t.time_scan_end = time_scan_end;
    DeskewPoints(t, xyz, rel_time, n, out);
}

// ---- EKF (ekf_algorithm.cpp) -------------------------------------------------------------------------------------------
size_t orc_ekf_state_size() { return sizeof(EkfStateBlob); }
void orc_ekf_init(const EkfConfig* c, EkfStateBlob* s) { EkfInit(*c, *s); }
int orc_ekf_predict_imu(const EkfConfig* c, EkfStateBlob* s, double t, const double* gyro, const double* acc) { return EkfPredictImu(*c, *s, t, gyro, acc) ? 1 : 0; }
int orc_ekf_update_pose(const EkfConfig* c, EkfStateBlob* s, const EkfMeasurement* m) { return EkfUpdatePose(*c, *s, *m) ? 1 : 0; }
void orc_ekf_get_current_state(EkfStateBlob* s, double* ego) { EkfGetCurrentState(*s, ego); }

// ---- FindGroundHeight (vhm.hpp:285-322): the reference's full scan over Pointcloud() ------------------------------------
int orc_find_ground_height(void* mp, double x, double y, double* ground_z) {
    auto* M = static_cast<VoxelHashMap*>(mp);
    const double range2 = 5.0 * 5.0;                                           // :286-287
    std::vector<double> zs;
    for (const auto& kv : M->map_)                                             // Pointcloud(), :288
        for (const PointStruct& p : kv.second.points) {
            const double dx = p.pose.x - x, dy = p.pose.y - y;
            if (dx * dx + dy * dy <= range2) zs.push_back(p.pose.z);           // :291-296
        }
    if (zs.size() <= 3) return 0;                                              // :298-300
    const size_t n = std::min<size_t>(5, zs.size());
    std::partial_sort(zs.begin(), zs.begin() + static_cast<std::ptrdiff_t>(n), zs.end());   // :303-306 (by z)
    double sum = 0.0;
    for (size_t i = 0; i < n; ++i) sum += zs[i];
    *ground_z = sum / static_cast<double>(n);                                  // :318-319
    return 1;
}

// ---- scan pre-processing: FilterPointsByDistance (pcm_matching.cpp:451-465) then VoxelDownsample (vhm.hpp:260-283) ----
// Writes the INPUT INDEX of every survivor (input order); returns their number.  max_dist <= 0 / voxel_size <= 0 skip a step.
size_t orc_scan_preprocess(const float* xyz, size_t n, double max_dist, double voxel_size, int32_t* index_out) {
    std::vector<size_t> kept;
    kept.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        if (max_dist > 0.0) {
            const double distance = std::sqrt(x * x + y * y + z * z);   // float expression, float sqrt, widened (cpp:456)
            if (distance > max_dist) continue;                          // cpp:457
        }
        kept.push_back(i);
    }
    size_t m = 0;
    if (voxel_size > 0.0) {
        std::vector<PointStruct> pts(kept.size());
        for (size_t k = 0; k < kept.size(); ++k) {
            pts[k].pose = V3(xyz[3 * kept[k]], xyz[3 * kept[k] + 1], xyz[3 * kept[k] + 2]);
            pts[k].local = pts[k].pose;
            pts[k].intensity = static_cast<double>(k);                  // carries the position through VoxelDownsample
        }
        VoxelHashMap grid_owner;
        grid_owner.Init(1.0, 1);
        for (const PointStruct& p : grid_owner.VoxelDownsample(pts, voxel_size)) index_out[m++] = static_cast<int32_t>(kept[static_cast<size_t>(p.intensity)]);
    } else {
        for (size_t k : kept) index_out[m++] = static_cast<int32_t>(k);
    }
    return m;
}

// ---- result shaping (PcmMatching::PublishPcmOdom, pcm_matching.cpp:1082-1098; NormalizeCovariance, pcm_matching.hpp:247-273)
// Test infrastructure like everything in oracle/: restates the reference line by line in plain loops.
static void orc_normalize_cov(const double in[9], double out[9]) {
    double c[9];
    for (int i = 0; i < 9; ++i) c[i] = in[i];
    double min_diag = std::min(c[0], std::min(c[4], c[8]));                 // hpp:252
    const double min_threshold = 1e-9;
    if (min_diag <= min_threshold) {                                        // hpp:256
        for (int i = 0; i < 9; ++i) c[i] *= 1e9;                            // hpp:257
        min_diag = std::min(c[0], std::min(c[4], c[8]));
        if (min_diag < min_threshold) min_diag = min_threshold;             // hpp:260
    }
    for (int i = 0; i < 9; ++i) out[i] = std::min(c[i] / min_diag, 5.0);    // hpp:264-268
}
void orc_shape_pcm_covariance(const double* R, const double* local_cov, double icp_pose_std_m, double* cov36) {
    const double std_m = std::max(icp_pose_std_m, 0.25);                    // cpp:1082
    double RC[9], tc[9], rc[9], tn[9], rn[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        RC[3 * i + j] = 0.0;
        for (int k = 0; k < 3; ++k) RC[3 * i + j] += R[3 * i + k] * local_cov[6 * k + j];
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {                // cpp:1085-1086: R C R^T
        tc[3 * i + j] = 0.0;
        for (int k = 0; k < 3; ++k) tc[3 * i + j] += RC[3 * i + k] * R[3 * j + k];
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rc[3 * i + j] = local_cov[6 * (i + 3) + (j + 3)];  // cpp:1090
    const double angle_std = std_m * M_PI / 180.0;                          // cpp:1093
    orc_normalize_cov(tc, tn);
    orc_normalize_cov(rc, rn);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {                // UpdateCovarianceField, hpp:275-290
        cov36[6 * i + j] = tn[3 * i + j] * std_m * std_m;
        cov36[6 * (i + 3) + (j + 3)] = rn[3 * i + j] * angle_std * angle_std;
    }
}

// ---- the restated third-party kernels of smallmat.hpp on their own, for tests/test_oracle_kernels.py (checked against LAPACK
// through numpy): everything row-major.
void orc_k_inverse(int n, const double* a, double* out) {
    if (n == 3) { M3 m; std::memcpy(m.m, a, sizeof m.m); const M3 r = inverse(m); std::memcpy(out, r.m, sizeof r.m); }
    else if (n == 4) { M4 m; std::memcpy(m.m, a, sizeof m.m); const M4 r = inverse(m); std::memcpy(out, r.m, sizeof r.m); }
    else if (n == 6) { M6 m; std::memcpy(m.m, a, sizeof m.m); const M6 r = inverse(m); std::memcpy(out, r.m, sizeof r.m); }
}
void orc_k_ldlt_solve(const double* a36, const double* b6, double* x6) {
    M6 A; V6 b;
    std::memcpy(A.m, a36, sizeof A.m);
    std::memcpy(b.v, b6, sizeof b.v);
    const V6 x = ldlt_solve(A, b);
    std::memcpy(x6, x.v, sizeof x.v);
}
void orc_k_sym_eig3(const double* a9, double* w3, double* v9) {
    M3 A, V;
    std::memcpy(A.m, a9, sizeof A.m);
    sym_eig3(A, w3, V);
    std::memcpy(v9, V.m, sizeof V.m);
}
void orc_k_plane_regularize(const double* a9, double* out9, double* normal3) {
    M3 A;
    std::memcpy(A.m, a9, sizeof A.m);
    V3 n;
    const M3 r = plane_regularize(A, &n);
    std::memcpy(out9, r.m, sizeof r.m);
    normal3[0] = n.x; normal3[1] = n.y; normal3[2] = n.z;
}
void orc_k_angle_axis_to_rot(double angle, const double* axis3, double* r9) {
    const M3 r = angle_axis_to_rot(angle, V3(axis3[0], axis3[1], axis3[2]));
    std::memcpy(r9, r.m, sizeof r.m);
}
double orc_k_rot_angle(const double* r9) {
    M3 R;
    std::memcpy(R.m, r9, sizeof R.m);
    return rot_angle(R);
}

}  // extern "C"
