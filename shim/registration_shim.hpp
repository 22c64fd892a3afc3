// registration_shim.hpp — drop-in for pcm_matching's registration.hpp / voxel_hash_map.hpp on top of the C ABI of
// libelimaloc_b200.so, so that pcm_matching.cpp compiles UNCHANGED (it keeps calling local_map_.Init / AddPoints /
// CalVoxelCovAll / CalPointCovAll / Pointcloud / Covariances / VoxelDownsample / FindGroundHeight and
// registration_.Init / RunRegister / TransformPoints exactly as at pcm_matching.cpp:82-105, 258, 280-282, 308-312, 387,
// 408-414).  Header-only; needs Eigen (the node already has it) and only its coefficient access, so it is independent of
// Eigen's storage order.  The build image has no Eigen/ROS: tests/test_shim.py compiles it against a minimal stand-in
// (tests/mock_eigen) and runs the node's call sequence through it, and tests/test_node_on_shim.py compiles the reference's
// own, unmodified pcm_matching.cpp against it.
//
// Reference interface replaced:
//   struct PointStruct / CovStruct          pcm_matching/include/voxel_hash_map.hpp:41-87   (layout kept)
//   struct VoxelHashMap                     pcm_matching/include/voxel_hash_map.hpp:89-335
//   struct RegistrationConfig / Registration pcm_matching/include/registration.hpp:60-230
#pragma once
#include <Eigen/Core>
#include <Eigen/Dense>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "elimaloc_b200.h"

#ifndef ELM_SHIM_DEVICE
#define ELM_SHIM_DEVICE 0  // CUDA device ordinal the node's map and registration live on
#endif

namespace Eigen {
using Matrix6d = Eigen::Matrix<double, 6, 6>;
}

struct CovStruct {
    Eigen::Matrix3d cov;
    Eigen::Vector3d mean;
    CovStruct() : cov(Eigen::Matrix3d::Identity()), mean(Eigen::Vector3d::Zero()) {}
};

struct PointStruct {  // 168 bytes, same field order as the reference
    Eigen::Vector3d pose;
    Eigen::Vector3d local;
    CovStruct covariance;
    float vel = 0, azi_angle = 0, ele_angle = 0;
    double intensity = 0;
    PointStruct() : pose(Eigen::Vector3d::Zero()), local(Eigen::Vector3d::Zero()) {}
};

typedef enum { P2P, GICP, VGICP, AVGICP } IcpMethod;

struct RegistrationConfig {  // registration.hpp:62-85, every field the node fills
    int i_max_thread;
    IcpMethod icp_method;
    int voxel_search_method;
    double gicp_cov_search_dist;
    bool use_radar_cov;
    int max_iteration;
    double max_search_dist, lm_lambda, icp_termination_threshold_m, min_overlap_ratio, max_fitness_score;
    double doppler_trans_lambda, range_variance_m, azimuth_variance_deg, elevation_variance_deg;
    Eigen::Vector3d ego_to_lidar_trans;
    Eigen::Matrix3d ego_to_lidar_rot, ego_to_imu_rot;
    bool b_debug_print;
};

namespace elm_shim {
inline void flatten_into(const std::vector<PointStruct>& pts, std::vector<float>& xyz) {
    xyz.resize(3 * pts.size());  // (a buffer that lives with its owner: no 1.5 MB allocation + first-touch page faults per scan)
    for (size_t i = 0; i < pts.size(); ++i) {  // Pcl2PointStruct widened float fields: narrowing back is lossless
        xyz[3 * i] = static_cast<float>(pts[i].pose.x());
        xyz[3 * i + 1] = static_cast<float>(pts[i].pose.y());
        xyz[3 * i + 2] = static_cast<float>(pts[i].pose.z());
    }
}
inline std::vector<float> flatten(const std::vector<PointStruct>& pts) {
    std::vector<float> xyz;
    flatten_into(pts, xyz);
    return xyz;
}
inline void check(int status) {
    // The reference never throws out of these calls; failures degrade to "ICP FAIL" in the caller.  Two kinds:
    //  - configuration errors (call order, out-of-scope options, map beyond the key range, bad arguments) do not go away by
    //    themselves: the reference would have run on default / stale data, here the call is refused — said LOUDLY, once per
    //    kind, so that a mis-configured drop-in does not just "never localise";
    //  - transient failures (CUDA / NCCL) stay the reference's quiet yellow line, every time.
    if (status == ELM_OK) return;
    const bool config = status == ELM_ERR_STATE || status == ELM_ERR_UNSUPPORTED || status == ELM_ERR_RANGE || status == ELM_ERR_INVALID;
    if (config) {
        static bool said[8] = {false, false, false, false, false, false, false, false};
        if (!said[status & 7]) {
            said[status & 7] = true;
            std::cerr << "\033[1;31m[elimaloc_b200] CONFIGURATION ERROR (status " << status << "): " << elm_last_error()
                      << " -- every registration will report ICP FAIL until this is fixed\033[0m" << std::endl;
        }
        return;
    }
    std::cout << "\033[1;33m[elimaloc_b200] " << elm_last_error() << "\033[0m" << std::endl;
}
}  // namespace elm_shim

struct VoxelHashMap {
    VoxelHashMap() {}
    ~VoxelHashMap() { if (h_) elm_map_destroy(h_); }
    VoxelHashMap(const VoxelHashMap&) = delete;
    VoxelHashMap& operator=(const VoxelHashMap&) = delete;
    void Init(double voxel_size, int max_points_per_voxel) {
        if (h_) elm_map_destroy(h_);
        h_ = nullptr;
        elm_shim::check(elm_map_create(&h_, voxel_size, max_points_per_voxel, ELM_SHIM_DEVICE));
        voxel_size_ = voxel_size;
        max_points_per_voxel_ = max_points_per_voxel;
    }
    void AddPoints(const std::vector<PointStruct>& points) {
        const std::vector<float> xyz = elm_shim::flatten(points);
        elm_shim::check(elm_map_add_points(h_, xyz.data(), points.size()));
    }
    void Update(const std::vector<PointStruct>& points, const Eigen::Vector3d& /*origin*/) { AddPoints(points); }  // voxel_hash_map.cpp:268
    void Clear() { Init(voxel_size_, max_points_per_voxel_); }                                                       // voxel_hash_map.hpp:324
    void CalVoxelCovAll() { elm_shim::check(elm_map_cal_voxel_cov(h_)); }
    void CalPointCovAll(double d_search_dist) { elm_shim::check(elm_map_cal_point_cov(h_, d_search_dist)); }
    bool Empty() const { return !h_ || elm_map_empty(h_); }
    bool FindGroundHeight(const Eigen::Matrix<double, 2, 1>& position, double& ground_z) const {  // voxel_hash_map.hpp:285-322
        int32_t found = 0;
        elm_shim::check(elm_map_find_ground_height(h_, position(0), position(1), &ground_z, &found));
        return found != 0;
    }
    std::vector<PointStruct> Pointcloud() const {  // visualisation only (pcm_matching.cpp:104)
        if (!h_) return {};
        std::vector<float> xyz(3 * elm_map_num_points(h_));
        elm_map_export(h_, nullptr, nullptr, nullptr, nullptr, xyz.data(), nullptr, nullptr);
        std::vector<PointStruct> out(xyz.size() / 3);
        for (size_t i = 0; i < out.size(); ++i) out[i].pose = out[i].local = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        return out;
    }
    std::vector<CovStruct> Covariances() const {  // visualisation only (pcm_matching.cpp:105): voxels with more than 2 points
        std::vector<CovStruct> out;
        if (!h_) return out;
        const size_t nv = elm_map_num_voxels(h_);
        std::vector<int32_t> counts(nv);
        std::vector<double> vmean(3 * nv), vcov(9 * nv);
        if (elm_map_export(h_, nullptr, counts.data(), vmean.data(), vcov.data(), nullptr, nullptr, nullptr) != ELM_OK) return out;
        for (size_t v = 0; v < nv; ++v) {
            if (counts[v] <= 2) continue;
            CovStruct c;
            for (int i = 0; i < 3; ++i) { c.mean(i) = vmean[3 * v + i]; for (int j = 0; j < 3; ++j) c.cov(i, j) = vcov[9 * v + 3 * i + j]; }
            out.push_back(c);
        }
        return out;
    }
    // first point of every floor-keyed voxel (voxel_hash_map.hpp:260-283), emitted in the reference's own order: the same container
    // (std::unordered_map), the same hash VALUES (VoxelHash, voxel_hash_map.hpp:150-155: 32-bit arithmetic, 20-bit mask), the same
    // reserve() and the same insertion sequence give the same iteration order on the same standard library — nothing here knows how
    // the library orders its nodes.  Host work on a few thousand points.
    std::vector<PointStruct> VoxelDownsample(const std::vector<PointStruct>& points, const double voxel_size) const {
        struct Key { int x, y, z; bool operator==(const Key& o) const { return x == o.x && y == o.y && z == o.z; } };
        struct Hash {
            size_t operator()(const Key& k) const {
                return ((1 << 20) - 1) & (static_cast<uint32_t>(k.x) * 73856093 ^ static_cast<uint32_t>(k.y) * 19349669 ^ static_cast<uint32_t>(k.z) * 83492791);
            }
        };
        std::unordered_map<Key, size_t, Hash> grid;
        grid.reserve(points.size());
        for (size_t i = 0; i < points.size(); ++i) {
            const PointStruct& p = points[i];
            const Key k{static_cast<int>(std::floor(p.pose.x() / voxel_size)), static_cast<int>(std::floor(p.pose.y() / voxel_size)),
                        static_cast<int>(std::floor(p.pose.z() / voxel_size))};
            if (grid.find(k) == grid.end()) grid.insert({k, i});
        }
        std::vector<PointStruct> out;
        out.reserve(grid.size());
        for (const auto& kv : grid) out.emplace_back(points[kv.second]);
        return out;
    }
    elm_map* handle() const { return h_; }  // for the scan chain of INTEGRATION.md section 6 (elm_scan_pipeline_*)
    elm_map* h_ = nullptr;
    double voxel_size_ = 1.0;
    int max_points_per_voxel_ = 30;
};

struct Registration {
    Registration() {}
    ~Registration() { if (h_) elm_registration_destroy(h_); }
    void Init(RegistrationConfig config) {
        config_ = config;
        if (!h_) elm_shim::check(elm_registration_create(&h_, ELM_SHIM_DEVICE, /*stream=*/nullptr));
    }
    // registration.hpp:122-124 — identical signature
    Eigen::Matrix4d RunRegister(const std::vector<PointStruct>& source_local, const VoxelHashMap& voxel_map,
                                const Eigen::Matrix4d& initial_guess, RegistrationConfig m_config, bool& is_success,
                                double& fitness_score, Eigen::Matrix6d& local_cov) {
        elm_reg_config c{};
        c.icp_method = static_cast<int32_t>(m_config.icp_method);
        c.max_iteration = m_config.max_iteration;
        c.max_thread = m_config.i_max_thread;
        c.use_radar_cov = m_config.use_radar_cov ? 1 : 0;
        c.debug_print = m_config.b_debug_print ? 1 : 0;
        c.max_search_dist = m_config.max_search_dist;
        c.lm_lambda = m_config.lm_lambda;
        c.icp_termination_threshold_m = m_config.icp_termination_threshold_m;
        c.min_overlap_ratio = m_config.min_overlap_ratio;
        c.max_fitness_score = m_config.max_fitness_score;
        c.range_variance_m = m_config.range_variance_m;
        c.azimuth_variance_deg = m_config.azimuth_variance_deg;
        c.elevation_variance_deg = m_config.elevation_variance_deg;
        std::vector<float>& xyz = xyz_;  // (RunRegister is never re-entered: both call sites of the node hold mutex_pcl_)
        elm_shim::flatten_into(source_local, xyz);
        double T0[16], T[16], cov[36];  // the ABI is row-major; coefficient access keeps this independent of Eigen's storage order
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T0[4 * i + j] = initial_guess(i, j);
        int32_t ok = 0;
        const int st = elm_run_register(h_, voxel_map.h_, xyz.data(), source_local.size(), T0, &c, T, &ok, &fitness_score, cov);
        if (st != ELM_OK) {  // CUDA / NCCL trouble degrades to the reference's soft failure (never throw)
            elm_shim::check(st);
            is_success = false;
            local_cov = Eigen::Matrix6d::Identity();
            return initial_guess;
        }
        is_success = ok != 0;
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) local_cov(i, j) = cov[6 * i + j];
        Eigen::Matrix4d out;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out(i, j) = T[4 * i + j];
        return out;
    }
    // registration.hpp:126-148 — host helpers of the node's visualisation clouds (pcm_matching.cpp:308, 312)
    void TransformPoints(const Eigen::Matrix4d& T, std::vector<PointStruct>& points) const {
        for (PointStruct& p : points) {
            const double x = p.pose.x(), y = p.pose.y(), z = p.pose.z();
            p.pose = Eigen::Vector3d(((T(0, 0) * x + T(0, 1) * y) + T(0, 2) * z) + T(0, 3), ((T(1, 0) * x + T(1, 1) * y) + T(1, 2) * z) + T(1, 3),
                                     ((T(2, 0) * x + T(2, 1) * y) + T(2, 2) * z) + T(2, 3));
        }
    }
    void TransformPoints(const Eigen::Matrix4d& T, const std::vector<PointStruct>& points, std::vector<PointStruct>& o_points) const {
        o_points = points;
        TransformPoints(T, o_points);
    }
    elm_registration* handle() const { return h_; }
    RegistrationConfig config_;
    elm_registration* h_ = nullptr;
    std::vector<float> xyz_;  // packed float xyz of the last scan (staging for elm_run_register)
};
