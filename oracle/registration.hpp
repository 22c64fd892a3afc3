// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.hpp header).
// Parity status: PINNED on oracle/_ref — the reference's own sources compiled against stand-in Eigen/oneTBB headers
// (tests/test_reference_build.py); Eigen's arithmetic kernels themselves stay restated (smallmat.hpp).
// CPU restatement of the reference's Registration
//   /root/reference/src/app/localization/pcm_matching/include/registration.hpp  (reg.hpp)
//   /root/reference/src/app/localization/pcm_matching/src/registration.cpp      (reg.cpp)
#pragma once
#include <vector>

#include "voxel_map.hpp"

namespace orc {

enum IcpMethod { P2P = 0, GICP = 1, VGICP = 2, AVGICP = 3 };  // reg.hpp:60

// reg.hpp:62-85 — only the fields the solver reads (reg.cpp:302-412) plus the thread cap.
struct RegistrationConfig {
    int i_max_thread = 1;
    int icp_method = P2P;
    bool use_radar_cov = false;
    int max_iteration = 10;
    double max_search_dist = 5.0;
    double lm_lambda = 0.5;
    double icp_termination_threshold_m = 0.02;
    double min_overlap_ratio = 0.4;
    double max_fitness_score = 0.5;
    double range_variance_m = 1.0;
    double azimuth_variance_deg = 0.4;
    double elevation_variance_deg = 0.4;
    bool b_debug_print = false;
    // NOT in the reference: "best-effort CPU" timing variant — AlignClouds* and TransformPoints run over
    // i_max_thread static chunks whose partial sums are joined in chunk order.  Off: serial loops exactly as reg.cpp has them.
    bool parallel_accumulate = false;
};

// One linearisation (what a single AlignClouds* call accumulates before its solve).
struct Linearization {
    M6 JTJ;
    V6 JTr;
    double residual_sum = 0.0;
    long long n_corr = 0;
};

// Per-iteration trace for the parity tests.
struct IterTrace {
    M4 pose_in;   // last_icp_pose going into the iteration
    Linearization lin;
    M4 pose_out;  // after the right-multiplied update
};

struct Registration {
    // reg.cpp:15-66 / 68-152 / 154-225.  `lin` (optional) receives the raw sums.
    M4 AlignCloudsLocal(const std::vector<PointStruct>& source_global, const std::vector<PointStruct>& target_global,
                        const M4& last_icp_pose, double trans_th, const RegistrationConfig& cfg, Linearization* lin);
    M4 AlignCloudsLocalPointCov(const std::vector<PointStruct>& source_global,
                                const std::vector<PointStruct>& target_global, M6& local_cov, const M4& last_icp_pose,
                                double trans_th, const RegistrationConfig& cfg, Linearization* lin);
    M4 AlignCloudsLocalVoxelCov(const std::vector<PointStruct>& source_global,
                                const std::vector<CovStruct>& target_cov_global, const M4& last_icp_pose,
                                double trans_th, const RegistrationConfig& cfg, Linearization* lin);
    // reg.cpp:274-418
    M4 RunRegister(const std::vector<PointStruct>& source_local, const VoxelHashMap& voxel_map, const M4& initial_guess,
                   const RegistrationConfig& cfg, bool& is_success, double& fitness_score, M6& local_cov,
                   std::vector<IterTrace>* trace = nullptr);
    // reg.hpp:136-148
    static void TransformPoints(const M4& T, const std::vector<PointStruct>& points, std::vector<PointStruct>& o_points);
    // search + one AlignClouds* accumulation at a fixed pose (no solve): the single-iteration parity hook.
    Linearization LinearizeOnce(const std::vector<PointStruct>& source_local, const VoxelHashMap& voxel_map,
                                const M4& pose, const RegistrationConfig& cfg);

    double d_fitness_score_ = 0.0;  // reg.hpp:229 — persists across calls
};

}  // namespace orc
