// Launch interface of ekf.cu
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>

#include "../../include/elimaloc_b200.h"

namespace elm {

cudaError_t launch_ekf_predict_imu(elm_ekf_state* s, const elm_ekf_config& c, double t, const double g[3], const double a[3], cudaStream_t st);
cudaError_t launch_ekf_update_pose(elm_ekf_state* s, const elm_ekf_config& c, const elm_ekf_measurement& m, cudaStream_t st);
void ekf_init_state(const elm_ekf_config& c, elm_ekf_state& s);
#if defined(__CUDACC__)
__host__ __device__
#endif
bool ekf_current_state(elm_ekf_state& s, double ego[26]);

// ring of EgoStates in HBM (PublishInThread's deque, ekf_localization.cpp:398-410): e[cap][8] = {t, x, y, z, roll, pitch, yaw, 0};
// meta[0] = index of the oldest entry, meta[1] = entries
struct EkfRing { double* e; int* meta; int cap; };
constexpr int kEkfRingCap = 1000;
// what ekf_update_from_icp_kernel needs beside the IcpState: stamp of the measurement (d_time_scan_end_), tf_ego_to_lidar^-1,
// RegistrationConfig::max_fitness_score, and whether the registration was skipped (empty map / empty scan)
struct EkfIcpParams { double stamp; double T_lidar_to_ego[16]; double max_fitness; int trivial; int reserved; };
struct IcpState;
cudaError_t launch_ekf_update_from_icp(elm_ekf_state* s, const elm_ekf_config& c, const IcpState* icp, const EkfIcpParams& p, const EkfRing& ring, cudaStream_t st);
cudaError_t launch_ekf_ring_push(elm_ekf_state* s, const EkfRing& ring, cudaStream_t st);

// PcmMatching::PublishPcmOdom's pose covariance (pcm_matching.cpp:1082-1098, pcm_matching.hpp:247-290): host and device
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void shape_pcm_covariance_hd(const double R_ego[9], const double local_cov[36], double icp_pose_std_m, double pose_cov[36]) {
    const double sd = icp_pose_std_m > 0.25 ? icp_pose_std_m : 0.25, ang = sd * 3.14159265358979323846 / 180.0;  // std::max(d_icp_pose_std_m, 0.25)
    double t[9], r[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = 0.0;
            for (int a = 0; a < 3; ++a) {
                double rc = 0.0;
                for (int k = 0; k < 3; ++k) rc += R_ego[3 * i + k] * local_cov[6 * k + a];
                acc += rc * R_ego[3 * j + a];
            }
            t[3 * i + j] = acc;
            r[3 * i + j] = local_cov[6 * (i + 3) + (j + 3)];
        }
    for (int b = 0; b < 2; ++b) {
        const double* in = b ? r : t;
        double scale = 1.0;
        auto min3 = [](double a, double c, double d) { const double m = a < c ? a : c; return m < d ? m : d; };
        double m = min3(in[0], in[4], in[8]);
        if (m <= 1e-9) { scale = 1e9; m = min3(in[0] * scale, in[4] * scale, in[8] * scale); if (m < 1e-9) m = 1e-9; }
        const double f = b ? ang * ang : sd * sd;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double v = in[3 * i + j] * scale / m;
                if (v > 5.0) v = 5.0;
                pose_cov[6 * (i + 3 * b) + (j + 3 * b)] = v * f;
            }
    }
}

}  // namespace elm
