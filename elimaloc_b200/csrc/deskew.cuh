// Launch interface of deskew.cu
#pragma once
#include <cuda_runtime.h>

namespace elm {

struct DeskewParams {
    int imu_pointer_cur;   // index of the last valid table entry (pcm_matching.cpp:580)
    int imu_available, odom_available;
    int table_stride;      // doubles between the four table rows {time, rot_x, rot_y, rot_z} in `table`
    float odom_incre_x, odom_incre_y, odom_incre_z;
    double time_scan_cur, time_scan_end;
    const int* n_dev;        // may be NULL: the real number of points sits in HBM (n is its upper bound)
    float rel_time_offset;   // subtracted from every point time (DeskewPointCloud's re-basing when the stamp is the scan end, :476-486)
    int run_deskew;          // cfg_.b_run_deskew: 0 copies the points through (:512-525)
};

cudaError_t launch_deskew_points(const float* xyz, const float* rel_time, int n, const DeskewParams& p, const double* table, float* out,
                                 int num_sms, cudaStream_t s);

}  // namespace elm
