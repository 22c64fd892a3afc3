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
};

cudaError_t launch_deskew_points(const float* xyz, const float* rel_time, int n, const DeskewParams& p, const double* table, float* out,
                                 int num_sms, cudaStream_t s);

}  // namespace elm
