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
bool ekf_current_state(elm_ekf_state& s, double ego[26]);

}  // namespace elm
