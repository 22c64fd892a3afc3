// Launch interface of icp_kernels.cu (see there for what each kernel restates).
#pragma once
#include "icp_device.cuh"

namespace elm {

constexpr int kIcpWarps = 8;
constexpr int kIcpThreads = kIcpWarps * 32;

struct Pose16 { double m[16]; };

// blocks the linearisation kernel is launched with (== rows of `partials`)
int icp_linearize_grid(const IcpParams& prm, int num_sms);

cudaError_t launch_icp_begin(IcpState* st, const double T0[16], cudaStream_t s);
cudaError_t launch_icp_linearize(const MapView& map, const float* scan, const IcpParams& prm, const IcpState* st,
                                 double* partials, int grid, cudaStream_t s);
cudaError_t launch_icp_reduce(IcpState* st, const double* partials, int nblocks, cudaStream_t s);
cudaError_t launch_icp_solve(IcpState* st, const IcpParams& prm, cudaStream_t s);
cudaError_t launch_icp_match(const MapView& map, const float* scan, int n, const double T[16], int method,
                             double max_dist2, int* count, double* target, int num_sms, cudaStream_t s);

}  // namespace elm
