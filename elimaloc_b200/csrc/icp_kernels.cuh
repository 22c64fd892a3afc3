// Launch interface of icp_kernels.cu (see there for what each kernel restates).
#pragma once
#include "icp_device.cuh"

namespace elm {

constexpr int kIcpWarps = 8;
constexpr int kIcpThreads = kIcpWarps * 32;

struct Pose16 { double m[16]; };

int icp_search_grid(const IcpParams& prm, int num_sms);
int icp_warm_grid(const IcpParams& prm, int num_sms);
// blocks of the accumulation kernel (== rows of `partials`)
int icp_accumulate_grid(const IcpParams& prm, int num_sms);

cudaError_t launch_icp_begin(IcpState* st, const double T0[16], unsigned int* ticket, unsigned long long* tile_ticket, cudaStream_t s);
// P2P / GICP / VGICP correspondence search -> wk.match[n] (+ wk.win, wk.memo for P2P / GICP); no-op for AVGICP, which searches
// inside the accumulation.  `orig` (may be NULL): scan is in binned order and the outputs are written at orig[i].
// fuse (P2P / GICP): the search kernel also linearises, reduces and — when solve_here — solves: one launch per iteration.
cudaError_t launch_icp_search(const MapView& map, const float* scan, const int* orig, const IcpParams& prm, IcpState* st, const IcpWork& wk,
                              int grid, int prune, int fuse, int solve_here, cudaStream_t s);
cudaError_t launch_icp_accumulate(const MapView& map, const float* scan, const IcpParams& prm, IcpState* st, const IcpWork& wk, int solve_here,
                                  int grid, cudaStream_t s);
// One WARM iteration of P2P / GICP — search from the previous iteration's wk.win / wk.memo / candidate lists of the SAME scan,
// linearisation, reduction, solve — as two launches (icp_kernels.cu); wk.partials needs reuse_grid + refresh_grid rows.
int icp_warm_refresh_grid(const IcpParams& prm, int num_sms);
cudaError_t launch_icp_warm_reuse(const MapView& map, const float* scan, const IcpParams& prm, const IcpState* st, const IcpWork& wk, int reuse_grid,
                                  cudaStream_t s);
cudaError_t launch_icp_warm_refresh(const MapView& map, const float* scan, const IcpParams& prm, IcpState* st, const IcpWork& wk, int reuse_grid,
                                    int refresh_grid, int solve_here, cudaStream_t s);
// One warm iteration as ONE launch (stragglers refreshed in place by their own warp); wk.partials needs `grid` rows.
cudaError_t launch_icp_warm(const MapView& map, const float* scan, const IcpParams& prm, IcpState* st, const IcpWork& wk, int grid, int solve_here,
                            cudaStream_t s);
// The refresh kernel running BESIDE the reuse kernel of the same iteration (wk.epoch != 0: the reuse kernel publishes every tile's
// work list as soon as it is complete); grid = icp_warm_refresh_async_grid blocks of 128 threads; wk.partials needs reuse_grid + grid rows.
int icp_warm_refresh_async_grid(int num_sms, int want = 0);  // want: blocks asked for (0: the default, 80)
cudaError_t launch_icp_warm_refresh_async(const MapView& map, const float* scan, const IcpParams& prm, IcpState* st, const IcpWork& wk, int reuse_grid,
                                          int grid, int solve_here, cudaStream_t s);
cudaError_t launch_icp_solve(IcpState* st, const IcpParams& prm, cudaStream_t s);
// spatial binning of the scan (scan_sort.cu)
int scan_bin_bits(int n);
cudaError_t launch_scan_binning(const float* scan, int n, const double T[16], double voxel_size, uint32_t* bin, uint32_t* hist,
                                float* sorted, int* orig, cudaStream_t s);

cudaError_t launch_icp_export(const MapView& map, const float* scan, const int* match, int n, const IcpState* st, int method,
                              double max_dist2, int* count, double* target, cudaStream_t s);

}  // namespace elm
