// GPU map build (map_build.cu): AddPoints / CalVoxelCovAll / CalPointCovAll on the device, bit-identical to host_map.cpp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace elm {

// the canonical arrays AddPoints produces (HostMap's vkey / vstart / pxyz / porig)
struct GpuCanonicalMap {
    std::vector<uint64_t> vkey;
    std::vector<uint32_t> vstart;
    std::vector<float> pxyz;
    std::vector<uint32_t> porig;
};

// All three run on the CURRENT device, synchronously, and return "" or an error message.
// AddPoints of n raw points into an EMPTY map (the node's only call, pcm_matching.cpp:87).
std::string gpu_add_points(const float* xyz, size_t n, double voxel_size, int cap, GpuCanonicalMap& out);
std::string gpu_cal_voxel_cov(const std::vector<float>& pxyz, const std::vector<uint32_t>& vstart, std::vector<double>& vmean, std::vector<double>& vcov);
// d_dslots / d_drows / bmask: the neighbourhood directory already published on the device (icp_device.cuh)
std::string gpu_cal_point_cov(const std::vector<float>& pxyz, const uint4* d_dslots, const uint32_t* d_drows, uint32_t bmask, double voxel_size,
                              double search_dist, std::vector<double>& pmean, std::vector<double>& pcov, std::vector<double>& pnormal);

}  // namespace elm
