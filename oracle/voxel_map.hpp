// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.hpp header).
// Parity status: PINNED on oracle/_ref — the reference's own sources compiled against stand-in Eigen/oneTBB headers
// (tests/test_reference_build.py); Eigen's arithmetic kernels themselves stay restated (smallmat.hpp).
//
// CPU restatement of the reference's VoxelHashMap
//   /root/reference/src/app/localization/pcm_matching/include/voxel_hash_map.hpp  (vhm.hpp)
//   /root/reference/src/app/localization/pcm_matching/src/voxel_hash_map.cpp      (vhm.cpp)
// keeping its data structure (std::unordered_map<Voxel, VoxelBlock> with the 20-bit hash, 168-byte
// AoS points) so that it doubles as the structure-faithful CPU timing baseline.
// oneTBB (absent here) only schedules the reference's loops; OpenMP static chunks joined in order
// stand in for tbb::parallel_reduce, whose join is left-to-right, so results are order-identical.
#pragma once
#include <cstdint>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "smallmat.hpp"

namespace orc {

// vhm.hpp:41-53 — default is (Identity, zero mean); observable through quirk Q2.
struct CovStruct {
    M3 cov;
    V3 mean;
    CovStruct() : cov(M3::Identity()), mean() {}
    CovStruct(const M3& c, const V3& m) : cov(c), mean(m) {}
};

// vhm.hpp:55-87 — same fields, same 168-byte footprint (the solver reads pose, local, covariance only).
struct PointStruct {
    V3 pose;
    V3 local;
    CovStruct covariance;
    float vel = 0.f, azi_angle = 0.f, ele_angle = 0.f;
    double intensity = 0.0;
};
static_assert(sizeof(PointStruct) == 168, "PointStruct must keep the reference's 168-byte stride");

struct Voxel {
    int x, y, z;
    bool operator==(const Voxel& o) const { return x == o.x && y == o.y && z == o.z; }
};

struct VoxelHashMap {
    // vhm.hpp:94-149
    struct VoxelBlock {
        std::vector<PointStruct> points;
        CovStruct covariance;
        int num_points;
        double map_resolution;
        void AddPointWithSpacing(const PointStruct& point);  // vhm.hpp:106-113
        void CalVoxelCov();                                  // vhm.hpp:114-148
    };
    // vhm.hpp:150-155 — only 2^20 distinct values (perf only; not observable in results)
    struct VoxelHash {
        size_t operator()(const Voxel& v) const {
            const uint32_t a = static_cast<uint32_t>(v.x), b = static_cast<uint32_t>(v.y), c = static_cast<uint32_t>(v.z);
            return ((1u << 20) - 1u) & (a * 73856093u ^ b * 19349669u ^ c * 83492791u);
        }
    };

    void Init(double voxel_size, int max_points_per_voxel) {  // vhm.cpp:26-29
        voxel_size_ = voxel_size;
        max_points_per_voxel_ = max_points_per_voxel;
    }
    // vhm.hpp:176-180 — QUERY key: floor
    Voxel PointToVoxel(const V3& p, double vs) const {
        return Voxel{static_cast<int>(std::floor(p.x / vs)), static_cast<int>(std::floor(p.y / vs)),
                     static_cast<int>(std::floor(p.z / vs))};
    }
    void AddPoints(const std::vector<PointStruct>& points);             // vhm.cpp:270-285 — INSERT key: truncation
    void CalVoxelCovAll();                                              // vhm.hpp:183-193
    void CalPointCovAll(double d_search_dist);                          // vhm.hpp:195-257
    int GetAdjacentVoxels(const V3& pose, int range, Voxel out[27]) const;  // vhm.cpp:208-243
    // vhm.cpp:31-88 / 90-151 / 153-206.  nthreads: OpenMP static chunks, concatenated in chunk order.
    std::tuple<std::vector<PointStruct>, std::vector<PointStruct>> GetCorrespondencePoints(
        const std::vector<PointStruct>& pts, double max_dist, int nthreads = 1) const;
    std::tuple<std::vector<PointStruct>, std::vector<CovStruct>> GetCorrespondencesCov(
        const std::vector<PointStruct>& pts, double max_dist, int nthreads = 1) const;
    std::tuple<std::vector<PointStruct>, std::vector<CovStruct>> GetCorrespondencesAllCov(
        const std::vector<PointStruct>& pts, double max_dist, int nthreads = 1) const;
    // vhm.hpp:260-283 — first point per FLOOR-keyed cell; output order = unordered_map iteration order in the
    // reference (implementation-defined); the oracle emits first-seen order instead (documented deviation).
    std::vector<PointStruct> VoxelDownsample(const std::vector<PointStruct>& pts, double voxel_size) const;
    bool Empty() const { return map_.empty(); }  // vhm.hpp:325
    size_t NumPoints() const;

    double voxel_size_ = 1.0;
    int max_points_per_voxel_ = 30;
    std::unordered_map<Voxel, VoxelBlock, VoxelHash> map_;
};

}  // namespace orc
