// Host-side (C++) builder of the voxel map: the product's equivalent of the reference's
// VoxelHashMap::AddPoints / CalVoxelCovAll / CalPointCovAll
//   (pcm_matching/src/voxel_hash_map.cpp:270-285, include/voxel_hash_map.hpp:106-148,183-257)
// re-designed for a device-resident layout: instead of an unordered_map of per-voxel vectors the map is
// kept CANONICALLY SORTED — voxels ordered by key (x, then y, then z), points in insertion order inside a
// voxel — which makes the three z-neighbours of a voxel column one contiguous run in memory, and an
// open-addressed hash table (64-bit packed key -> {first point, count}) is built over it for upload.
// The order-dependent AddPointWithSpacing rule is preserved exactly: a stable sort by voxel keeps insertion
// order, and voxels are independent of each other, so the per-voxel sequential filter runs in parallel.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "voxel_key.hpp"

namespace elm {

// One open-addressed slot, 16 bytes: {key lo, key hi, first stored point, stored count}.
struct Slot { uint32_t key_lo, key_hi, start, count; };

struct HostMap {
    double voxel_size = 1.0;
    int cap = 30;

    // canonical sorted storage
    std::vector<uint64_t> vkey;     // V packed keys, ascending
    std::vector<uint32_t> vstart;   // V+1 prefix of stored points
    std::vector<float> pxyz;        // 3P
    std::vector<uint32_t> porig;    // P: running insertion index of the raw point (diagnostics)
    uint64_t n_raw_seen = 0;

    // covariances (row-major 3x3)
    bool has_vcov = false, has_pcov = false;
    std::vector<double> vmean, vcov;            // 3V, 9V
    std::vector<double> pmean, pcov, pnormal;   // 3P, 9P, 3P

    // open-addressed table over vkey (load <= 0.5); slot i <-> voxel slot_voxel[i] (or -1)
    std::vector<Slot> slots;
    std::vector<int32_t> slot_voxel;
    uint32_t mask = 0;

    size_t V() const { return vkey.size(); }
    size_t P() const { return pxyz.size() / 3; }

    // returns "" on success, else an error message (range violation)
    std::string add_points(const float* xyz, size_t n);
    void cal_voxel_cov();
    void cal_point_cov(double search_dist);
    void build_table();
    // voxel index of a packed key or -1
    int64_t find(uint64_t key) const;
};

// Symmetric 3x3 "plane regularisation" used by both covariance passes (voxel_hash_map.hpp:141-144, 241-244):
// returns I - (1 - 1e-3) n n^T with n the unit eigenvector of the smallest eigenvalue (== U diag(1,1,1e-3) V^T of the
// reference's JacobiSVD for a symmetric PSD input).  Degenerate smallest pair: convention documented in DESIGN.md.
void plane_regularize(const double cov[9], double out[9], double normal[3]);

}  // namespace elm
