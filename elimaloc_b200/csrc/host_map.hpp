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

namespace elm {

constexpr int kKeyBits = 21;                       // per-axis field width of a packed key
constexpr int32_t kKeyBias = 1 << (kKeyBits - 1);  // keys in [-2^20, 2^20) per axis
constexpr uint64_t kEmptyKey = ~0ull;

inline bool key_in_range(int32_t k) { return k >= -kKeyBias && k < kKeyBias; }
inline uint64_t pack_key(int32_t x, int32_t y, int32_t z) {
    return (static_cast<uint64_t>(static_cast<uint32_t>(x + kKeyBias)) << (2 * kKeyBits)) |
           (static_cast<uint64_t>(static_cast<uint32_t>(y + kKeyBias)) << kKeyBits) |
           static_cast<uint64_t>(static_cast<uint32_t>(z + kKeyBias));
}
inline void unpack_key(uint64_t k, int32_t& x, int32_t& y, int32_t& z) {
    const uint64_t m = (1ull << kKeyBits) - 1;
    x = static_cast<int32_t>((k >> (2 * kKeyBits)) & m) - kKeyBias;
    y = static_cast<int32_t>((k >> kKeyBits) & m) - kKeyBias;
    z = static_cast<int32_t>(k & m) - kKeyBias;
}
// murmur3 finaliser; the reference's 20-bit hash (voxel_hash_map.hpp:150-155) is not observable behaviour.
inline uint64_t mix_key(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

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
