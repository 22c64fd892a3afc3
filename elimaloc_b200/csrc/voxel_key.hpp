// Packed voxel keys and the table's home-slot function, shared by the host builder and the device kernels.
//
// A voxel key is three signed coordinates in [-2^20, 2^20), biased and packed x:y:z into 63 bits; all-ones = empty slot.
// The reference's 20-bit hash
// (voxel_hash_map.hpp:150-155) is not observable behaviour — only key equality is.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ELM_HD __host__ __device__ __forceinline__
#else
#define ELM_HD inline
#endif

namespace elm {

constexpr int kKeyBits = 21;                       // per-axis field width of a packed key
constexpr int32_t kKeyBias = 1 << (kKeyBits - 1);  // keys in [-2^20, 2^20) per axis
constexpr uint64_t kEmptyKey = ~0ull;

ELM_HD bool key_in_range(int32_t k) { return k >= -kKeyBias && k < kKeyBias; }
ELM_HD uint64_t pack_key(int32_t x, int32_t y, int32_t z) {
    return (static_cast<uint64_t>(static_cast<uint32_t>(x + kKeyBias)) << (2 * kKeyBits)) |
           (static_cast<uint64_t>(static_cast<uint32_t>(y + kKeyBias)) << kKeyBits) |
           static_cast<uint64_t>(static_cast<uint32_t>(z + kKeyBias));
}
ELM_HD void unpack_key(uint64_t k, int32_t& x, int32_t& y, int32_t& z) {
    const uint64_t m = (1ull << kKeyBits) - 1;
    x = static_cast<int32_t>((k >> (2 * kKeyBits)) & m) - kKeyBias;
    y = static_cast<int32_t>((k >> kKeyBits) & m) - kKeyBias;
    z = static_cast<int32_t>(k & m) - kKeyBias;
}
// home slot (before masking): a 32-bit multiplicative mix of the whole key.  Locality-preserving variants (keeping
// the low z bits so that z-neighbours share a sector) were measured and rejected: on a dense map they chain whole
// voxel columns into long linear-probing clusters (avg hit 1.9-4.6 probes vs 1.45 for this one at load 0.48).
ELM_HD uint32_t home_slot(uint64_t key) {
    uint32_t h = (static_cast<uint32_t>(key) * 0x9E3779B1u) ^ (static_cast<uint32_t>(key >> 32) * 0x85EBCA77u);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

// second, independent mix: the alternative bucket of the 2-choice (cuckoo) neighbourhood directory
ELM_HD uint32_t home_slot2(uint64_t key) {
    uint32_t h = (static_cast<uint32_t>(key) * 0xC2B2AE3Du) + (static_cast<uint32_t>(key >> 32) * 0x27D4EB2Fu) + 0x165667B1u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
// the two candidate buckets of a key (bucket = 2 slots = one 32-byte sector); always distinct
ELM_HD void dir_buckets(uint64_t key, uint32_t bmask, uint32_t& b1, uint32_t& b2) {
    b1 = home_slot(key) & bmask;
    b2 = home_slot2(key) & bmask;
    if (b2 == b1) b2 = b1 ^ 1u;
}
constexpr uint32_t kDirCountBits = 10;                       // per-voxel count field of a column descriptor
constexpr uint32_t kDirCountMask = (1u << kDirCountBits) - 1;

// ---- directory row: 80 words (320 bytes = ten 32-byte sectors) per directory SLOT ----------------------------------
//   words 0..7   header  {key_lo, key_hi, first VGICP candidate, candidate count, 27-bit occupancy mask, flags, 0, 0}
//   words 8+8c.. column c = 3 (dx + 1) + (dy + 1), one sector:
//                  [0] first stored point of the z-column (x+dx, y+dy, z-1..z+1)   [1] n(z-1) | n(z) << 10 | n(z+1) << 20
//                  [2..7] three 64-bit OCTANT words, one per voxel of the column (z-1, z, z+1):
//                         byte k-1 (k = 1..7) = number of the voxel's points in octants < k, byte 7 = min(n, 255)
// The stored points of a voxel are laid out on the device sorted by octant (hz << 2 | hy << 1 | hx, stable in the canonical
// order inside an octant), so "octants a..b of voxel v" is the contiguous run [first(v) + byte(a-1), first(v) + byte(b)) —
// what the warm-started search reads instead of whole voxels.  Along one axis a voxel with stored key c holds p / vs in
// [c, c+1) for c > 0, (c-1, c] for c < 0 and (-1, 1) for c == 0 (insert keys truncate toward zero, vhm.cpp:275); its upper
// half is p / vs >= the middle of that span.
constexpr int kRowWords = 80, kRowHeaderWords = 8, kRowColWords = 8;
constexpr int kRowKeyLo = 0, kRowKeyHi = 1, kRowCandFirst = 2, kRowCandCount = 3, kRowOccMask = 4, kRowFlags = 5;
constexpr uint32_t kRowFlagOctants = 1u;   // octant words valid (max_points_per_voxel <= 255)
constexpr int kOctantCapMax = 255;
ELM_HD size_t row_word(size_t row, int w) { return row * kRowWords + static_cast<size_t>(w); }
ELM_HD size_t row_col_word(size_t row, int c) { return row * kRowWords + kRowHeaderWords + static_cast<size_t>(c) * kRowColWords; }
// upper (1) or lower (0) half of the span of stored key c along one axis, q = p / voxel_size
ELM_HD int axis_half(double q, int32_t c) {
    const double mid = (c > 0) ? static_cast<double>(c) + 0.5 : ((c < 0) ? static_cast<double>(c) - 0.5 : 0.0);
    return q >= mid ? 1 : 0;
}


// ---- VGICP candidate record (8 bytes) ---------------------------------------------------------------------------------
// The mean of a voxel of an entry's 27-neighbourhood relative to the entry's key, in voxel sizes, lies in (-2, 2) (stored
// keys truncate toward zero, so the voxel at offset o spans (o - 1, o + 1) at most).  13 bits per axis over [-2, 2]:
// step 4 / 8191, error <= 2.45e-4 per axis (4.3e-4 on the vector) — the search's fp32 pre-filter keeps a band of 1.2e-3 voxel
// sizes around the smallest distance and decides anything inside it with the exact fp64 means.  Bits 39..63: voxel index.
constexpr uint32_t kVcandAxisMax = 8191;
constexpr uint32_t kVcandVoxelBits = 25;  // voxel indices < 2^25
ELM_HD uint64_t pack_vcand(double ox, double oy, double oz, uint32_t voxel) {
    auto q = [](double o) {
        double t = (o + 2.0) * (kVcandAxisMax / 4.0) + 0.5;
        t = t < 0.0 ? 0.0 : (t > static_cast<double>(kVcandAxisMax) ? static_cast<double>(kVcandAxisMax) : t);
        return static_cast<uint64_t>(t);
    };
    return q(ox) | (q(oy) << 13) | (q(oz) << 26) | (static_cast<uint64_t>(voxel) << 39);
}
ELM_HD void unpack_vcand(uint64_t r, float& ox, float& oy, float& oz, uint32_t& voxel) {
    const float s = 4.0f / static_cast<float>(kVcandAxisMax);
    ox = static_cast<float>(static_cast<uint32_t>(r) & kVcandAxisMax) * s - 2.0f;
    oy = static_cast<float>(static_cast<uint32_t>(r >> 13) & kVcandAxisMax) * s - 2.0f;
    oz = static_cast<float>(static_cast<uint32_t>(r >> 26) & kVcandAxisMax) * s - 2.0f;
    voxel = static_cast<uint32_t>(r >> 39);
}

}  // namespace elm
