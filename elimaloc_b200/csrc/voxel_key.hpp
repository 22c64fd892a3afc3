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

}  // namespace elm
