// Scan pre-processing on the GPU: the two serial steps the node runs on every scan right before
// Registration::RunRegister,
//   PcmMatching::FilterPointsByDistance   pcm_matching/src/pcm_matching.cpp:451-465   keep iff sqrt(x^2+y^2+z^2) <= max_dist
//   VoxelHashMap::VoxelDownsample         pcm_matching/include/voxel_hash_map.hpp:260-283  first point of every
//                                         FLOOR-keyed voxel (PointToVoxel, :176-180) in input order
// fused into one stable compaction so that a raw scan is uploaded once and never comes back to the host.
//
// Arithmetic follows the reference exactly: the distance is float32 (x*x + y*y + z*z with one rounding per operation,
// std::sqrt(float)) widened to double for the comparison; the voxel key is floor(double(p) / voxel_size) per axis.
// Order: the reference copies the survivors out of a std::unordered_map, i.e. in an implementation-defined order; here
// they keep their input order — the ICP sums do not depend on the order of the scan beyond rounding.
//
//   mark kernel     thread per point: distance test, voxel key, claim the voxel in an open-addressed table
//                   (atomicCAS on the key, atomicMin on the smallest input index seen for it)
//   select kernel   thread per point: survivor iff it passed the distance test and owns its voxel's minimum index;
//                   per-block survivor counts
//   offsets kernel  one block: exclusive scan of the block counts (+ the total)
//   scatter kernel  thread per point: block-local rank by ballot/popc + block offset -> stable output position
#include "scan_prep.cuh"

#include "voxel_key.hpp"

namespace elm {

namespace {

constexpr int kPrepThreads = 256;

__device__ __forceinline__ bool passes_distance(float x, float y, float z, double max_dist) {
    if (!(max_dist > 0.0)) return true;  // filter disabled
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    return !(static_cast<double>(d) > max_dist);  // remove_if(distance > d_input_max_dist), :457-458
}

// floor(p / vs) per axis (vhm.hpp:176-180); false when a coordinate is not finite or leaves the packable key range
__device__ __forceinline__ bool voxel_key_of(float x, float y, float z, double vs, uint64_t& key) {
    const double qx = floor(__ddiv_rn(static_cast<double>(x), vs)), qy = floor(__ddiv_rn(static_cast<double>(y), vs)),
                 qz = floor(__ddiv_rn(static_cast<double>(z), vs));
    const double lim = static_cast<double>(kKeyBias - 1);
    if (!(fabs(qx) < lim && fabs(qy) < lim && fabs(qz) < lim)) return false;
    key = pack_key(static_cast<int32_t>(qx), static_cast<int32_t>(qy), static_cast<int32_t>(qz));
    return true;
}

__global__ void __launch_bounds__(kPrepThreads)
prep_mark_kernel(const float* __restrict__ xyz, int n, const int* __restrict__ n_dev, double max_dist, double voxel_size, unsigned long long* __restrict__ tkeys,
                 uint32_t* __restrict__ tmin, uint32_t tmask, int* __restrict__ error) {
    const int i = blockIdx.x * kPrepThreads + threadIdx.x;
    if (n_dev) n = min(n, *n_dev);  // the real length sits in HBM (an earlier compaction's count): n is only its upper bound
    if (i >= n) return;
    const float x = xyz[3 * static_cast<size_t>(i)], y = xyz[3 * static_cast<size_t>(i) + 1], z = xyz[3 * static_cast<size_t>(i) + 2];
    if (!passes_distance(x, y, z, max_dist)) return;
    uint64_t key;
    if (!voxel_key_of(x, y, z, voxel_size, key)) { atomicExch(error, 1); return; }
    uint32_t h = home_slot(key) & tmask;
    for (uint32_t probes = 0; probes <= tmask; ++probes) {
        const unsigned long long old = atomicCAS(tkeys + h, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
        if (old == kEmptyKey || old == key) { atomicMin(tmin + h, static_cast<uint32_t>(i)); return; }
        h = (h + 1) & tmask;
    }
}

__device__ __forceinline__ bool survives(const float* __restrict__ xyz, int i, double max_dist, double voxel_size,
                                         const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tmin, uint32_t tmask) {
    const float x = xyz[3 * static_cast<size_t>(i)], y = xyz[3 * static_cast<size_t>(i) + 1], z = xyz[3 * static_cast<size_t>(i) + 2];
    if (!passes_distance(x, y, z, max_dist)) return false;
    if (!(voxel_size > 0.0)) return true;  // down-sampling disabled
    uint64_t key;
    if (!voxel_key_of(x, y, z, voxel_size, key)) return false;
    uint32_t h = home_slot(key) & tmask;
    for (uint32_t probes = 0; probes <= tmask; ++probes) {
        const unsigned long long k = tkeys[h];
        if (k == key) return tmin[h] == static_cast<uint32_t>(i);
        if (k == kEmptyKey) return false;
        h = (h + 1) & tmask;
    }
    return false;
}

__global__ void __launch_bounds__(kPrepThreads)
prep_select_kernel(const float* __restrict__ xyz, int n, const int* __restrict__ n_dev, double max_dist, double voxel_size, const unsigned long long* __restrict__ tkeys,
                   const uint32_t* __restrict__ tmin, uint32_t tmask, uint8_t* __restrict__ keep, uint32_t* __restrict__ block_count) {
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if (n_dev) n = min(n, *n_dev);
    const int i = blockIdx.x * kPrepThreads + threadIdx.x;
    const bool k = i < n && survives(xyz, i, max_dist, voxel_size, tkeys, tmin, tmask);
    if (i < n) keep[i] = k ? 1 : 0;
    const uint32_t b = __ballot_sync(0xffffffffu, k);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_cnt, static_cast<uint32_t>(__popc(b)));
    __syncthreads();
    if (threadIdx.x == 0) block_count[blockIdx.x] = s_cnt;
}

// one block: block_offset[b] = sum of block_count[0..b), total -> *n_out
__global__ void __launch_bounds__(1024) prep_offsets_kernel(const uint32_t* __restrict__ block_count, int nblocks, uint32_t* __restrict__ block_offset,
                                                            int* __restrict__ n_out) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int b = base + threadIdx.x;
        const uint32_t v = b < nblocks ? block_count[b] : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = s_warp[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += t; }
            s_warp[threadIdx.x] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t warp_excl = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u;
        if (b < nblocks) block_offset[b] = s_carry + warp_excl + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = static_cast<int>(s_carry);
}

__global__ void __launch_bounds__(kPrepThreads)
prep_scatter_kernel(const float* __restrict__ xyz, const float* __restrict__ aux, int n, const int* __restrict__ n_dev, const uint8_t* __restrict__ keep,
                    const uint32_t* __restrict__ block_offset, float* __restrict__ xyz_out, float* __restrict__ aux_out, int* __restrict__ index_out) {
    __shared__ uint32_t s_warp[kPrepThreads / 32];
    if (n_dev) n = min(n, *n_dev);
    const int i = blockIdx.x * kPrepThreads + threadIdx.x;
    const bool k = i < n && keep[i];
    const uint32_t b = __ballot_sync(0xffffffffu, k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = static_cast<uint32_t>(__popc(b));
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (!k) return;
    const size_t o = static_cast<size_t>(block_offset[blockIdx.x]) + before + static_cast<uint32_t>(__popc(b & ((1u << lane) - 1u)));
    xyz_out[3 * o] = xyz[3 * static_cast<size_t>(i)];
    xyz_out[3 * o + 1] = xyz[3 * static_cast<size_t>(i) + 1];
    xyz_out[3 * o + 2] = xyz[3 * static_cast<size_t>(i) + 2];
    if (aux_out) aux_out[o] = aux[i];
    if (index_out) index_out[o] = i;
}

}  // namespace

size_t scan_prep_table_slots(size_t n) {
    size_t c = 64;
    while (c < 2 * n) c <<= 1;
    return c;
}

cudaError_t launch_scan_prep(const float* xyz, const float* aux, int n, double max_dist, double voxel_size, const ScanPrepScratch& w,
                             float* xyz_out, float* aux_out, int* index_out, int* n_out, cudaStream_t s, const int* n_dev) {
    const int blocks = (n + kPrepThreads - 1) / kPrepThreads;
    cudaError_t e = cudaMemsetAsync(w.error, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    if (n == 0) return cudaMemsetAsync(n_out, 0, sizeof(int), s);
    const uint32_t tmask = static_cast<uint32_t>(w.table_slots - 1);
    if (voxel_size > 0.0) {
        e = cudaMemsetAsync(w.tkeys, 0xff, w.table_slots * sizeof(unsigned long long), s);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(w.tmin, 0xff, w.table_slots * sizeof(uint32_t), s);
        if (e != cudaSuccess) return e;
        prep_mark_kernel<<<blocks, kPrepThreads, 0, s>>>(xyz, n, n_dev, max_dist, voxel_size, w.tkeys, w.tmin, tmask, w.error);
    }
    prep_select_kernel<<<blocks, kPrepThreads, 0, s>>>(xyz, n, n_dev, max_dist, voxel_size, w.tkeys, w.tmin, tmask, w.keep, w.block_count);
    prep_offsets_kernel<<<1, 1024, 0, s>>>(w.block_count, blocks, w.block_offset, n_out);
    prep_scatter_kernel<<<blocks, kPrepThreads, 0, s>>>(xyz, aux, n, n_dev, w.keep, w.block_offset, xyz_out, aux_out, index_out);
    return cudaGetLastError();
}

}  // namespace elm
