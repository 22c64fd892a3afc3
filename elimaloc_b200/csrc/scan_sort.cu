// Spatial binning of the scan, once per RunRegister call.
//
// The reference hands RunRegister the scan in unordered_map iteration order (VoxelDownsample, voxel_hash_map.hpp:
// 260-283), i.e. spatially random: consecutive queries touch unrelated parts of the map and every probe / point run is
// a cold miss.  On the GPU a counting sort of the scan by the voxel it falls into under the INITIAL guess costs a few
// microseconds and makes the 256 queries of a search tile neighbours in the map, so their probes and point runs hit
// L1/L2.  Only the SEARCH runs in binned order (match[] is written back under the original index); the accumulation
// keeps the caller's order, so the sums — and therefore every result — are independent of how ties inside a bin fall.
#include "icp_kernels.cuh"
#include "voxel_key.hpp"

namespace elm {

namespace {

__device__ __forceinline__ uint32_t bin_of(const double* T, float sx, float sy, float sz, double inv_vs, int bits) {
    // plain fp64 here: the bin only steers locality, it never influences a result
    const double px = T[0] * sx + T[1] * sy + T[2] * sz + T[3];
    const double py = T[4] * sx + T[5] * sy + T[6] * sz + T[7];
    const double pz = T[8] * sx + T[9] * sy + T[10] * sz + T[11];
    const uint32_t m = (1u << bits) - 1u;
    const uint32_t kx = static_cast<uint32_t>(static_cast<long long>(floor(px * inv_vs))) & m;
    const uint32_t ky = static_cast<uint32_t>(static_cast<long long>(floor(py * inv_vs))) & m;
    const uint32_t kz = static_cast<uint32_t>(static_cast<long long>(floor(pz * inv_vs))) & m;
    return (kx << (2 * bits)) | (ky << bits) | kz;  // x-major like the map's own point order
}

__global__ void __launch_bounds__(256) bin_count_kernel(const float* __restrict__ scan, int n, Pose16 T, double inv_vs, int bits,
                                                        uint32_t* __restrict__ bin, uint32_t* __restrict__ hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = bin_of(T.m, scan[3 * static_cast<size_t>(i)], scan[3 * static_cast<size_t>(i) + 1], scan[3 * static_cast<size_t>(i) + 2], inv_vs, bits);
    bin[i] = b;
    atomicAdd(hist + b, 1u);
}

// exclusive scan of hist[nbins] in place, one block of 1024 threads (nbins <= 2^18)
__global__ void __launch_bounds__(1024) bin_scan_kernel(uint32_t* __restrict__ hist, int nbins) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nbins; base += 4096) {
        // 4 consecutive bins per thread
        const int i0 = base + tid * 4;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i0 + k < nbins) ? hist[i0 + k] : 0u;
        const uint32_t tsum = v[0] + v[1] + v[2] + v[3];
        uint32_t x = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;
        }
        __syncthreads();
        uint32_t excl = s_carry + (warp ? s_warp[warp - 1] : 0u) + (x - tsum);
#pragma unroll
        for (int k = 0; k < 4; ++k) { if (i0 + k < nbins) hist[i0 + k] = excl; excl += v[k]; }
        __syncthreads();
        if (tid == 1023) s_carry += s_warp[31];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) bin_scatter_kernel(const float* __restrict__ scan, int n, const uint32_t* __restrict__ bin,
                                                          uint32_t* __restrict__ cursor, float* __restrict__ sorted, int* __restrict__ orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pos = atomicAdd(cursor + bin[i], 1u);
    sorted[3 * static_cast<size_t>(pos)] = scan[3 * static_cast<size_t>(i)];
    sorted[3 * static_cast<size_t>(pos) + 1] = scan[3 * static_cast<size_t>(i) + 1];
    sorted[3 * static_cast<size_t>(pos) + 2] = scan[3 * static_cast<size_t>(i) + 2];
    orig[pos] = i;
}

}  // namespace

int scan_bin_bits(int n) {
    int bits = 3;
    while (bits < 6 && (1 << (3 * bits)) < 2 * n) ++bits;
    return bits;
}

// scan[n] -> sorted[n] (same points, binned order) + orig[n] (original index of each sorted point).
// scratch: bin[n] and hist[1 << (3 * bits)] (zeroed here).  Three small launches + one memset.
cudaError_t launch_scan_binning(const float* scan, int n, const double T[16], double voxel_size, uint32_t* bin, uint32_t* hist, float* sorted,
                                int* orig, cudaStream_t s) {
    const int bits = scan_bin_bits(n);
    const int nbins = 1 << (3 * bits);
    Pose16 p;
    for (int i = 0; i < 16; ++i) p.m[i] = T[i];
    cudaError_t e = cudaMemsetAsync(hist, 0, static_cast<size_t>(nbins) * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    const int blocks = (n + 255) / 256;
    bin_count_kernel<<<blocks, 256, 0, s>>>(scan, n, p, 1.0 / voxel_size, bits, bin, hist);
    bin_scan_kernel<<<1, 1024, 0, s>>>(hist, nbins);
    bin_scatter_kernel<<<blocks, 256, 0, s>>>(scan, n, bin, hist, sorted, orig);
    return cudaGetLastError();
}

}  // namespace elm
