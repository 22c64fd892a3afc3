// GPU map build: the three reference-defined passes that produce the static map, on the device, bit-identical to
// the host builder (host_map.cpp) and through it to the reference:
//   VoxelHashMap::AddPoints + AddPointWithSpacing   pcm_matching/src/voxel_hash_map.cpp:270-285, include/voxel_hash_map.hpp:106-113
//   VoxelHashMap::CalVoxelCovAll / CalVoxelCov      voxel_hash_map.hpp:114-148, 183-193
//   VoxelHashMap::CalPointCovAll / ProcessVoxelBlock voxel_hash_map.hpp:195-257 (self counted twice, Q5)
// AddPoints is order dependent (the first point of a voxel is always kept, a later one iff the voxel is below its cap and no
// kept point lies within sqrt(vs^2 / cap)), but voxels never interact and inside a voxel only the ARRIVAL order matters:
//   keys (truncation toward zero, :275) -> STABLE radix sort by (voxel key; arrival index as payload) -> run-length encode the
//   voxels -> one thread per voxel replays its points in arrival order through the spacing test -> scan of the kept counts ->
//   compaction into the canonical arrays (voxels ascending by key, arrival order inside).
// The sort / run-length / scan primitives are CUB (a plain library primitive, like cuBLAS for a plain GEMM); everything that
// carries the reference's semantics is written here.  This file is compiled with -fmad=false: every double operation is a plain
// IEEE multiply / add / divide / sqrt in the association order of the host builder (cov_math.hpp is the shared source), so the
// covariances agree with the host builder bit for bit, not just to rounding.
#include "map_build.cuh"

#include <cub/cub.cuh>

#include <cmath>
#include <cstring>

#include "cov_math.hpp"

namespace elm {

namespace {

#define MB_CUDA(call)                                                                                \
    do {                                                                                             \
        const cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) return std::string(#call) + ": " + cudaGetErrorString(e__);          \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(reinterpret_cast<void**>(&p), (n ? n : 1) * sizeof(T)); }
};

constexpr int kThreads = 256;
inline int blocks_for(size_t n) { return static_cast<int>((n + kThreads - 1) / kThreads); }

// insert key = static_cast<int>(p / voxel_size) per axis: truncation toward zero (vhm.cpp:275)
__global__ void __launch_bounds__(kThreads) keys_kernel(const float* __restrict__ xyz, size_t n, double vs, unsigned long long* __restrict__ keys,
                                                         uint32_t* __restrict__ idx, int* __restrict__ bad) {
    const size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (i >= n) return;
    const double qx = static_cast<double>(xyz[3 * i]) / vs, qy = static_cast<double>(xyz[3 * i + 1]) / vs, qz = static_cast<double>(xyz[3 * i + 2]) / vs;
    const double lim = static_cast<double>(kKeyBias - 2);
    idx[i] = static_cast<uint32_t>(i);
    if (!(fabs(qx) < lim && fabs(qy) < lim && fabs(qz) < lim)) { *bad = 1; keys[i] = 0; return; }
    keys[i] = pack_key(static_cast<int32_t>(qx), static_cast<int32_t>(qy), static_cast<int32_t>(qz));
}

// AddPointWithSpacing replayed per voxel in arrival order.  The kept points of voxel v are written to kept_xyz / kept_idx at the
// voxel's raw offset (they are also the list the later points are tested against).
__global__ void __launch_bounds__(kThreads) spacing_filter_kernel(const float* __restrict__ xyz, const uint32_t* __restrict__ idx_sorted,
                                                                   const uint32_t* __restrict__ rstart, const uint32_t* __restrict__ rcount, uint32_t nvox, int cap,
                                                                   double d2_limit, float* __restrict__ kept_xyz, uint32_t* __restrict__ kept_idx,
                                                                   uint32_t* __restrict__ kcount) {
    const uint32_t v = blockIdx.x * kThreads + threadIdx.x;
    if (v >= nvox) return;
    const size_t s = rstart[v];
    const uint32_t m = rcount[v];
    uint32_t kept = 0;
    for (uint32_t j = 0; j < m; ++j) {
        if (kept >= static_cast<uint32_t>(cap)) break;  // full: nothing later can enter (vhm.hpp:108)
        const uint32_t i = idx_sorted[s + j];
        const float fx = xyz[3 * static_cast<size_t>(i)], fy = xyz[3 * static_cast<size_t>(i) + 1], fz = xyz[3 * static_cast<size_t>(i) + 2];
        const double x = fx, y = fy, z = fz;
        bool ok = true;
        for (uint32_t k = 0; ok && k < kept; ++k) {  // (the first point of a voxel is always kept, vhm.cpp:281-283)
            const double dx = static_cast<double>(kept_xyz[3 * (s + k)]) - x, dy = static_cast<double>(kept_xyz[3 * (s + k) + 1]) - y,
                         dz = static_cast<double>(kept_xyz[3 * (s + k) + 2]) - z;
            if ((dx * dx + dy * dy) + dz * dz < d2_limit) ok = false;
        }
        if (ok) {
            kept_xyz[3 * (s + kept)] = fx; kept_xyz[3 * (s + kept) + 1] = fy; kept_xyz[3 * (s + kept) + 2] = fz;
            kept_idx[s + kept] = i;
            ++kept;
        }
    }
    kcount[v] = kept;
}

__global__ void __launch_bounds__(kThreads) compact_kernel(const float* __restrict__ kept_xyz, const uint32_t* __restrict__ kept_idx,
                                                            const uint32_t* __restrict__ rstart, const uint32_t* __restrict__ vstart, const uint32_t* __restrict__ kcount,
                                                            uint32_t nvox, float* __restrict__ pxyz, uint32_t* __restrict__ porig) {
    const uint32_t v = blockIdx.x * kThreads + threadIdx.x;
    if (v >= nvox) return;
    const size_t s = rstart[v], o = vstart[v];
    const uint32_t m = kcount[v];
    for (uint32_t k = 0; k < m; ++k) {
        pxyz[3 * (o + k)] = kept_xyz[3 * (s + k)]; pxyz[3 * (o + k) + 1] = kept_xyz[3 * (s + k) + 1]; pxyz[3 * (o + k) + 2] = kept_xyz[3 * (s + k) + 2];
        porig[o + k] = kept_idx[s + k];
    }
}

// CalVoxelCov (vhm.hpp:114-148): n == 1 -> (I, p); n >= 2 -> mean, sample covariance /(n - 1), regularised
__global__ void __launch_bounds__(kThreads) voxel_cov_kernel(const float* __restrict__ pxyz, const uint32_t* __restrict__ vstart, uint32_t nvox,
                                                              double* __restrict__ vmean, double* __restrict__ vcov) {
    const uint32_t v = blockIdx.x * kThreads + threadIdx.x;
    if (v >= nvox) return;
    const size_t s = vstart[v];
    const uint32_t cnt = vstart[v + 1] - vstart[v];
    double* m = vmean + 3 * static_cast<size_t>(v);
    double* c = vcov + 9 * static_cast<size_t>(v);
    if (cnt == 1) {
        m[0] = pxyz[3 * s]; m[1] = pxyz[3 * s + 1]; m[2] = pxyz[3 * s + 2];
        for (int i = 0; i < 9; ++i) c[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    double mean[3], cov[9];
    mean_cov_regularized_seq(cnt, [&](size_t i, double* p) { p[0] = pxyz[3 * (s + i)]; p[1] = pxyz[3 * (s + i) + 1]; p[2] = pxyz[3 * (s + i) + 2]; }, mean, cov,
                             static_cast<double*>(nullptr));
    for (int i = 0; i < 3; ++i) m[i] = mean[i];
    for (int i = 0; i < 9; ++i) c[i] = cov[i];
}

// directory lookup of a centre key (same 2-choice table the search kernels read, icp_device.cuh); -1: not present
__device__ int dir_find(const uint4* __restrict__ dslots, uint32_t bmask, uint64_t key) {
    uint32_t b1, b2;
    dir_buckets(key, bmask, b1, b2);
    const uint32_t klo = static_cast<uint32_t>(key), khi = static_cast<uint32_t>(key >> 32);
    const uint32_t cand[4] = {2 * b1, 2 * b1 + 1, 2 * b2, 2 * b2 + 1};
    for (int k = 0; k < 4; ++k) {
        const uint4 s = dslots[cand[k]];
        if (s.x == klo && s.y == khi) return static_cast<int>(cand[k]);
    }
    return -1;
}

// ProcessVoxelBlock (vhm.hpp:195-250) per stored point: neighbours = {self} U {stored points of the 27 voxels around
// FLOOR(p / vs) with d^2 <= r^2} in the reference's visit order (x outer, y, z inner; insertion order inside a voxel) — self is in
// that set too and therefore counted twice (Q5) — then mean, covariance /(n - 1), regularisation, normal.
__global__ void __launch_bounds__(kThreads) point_cov_kernel(const float* __restrict__ pxyz, size_t npts, const uint4* __restrict__ dslots,
                                                              const uint32_t* __restrict__ drows, uint32_t bmask, double vs, double r2,
                                                              double* __restrict__ pmean, double* __restrict__ pcov, double* __restrict__ pnormal) {
    const size_t p = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (p >= npts) return;
    const double x = pxyz[3 * p], y = pxyz[3 * p + 1], z = pxyz[3 * p + 2];
    const int32_t kx = static_cast<int32_t>(floor(x / vs)), ky = static_cast<int32_t>(floor(y / vs)), kz = static_cast<int32_t>(floor(z / vs));
    const int row = dir_find(dslots, bmask, pack_key(kx, ky, kz));
    // runs of the nine z-columns {first point, n(z-1) | n(z) << 10 | n(z+1) << 20}: the 27 voxels in visit order
    uint32_t first[9], counts[9];
    for (int c = 0; c < 9; ++c) {
        first[c] = 0; counts[c] = 0;
        if (row >= 0) { const size_t w = row_col_word(static_cast<size_t>(row), c); first[c] = drows[w]; counts[c] = drows[w + 1]; }
    }
    auto gap2 = [vs](double q0, int32_t c) {
        const double q = q0 / vs;
        const double lo = static_cast<double>(c <= 0 ? c - 1 : c), hi = static_cast<double>(c >= 0 ? c + 1 : c);
        const double g = cm_max(cm_max(lo - q, q - hi), 0.0) * vs * (1.0 - 1e-9);
        return g * g;
    };
    const double r2_skip = r2 * (1.0 + 1e-9);
    // visits self, then every neighbour within r, in order; called once per pass
    auto visit = [&](auto&& f) {
        const double self[3] = {x, y, z};
        f(self);
        for (int c = 0; c < 9; ++c) {
            const int i = kx + c / 3 - 1, j = ky + c % 3 - 1;
            uint32_t q = first[c];
            for (int dz = 0; dz < 3; ++dz) {
                const uint32_t n = (counts[c] >> (kDirCountBits * dz)) & kDirCountMask;
                const int k = kz + dz - 1;
                if (n && !(gap2(x, i) + gap2(y, j) + gap2(z, k) > r2_skip)) {
                    for (uint32_t t = q; t < q + n; ++t) {
                        const double px = pxyz[3 * static_cast<size_t>(t)], py = pxyz[3 * static_cast<size_t>(t) + 1], pz = pxyz[3 * static_cast<size_t>(t) + 2];
                        const double dx = px - x, dy = py - y, dz2 = pz - z;
                        if ((dx * dx + dy * dy) + dz2 * dz2 <= r2) { const double nb[3] = {px, py, pz}; f(nb); }
                    }
                }
                q += n;
            }
        }
    };
    // pass 1: count + sum; pass 2: covariance — the arithmetic of mean_cov_regularized_seq on a sequence that is generated twice
    double s[3] = {0, 0, 0};
    size_t n = 0;
    visit([&](const double* q) { s[0] += q[0]; s[1] += q[1]; s[2] += q[2]; ++n; });
    const double dn = static_cast<double>(n);
    double mean[3] = {s[0] / dn, s[1] / dn, s[2] / dn};
    double c9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    visit([&](const double* q) {
        const double d[3] = {q[0] - mean[0], q[1] - mean[1], q[2] - mean[2]};
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) c9[a * 3 + b] += d[a] * d[b];
    });
    for (int i = 0; i < 9; ++i) c9[i] /= (dn - 1.0);
    double cov[9], nrm[3];
    plane_regularize_hd(c9, cov, nrm);
    for (int i = 0; i < 3; ++i) { pmean[3 * p + i] = mean[i]; pnormal[3 * p + i] = nrm[i]; }
    for (int i = 0; i < 9; ++i) pcov[9 * p + i] = cov[i];
}

}  // namespace

std::string gpu_add_points(const float* xyz, size_t n, double voxel_size, int cap, GpuCanonicalMap& out) {
    out = GpuCanonicalMap{};
    if (n == 0) return "";
    if (n >= (1ull << 31)) return "more than 2^31 points";
    DevBuf<float> d_xyz, d_kept_xyz, d_pxyz;
    DevBuf<unsigned long long> d_keys, d_keys_sorted, d_ukeys;
    DevBuf<uint32_t> d_idx, d_idx_sorted, d_rcount, d_rstart, d_kept_idx, d_kcount, d_vstart, d_porig;
    DevBuf<int> d_flags;  // [0] bad key, [1] number of runs
    DevBuf<unsigned char> d_tmp;
    MB_CUDA(d_xyz.alloc(3 * n)); MB_CUDA(d_keys.alloc(n)); MB_CUDA(d_keys_sorted.alloc(n)); MB_CUDA(d_idx.alloc(n)); MB_CUDA(d_idx_sorted.alloc(n));
    MB_CUDA(d_flags.alloc(2));
    MB_CUDA(cudaMemcpy(d_xyz.p, xyz, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    MB_CUDA(cudaMemset(d_flags.p, 0, 2 * sizeof(int)));
    keys_kernel<<<blocks_for(n), kThreads>>>(d_xyz.p, n, voxel_size, d_keys.p, d_idx.p, d_flags.p);
    MB_CUDA(cudaGetLastError());
    // stable radix sort by the 63-bit key; the arrival index travels as the payload, so equal keys stay in arrival order
    size_t tmp_bytes = 0, need = 0;
    const int in = static_cast<int>(n);
    MB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, d_keys.p, d_keys_sorted.p, d_idx.p, d_idx_sorted.p, in, 0, 63));
    tmp_bytes = need;
    MB_CUDA(d_ukeys.alloc(n)); MB_CUDA(d_rcount.alloc(n + 1));
    MB_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, need, d_keys_sorted.p, d_ukeys.p, d_rcount.p, d_flags.p + 1, in));
    tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
    MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, d_rcount.p, d_rcount.p, in + 1));
    tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
    MB_CUDA(d_tmp.alloc(tmp_bytes));
    need = tmp_bytes;
    MB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, need, d_keys.p, d_keys_sorted.p, d_idx.p, d_idx_sorted.p, in, 0, 63));
    need = tmp_bytes;
    MB_CUDA(cub::DeviceRunLengthEncode::Encode(d_tmp.p, need, d_keys_sorted.p, d_ukeys.p, d_rcount.p, d_flags.p + 1, in));
    int flags[2] = {0, 0};
    MB_CUDA(cudaMemcpy(flags, d_flags.p, sizeof flags, cudaMemcpyDeviceToHost));
    if (flags[0]) return "map point outside +-2^20 voxels per axis (or not finite)";
    const uint32_t nvox = static_cast<uint32_t>(flags[1]);
    MB_CUDA(d_rstart.alloc(nvox + 1)); MB_CUDA(d_kcount.alloc(nvox + 1)); MB_CUDA(d_vstart.alloc(nvox + 1));
    need = tmp_bytes;
    MB_CUDA(cudaMemset(d_rcount.p + nvox, 0, sizeof(uint32_t)));
    MB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, need, d_rcount.p, d_rstart.p, static_cast<int>(nvox) + 1));
    // AddPointWithSpacing: sqrt(d2) < map_resolution <=> d2 < d2_limit, the smallest double whose correctly rounded square root
    // reaches map_resolution (the host builder's formulation of the same test)
    const double map_resolution = std::sqrt(voxel_size * voxel_size / cap);
    double d2_limit = map_resolution * map_resolution;
    while (std::sqrt(d2_limit) >= map_resolution) d2_limit = std::nextafter(d2_limit, 0.0);
    while (std::sqrt(d2_limit) < map_resolution) d2_limit = std::nextafter(d2_limit, HUGE_VAL);
    MB_CUDA(d_kept_xyz.alloc(3 * n)); MB_CUDA(d_kept_idx.alloc(n));
    spacing_filter_kernel<<<blocks_for(nvox), kThreads>>>(d_xyz.p, d_idx_sorted.p, d_rstart.p, d_rcount.p, nvox, cap, d2_limit, d_kept_xyz.p, d_kept_idx.p,
                                                          d_kcount.p);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaMemset(d_kcount.p + nvox, 0, sizeof(uint32_t)));
    need = tmp_bytes;
    MB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, need, d_kcount.p, d_vstart.p, static_cast<int>(nvox) + 1));
    uint32_t npts = 0;
    MB_CUDA(cudaMemcpy(&npts, d_vstart.p + nvox, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    MB_CUDA(d_pxyz.alloc(3 * static_cast<size_t>(npts))); MB_CUDA(d_porig.alloc(npts));
    compact_kernel<<<blocks_for(nvox), kThreads>>>(d_kept_xyz.p, d_kept_idx.p, d_rstart.p, d_vstart.p, d_kcount.p, nvox, d_pxyz.p, d_porig.p);
    MB_CUDA(cudaGetLastError());
    out.vkey.resize(nvox); out.vstart.resize(nvox + 1); out.pxyz.resize(3 * static_cast<size_t>(npts)); out.porig.resize(npts);
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "key width");
    MB_CUDA(cudaMemcpy(out.vkey.data(), d_ukeys.p, nvox * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    MB_CUDA(cudaMemcpy(out.vstart.data(), d_vstart.p, (static_cast<size_t>(nvox) + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    MB_CUDA(cudaMemcpy(out.pxyz.data(), d_pxyz.p, out.pxyz.size() * sizeof(float), cudaMemcpyDeviceToHost));
    MB_CUDA(cudaMemcpy(out.porig.data(), d_porig.p, out.porig.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return "";
}

std::string gpu_cal_voxel_cov(const std::vector<float>& pxyz, const std::vector<uint32_t>& vstart, std::vector<double>& vmean, std::vector<double>& vcov) {
    const size_t nvox = vstart.empty() ? 0 : vstart.size() - 1;
    vmean.assign(3 * nvox, 0.0); vcov.assign(9 * nvox, 0.0);
    if (nvox == 0) return "";
    DevBuf<float> d_pxyz;
    DevBuf<uint32_t> d_vstart;
    DevBuf<double> d_mean, d_cov;
    MB_CUDA(d_pxyz.alloc(pxyz.size())); MB_CUDA(d_vstart.alloc(vstart.size())); MB_CUDA(d_mean.alloc(3 * nvox)); MB_CUDA(d_cov.alloc(9 * nvox));
    MB_CUDA(cudaMemcpy(d_pxyz.p, pxyz.data(), pxyz.size() * sizeof(float), cudaMemcpyHostToDevice));
    MB_CUDA(cudaMemcpy(d_vstart.p, vstart.data(), vstart.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    voxel_cov_kernel<<<blocks_for(nvox), kThreads>>>(d_pxyz.p, d_vstart.p, static_cast<uint32_t>(nvox), d_mean.p, d_cov.p);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaMemcpy(vmean.data(), d_mean.p, vmean.size() * sizeof(double), cudaMemcpyDeviceToHost));
    MB_CUDA(cudaMemcpy(vcov.data(), d_cov.p, vcov.size() * sizeof(double), cudaMemcpyDeviceToHost));
    return "";
}

std::string gpu_cal_point_cov(const std::vector<float>& pxyz, const uint4* d_dslots, const uint32_t* d_drows, uint32_t bmask, double voxel_size,
                              double search_dist, std::vector<double>& pmean, std::vector<double>& pcov, std::vector<double>& pnormal) {
    const size_t np = pxyz.size() / 3;
    pmean.assign(3 * np, 0.0); pcov.assign(9 * np, 0.0); pnormal.assign(3 * np, 0.0);
    if (np == 0) return "";
    DevBuf<float> d_pxyz;
    DevBuf<double> d_mean, d_cov, d_nrm;
    MB_CUDA(d_pxyz.alloc(pxyz.size())); MB_CUDA(d_mean.alloc(3 * np)); MB_CUDA(d_cov.alloc(9 * np)); MB_CUDA(d_nrm.alloc(3 * np));
    MB_CUDA(cudaMemcpy(d_pxyz.p, pxyz.data(), pxyz.size() * sizeof(float), cudaMemcpyHostToDevice));
    point_cov_kernel<<<blocks_for(np), kThreads>>>(d_pxyz.p, np, d_dslots, d_drows, bmask, voxel_size, search_dist * search_dist, d_mean.p, d_cov.p, d_nrm.p);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaMemcpy(pmean.data(), d_mean.p, pmean.size() * sizeof(double), cudaMemcpyDeviceToHost));
    MB_CUDA(cudaMemcpy(pcov.data(), d_cov.p, pcov.size() * sizeof(double), cudaMemcpyDeviceToHost));
    MB_CUDA(cudaMemcpy(pnormal.data(), d_nrm.p, pnormal.size() * sizeof(double), cudaMemcpyDeviceToHost));
    return "";
}

}  // namespace elm
