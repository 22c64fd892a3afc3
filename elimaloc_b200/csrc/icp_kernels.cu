// Hand-written sm_100a kernels of the ICP hot loop.
//
//   icp_begin_kernel        state <- initial guess (+ inverse)                     reg.cpp:298,24,79
//   icp_points_kernel<M>    P2P / GICP : TransformPoints + GetCorrespondencePoints + AlignCloudsLocal{,PointCov}
//                           accumulation, fused                                     reg.hpp:136-148, vhm.cpp:31-88,
//                                                                                   reg.cpp:28-51 / 85-132
//   icp_voxels_kernel<M>    VGICP / AVGICP : TransformPoints + GetCorrespondences{Cov,AllCov} +
//                           AlignCloudsLocalVoxelCov accumulation, fused            vhm.cpp:90-206, reg.cpp:171-208
//   icp_reduce_kernel       per-block partials -> 30 accumulators (fixed order)     (multi-GPU: followed by ncclAllReduce)
//   icp_solve_kernel        overlap gate, LM-damped LDLT solve, exp map, pose update, termination test
//                                                                                   reg.cpp:349-356, 53-65, 136-151, 378-387
//   icp_match_kernel        correspondence dump for the parity tests
//
// Exactness: the transformed scan point, its voxel key and every candidate distance are computed in fp64 with
// explicit round-to-nearest mul/add (never contracted into FMA) in the same association order as the CPU reference,
// so the nearest-neighbour choice — including its first-in-visit-order tie-break — is bit-identical to the reference.
// The accumulation that follows is plain fp64 (FMA allowed); it is compared with a tolerance.
#include "icp_kernels.cuh"

namespace elm {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr int kKeyBits = 21;
constexpr int kKeyBias = 1 << (kKeyBits - 1);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- exact helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double sq3_exact(double dx, double dy, double dz) {  // (dx^2 + dy^2) + dz^2, no FMA
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
// row r of  T * (x, y, z, 1):  ((T0 x + T1 y) + T2 z) + T3, no FMA  (reg.hpp:142-145)
__device__ __forceinline__ double row_apply_exact(const double* T, int r, double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4 * r], x), __dmul_rn(T[4 * r + 1], y)), __dmul_rn(T[4 * r + 2], z)), T[4 * r + 3]);
}
__device__ __forceinline__ int voxel_floor(double p, double vs) {  // PointToVoxel, vhm.hpp:176-180
    const double q = floor(__ddiv_rn(p, vs));
    // saturate far outside the table's key range instead of the reference's undefined int overflow
    return (q >= 2.0e9) ? 2000000000 : ((q <= -2.0e9) ? -2000000000 : static_cast<int>(q));
}
__device__ __forceinline__ bool key_ok(int k) { return k >= -kKeyBias && k < kKeyBias; }
__device__ __forceinline__ uint64_t pack_key(int x, int y, int z) {
    return (static_cast<uint64_t>(static_cast<uint32_t>(x + kKeyBias)) << (2 * kKeyBits)) |
           (static_cast<uint64_t>(static_cast<uint32_t>(y + kKeyBias)) << kKeyBits) |
           static_cast<uint64_t>(static_cast<uint32_t>(z + kKeyBias));
}
__device__ __forceinline__ uint32_t hash_key(uint64_t k) {  // murmur3 finaliser, same as host_map.hpp
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return static_cast<uint32_t>(k);
}
// Linear probe; returns the slot index or -1.  start/count filled on a hit.
__device__ __forceinline__ int probe(const uint4* __restrict__ slots, uint32_t mask, uint64_t key, uint32_t& start, uint32_t& count) {
    uint32_t h = hash_key(key) & mask;
    const uint32_t klo = static_cast<uint32_t>(key), khi = static_cast<uint32_t>(key >> 32);
    for (;;) {
        const uint4 s = __ldg(slots + h);
        if (s.x == klo && s.y == khi) { start = s.z; count = s.w; return static_cast<int>(h); }
        if ((s.x & s.y) == 0xffffffffu) return -1;
        h = (h + 1) & mask;
    }
}

// ---- warp argmin over (fp64 distance >= 0, visit order) --------------------------------------------------------
// Returns the lane holding the smallest (d2, ord) pair, or -1 when no lane has a candidate (ord == 0xffffffff).
__device__ __forceinline__ int warp_argmin(double d2, uint32_t ord) {
    const uint32_t hi = static_cast<uint32_t>(__double2hiint(d2));
    const uint32_t lo = static_cast<uint32_t>(__double2loint(d2));
    const uint32_t mhi = __reduce_min_sync(kFull, hi);
    const bool a = (hi == mhi);
    const uint32_t mlo = __reduce_min_sync(kFull, a ? lo : 0xffffffffu);
    const bool b = a && (lo == mlo);
    const uint32_t mord = __reduce_min_sync(kFull, b ? ord : 0xffffffffu);
    if (mord == 0xffffffffu) return -1;
    const uint32_t who = __ballot_sync(kFull, b && ord == mord);
    return __ffs(who) - 1;
}

// Nearest stored map point of the 27 voxels around key (kx,ky,kz), warp-cooperative.
// Visit order of the reference: voxels x-outer / y / z-inner (vhm.cpp:234-240), insertion order inside a voxel, strict <
// (vhm.cpp:45).  Lane L < 27 probes voxel L; the three z-voxels of a column are one contiguous run of `pts`.
// Returns the winning point index (warp-uniform) or -1.
__device__ __forceinline__ int nearest_point_27(const MapView& map, double px, double py, double pz, int kx, int ky, int kz, int lane) {
    uint32_t start = 0, count = 0;
    if (lane < 27) {
        const int x = kx + lane / 9 - 1, y = ky + (lane / 3) % 3 - 1, z = kz + lane % 3 - 1;
        if (key_ok(x) && key_ok(y) && key_ok(z)) {
            if (probe(map.slots, map.mask, pack_key(x, y, z), start, count) < 0) count = 0;
        }
    }
    // column c = lanes 3c..3c+2
    const uint32_t c1 = __shfl_down_sync(kFull, count, 1), c2 = __shfl_down_sync(kFull, count, 2);
    const uint32_t s1 = __shfl_down_sync(kFull, start, 1), s2 = __shfl_down_sync(kFull, start, 2);
    const uint32_t runlen = count + c1 + c2;
    const uint32_t runstart = count ? start : (c1 ? s1 : s2);

    double best = 1.7976931348623157e308;
    uint32_t bord = 0xffffffffu;
    uint32_t bidx = 0;
    uint32_t rs[9], rl[9];
    float4 m[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        rs[c] = __shfl_sync(kFull, runstart, 3 * c);
        rl[c] = __shfl_sync(kFull, runlen, 3 * c);
    }
#pragma unroll
    for (int c = 0; c < 9; ++c)
        if (static_cast<uint32_t>(lane) < rl[c]) m[c] = __ldg(map.pts + rs[c] + lane);
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        if (static_cast<uint32_t>(lane) < rl[c]) {
            const double d2 = sq3_exact(static_cast<double>(m[c].x) - px, static_cast<double>(m[c].y) - py, static_cast<double>(m[c].z) - pz);
            const uint32_t ord = (static_cast<uint32_t>(c) << 20) | static_cast<uint32_t>(lane);
            if (d2 < best || (d2 == best && ord < bord)) { best = d2; bord = ord; bidx = rs[c] + lane; }
        }
    }
    // runs longer than a warp (more than 32 stored points in one column)
#pragma unroll 1
    for (int c = 0; c < 9; ++c) {
        for (uint32_t o = lane + 32; o < rl[c]; o += 32) {
            const float4 q = __ldg(map.pts + rs[c] + o);
            const double d2 = sq3_exact(static_cast<double>(q.x) - px, static_cast<double>(q.y) - py, static_cast<double>(q.z) - pz);
            const uint32_t ord = (static_cast<uint32_t>(c) << 20) | o;
            if (d2 < best || (d2 == best && ord < bord)) { best = d2; bord = ord; bidx = rs[c] + o; }
        }
    }
    const int wl = warp_argmin(best, bord);
    if (wl < 0) return -1;
    return static_cast<int>(__shfl_sync(kFull, bidx, wl));
}

// Nearest voxel MEAN of the 27 voxels (VGICP, vhm.cpp:90-151).  Returns the winning slot index or -1.
__device__ __forceinline__ int nearest_mean_27(const MapView& map, double px, double py, double pz, int kx, int ky, int kz, int lane) {
    double d2 = 1.7976931348623157e308;
    uint32_t ord = 0xffffffffu;
    int slot = -1;
    if (lane < 27) {
        const int x = kx + lane / 9 - 1, y = ky + (lane / 3) % 3 - 1, z = kz + lane % 3 - 1;
        if (key_ok(x) && key_ok(y) && key_ok(z)) {
            uint32_t st, cnt;
            slot = probe(map.slots, map.mask, pack_key(x, y, z), st, cnt);
            if (slot >= 0) {
                const double4 vm = map.vslots[slot];
                d2 = sq3_exact(vm.y - px, vm.z - py, vm.w - pz);
                ord = lane;
            }
        }
    }
    const int wl = warp_argmin(d2, ord);
    if (wl < 0) return -1;
    return __shfl_sync(kFull, slot, wl);
}

// ---- accumulation ---------------------------------------------------------------------------------------------------
// P2P keeps 18 structured sums (J = [I | -skew(s)] makes most of JtJ redundant); the others keep all 29.
//   P2P layout:  0 W=sum w | 1..3 sum w s | 4..9 (33,34,35,44,45,55) of sum w(|s|^2 I - s s^T) | 10..12 sum w r
//                13..15 sum w (s x r) | 16 residual | 17 count
template <int METHOD> struct AccSize { static constexpr int value = 29; };
template <> struct AccSize<0> { static constexpr int value = 18; };

__device__ __forceinline__ void acc_p2p(double* a, double sx, double sy, double sz, double rx, double ry, double rz, double th) {
    const double r2 = rx * rx + ry * ry + rz * rz;
    const double den = th + r2;
    const double w = (th * th) / (den * den);  // reg.cpp:44
    a[0] += w;
    a[1] += w * sx; a[2] += w * sy; a[3] += w * sz;
    a[4] += w * (sy * sy + sz * sz); a[5] -= w * sx * sy; a[6] -= w * sx * sz;
    a[7] += w * (sx * sx + sz * sz); a[8] -= w * sy * sz; a[9] += w * (sx * sx + sy * sy);
    a[10] += w * rx; a[11] += w * ry; a[12] += w * rz;
    a[13] += w * (sy * rz - sz * ry); a[14] += w * (sz * rx - sx * rz); a[15] += w * (sx * ry - sy * rx);
    a[16] += sqrt(r2);  // reg.cpp:50
    a[17] += 1.0;
}

// expand the P2P sums into the canonical 29 (upper JtJ row-major, Jtr, residual, count)
__device__ __forceinline__ void expand_p2p(const double* a, double* o) {
    for (int i = 0; i < 29; ++i) o[i] = 0.0;
    o[0] = a[0]; o[6] = a[0]; o[11] = a[0];            // I block
    o[4] = a[3]; o[5] = -a[2];                         // -skew(B): (0,4)=bz (0,5)=-by
    o[8] = -a[3]; o[10] = a[1];                        // (1,3)=-bz (1,5)=bx
    o[12] = a[2]; o[13] = -a[1];                       // (2,3)=by (2,4)=-bx
    o[15] = a[4]; o[16] = a[5]; o[17] = a[6]; o[18] = a[7]; o[19] = a[8]; o[20] = a[9];
    o[21] = a[10]; o[22] = a[11]; o[23] = a[12]; o[24] = a[13]; o[25] = a[14]; o[26] = a[15];
    o[27] = a[16]; o[28] = a[17];
}

// JtJ += w J^T M J, Jtr += w J^T M r with J = [I | A], A = -skew(s)   (reg.cpp:124-125, 204-205)
__device__ __forceinline__ void acc_mahalanobis(double* a, const double* M, double sx, double sy, double sz, double rx, double ry, double rz, double w) {
    const double A[3][3] = {{0.0, sz, -sy}, {-sz, 0.0, sx}, {sy, -sx, 0.0}};
    double MA[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) MA[i][j] = M[3 * i] * A[0][j] + M[3 * i + 1] * A[1][j] + M[3 * i + 2] * A[2][j];
    const double Mr[3] = {M[0] * rx + M[1] * ry + M[2] * rz, M[3] * rx + M[4] * ry + M[5] * rz, M[6] * rx + M[7] * ry + M[8] * rz};
    // rows 0..2: [M | MA]
    a[0] += w * M[0]; a[1] += w * M[1]; a[2] += w * M[2]; a[3] += w * MA[0][0]; a[4] += w * MA[0][1]; a[5] += w * MA[0][2];
    a[6] += w * M[4]; a[7] += w * M[5]; a[8] += w * MA[1][0]; a[9] += w * MA[1][1]; a[10] += w * MA[1][2];
    a[11] += w * M[8]; a[12] += w * MA[2][0]; a[13] += w * MA[2][1]; a[14] += w * MA[2][2];
    // rows 3..5: A^T M A (upper)
    int k = 15;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) { a[k] += w * (A[0][i] * MA[0][j] + A[1][i] * MA[1][j] + A[2][i] * MA[2][j]); ++k; }
    a[21] += w * Mr[0]; a[22] += w * Mr[1]; a[23] += w * Mr[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) a[24 + i] += w * (A[0][i] * Mr[0] + A[1][i] * Mr[1] + A[2][i] * Mr[2]);
}

// M = (Rinv C Rinv^T)^-1   (reg.cpp:107,113 / 187,191), cofactor inverse
__device__ __forceinline__ void mahalanobis_local(const double* Rinv, const double* C, double* M) {
    double RC[9], S[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) RC[3 * i + j] = Rinv[3 * i] * C[j] + Rinv[3 * i + 1] * C[3 + j] + Rinv[3 * i + 2] * C[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) S[3 * i + j] = RC[3 * i] * Rinv[3 * j] + RC[3 * i + 1] * Rinv[3 * j + 1] + RC[3 * i + 2] * Rinv[3 * j + 2];
    const double c00 = S[4] * S[8] - S[5] * S[7], c01 = S[2] * S[7] - S[1] * S[8], c02 = S[1] * S[5] - S[2] * S[4];
    const double c10 = S[5] * S[6] - S[3] * S[8], c11 = S[0] * S[8] - S[2] * S[6], c12 = S[2] * S[3] - S[0] * S[5];
    const double c20 = S[3] * S[7] - S[4] * S[6], c21 = S[1] * S[6] - S[0] * S[7], c22 = S[0] * S[4] - S[1] * S[3];
    const double inv = 1.0 / (S[0] * c00 + S[1] * c10 + S[2] * c20);
    M[0] = c00 * inv; M[1] = c01 * inv; M[2] = c02 * inv;
    M[3] = c10 * inv; M[4] = c11 * inv; M[5] = c12 * inv;
    M[6] = c20 * inv; M[7] = c21 * inv; M[8] = c22 * inv;
}

// block reduction of per-lane accumulators -> partials[block][kAcc], fixed order
template <int NACC, bool IS_P2P>
__device__ __forceinline__ void block_reduce_store(double* acc, double (*s_red)[kAcc], double* __restrict__ partials, int n_local) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        acc[k] = v;
    }
    if (lane == 0) {
        if (IS_P2P) {
            double o[29];
            expand_p2p(acc, o);
            for (int k = 0; k < 29; ++k) s_red[warp][k] = o[k];
        } else {
            for (int k = 0; k < 29; ++k) s_red[warp][k] = acc[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < kAcc) {
        double v = 0.0;
        if (threadIdx.x < 29) {
            for (int w = 0; w < kIcpWarps; ++w) v += s_red[w][threadIdx.x];
        } else if (threadIdx.x == kIdxNtotal) {
            v = (blockIdx.x == 0) ? static_cast<double>(n_local) : 0.0;
        }
        partials[blockIdx.x * kAcc + threadIdx.x] = v;
    }
}

}  // namespace

// ======================================================================================================================
// P2P / GICP
// ======================================================================================================================
template <int METHOD>
__global__ void __launch_bounds__(kIcpThreads, 2)
icp_points_kernel(MapView map, const float* __restrict__ scan, IcpParams prm, const IcpState* __restrict__ st, double* __restrict__ partials) {
    constexpr int NACC = AccSize<METHOD>::value;
    __shared__ __align__(16) float s_tile[2][kIcpWarps * 32 * 3];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_T[16], s_Tinv[16], s_Rinv[9];
    __shared__ double s_qp[kIcpWarps][3][32];
    __shared__ int s_qk[kIcpWarps][3][32];
    __shared__ double s_red[kIcpWarps][kAcc];

    if (st->done) return;  // loop already left (termination / overlap failure): the solve kernel skips too
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) { s_T[tid] = st->T[tid]; s_Tinv[tid] = st->Tinv[tid]; }
    if (tid < 9) s_Rinv[tid] = st->Rinv[tid];

    const int B = prm.queries_per_warp;
    const int tile_pts = kIcpWarps * B;
    const int ntiles = (prm.n + tile_pts - 1) / tile_pts;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(scan) & 15) == 0;
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
    __syncthreads();

    // scan tile -> shared memory through the TMA bulk-copy engine (packed xyz, 12 B/point)
    auto tile_count = [&](int t) { return min(tile_pts, prm.n - t * tile_pts); };
    auto tile_tma_ok = [&](int t) { return base_aligned && ((tile_count(t) * 12) & 15) == 0; };
    auto issue = [&](int t, int buf) {
        const uint32_t bytes = tile_count(t) * 12;
        mbar_expect_tx(&s_bar[buf], bytes);
        tma_load_1d(s_tile[buf], scan + static_cast<size_t>(t) * tile_pts * 3, bytes, &s_bar[buf]);
    };

    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;

    int tile = blockIdx.x, buf = 0;
    uint32_t phase[2] = {0, 0};
    if (tile < ntiles && tid == 0 && tile_tma_ok(tile)) issue(tile, 0);
    for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const int next = tile + gridDim.x;
        if (tid == 0 && next < ntiles && tile_tma_ok(next)) issue(next, buf ^ 1);  // prefetch the next tile
        if (tile_tma_ok(tile)) {
            mbar_wait(&s_bar[buf], phase[buf]);
            phase[buf] ^= 1;
        } else {  // ragged last tile / unaligned base: plain cooperative copy
            const int nf = tile_count(tile) * 3;
            for (int i = tid; i < nf; i += kIcpThreads) s_tile[buf][i] = scan[static_cast<size_t>(tile) * tile_pts * 3 + i];
            __syncthreads();
        }
        const int cnt = tile_count(tile) - warp * B;  // queries of this warp in this tile (may be <= 0)
        const int nq = min(B, cnt);
        // stage 1: each lane transforms its own query point, publishes it to the warp
        double sx = 0, sy = 0, sz = 0, px = 0, py = 0, pz = 0;
        if (lane < nq) {
            const float* sp = &s_tile[buf][(warp * B + lane) * 3];
            sx = sp[0]; sy = sp[1]; sz = sp[2];
            px = row_apply_exact(s_T, 0, sx, sy, sz);
            py = row_apply_exact(s_T, 1, sx, sy, sz);
            pz = row_apply_exact(s_T, 2, sx, sy, sz);
            s_qp[warp][0][lane] = px; s_qp[warp][1][lane] = py; s_qp[warp][2][lane] = pz;
            s_qk[warp][0][lane] = voxel_floor(px, map.voxel_size);
            s_qk[warp][1][lane] = voxel_floor(py, map.voxel_size);
            s_qk[warp][2][lane] = voxel_floor(pz, map.voxel_size);
        }
        __syncwarp();
        // stage 2: the warp searches the queries one after another
        int my_idx = -1;
        for (int q = 0; q < nq; ++q) {
            const int w = nearest_point_27(map, s_qp[warp][0][q], s_qp[warp][1][q], s_qp[warp][2][q], s_qk[warp][0][q],
                                           s_qk[warp][1][q], s_qk[warp][2][q], lane);
            if (lane == q) my_idx = w;
        }
        // stage 3: each lane linearises its own correspondence
        if (lane < nq) {
            double tx = 0.0, ty = 0.0, tz = 0.0;  // default-constructed neighbour at the origin (Q2, vhm.cpp:37)
            if (my_idx >= 0) { const float4 t = __ldg(map.pts + my_idx); tx = t.x; ty = t.y; tz = t.z; }
            const double d2 = sq3_exact(tx - px, ty - py, tz - pz);
            if (d2 < prm.max_dist2) {  // vhm.cpp:66
                if (METHOD == 0) {
                    const double lx = s_Tinv[0] * tx + s_Tinv[1] * ty + s_Tinv[2] * tz + s_Tinv[3];
                    const double ly = s_Tinv[4] * tx + s_Tinv[5] * ty + s_Tinv[6] * tz + s_Tinv[7];
                    const double lz = s_Tinv[8] * tx + s_Tinv[9] * ty + s_Tinv[10] * tz + s_Tinv[11];
                    acc_p2p(acc, sx, sy, sz, lx - sx, ly - sy, lz - sz, prm.th);
                } else {
                    double mean[3] = {0.0, 0.0, 0.0}, C[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, nrm[3] = {1.0, 0.0, 0.0};
                    if (my_idx >= 0) {
                        const double2* r = reinterpret_cast<const double2*>(map.prec + static_cast<size_t>(my_idx) * 16);
                        const double2 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3), r4 = __ldg(r + 4),
                                      r5 = __ldg(r + 5), r6 = __ldg(r + 6), r7 = __ldg(r + 7);
                        mean[0] = r0.x; mean[1] = r0.y; mean[2] = r1.x;
                        C[0] = r1.y; C[1] = r2.x; C[2] = r2.y; C[3] = r3.x; C[4] = r3.y; C[5] = r4.x; C[6] = r4.y; C[7] = r5.x; C[8] = r5.y;
                        nrm[0] = r6.x; nrm[1] = r6.y; nrm[2] = r7.x;
                    }
                    // residual to the neighbourhood MEAN, not the matched point (Q4, reg.cpp:97-101)
                    const double lx = s_Tinv[0] * mean[0] + s_Tinv[1] * mean[1] + s_Tinv[2] * mean[2] + s_Tinv[3];
                    const double ly = s_Tinv[4] * mean[0] + s_Tinv[5] * mean[1] + s_Tinv[6] * mean[2] + s_Tinv[7];
                    const double lz = s_Tinv[8] * mean[0] + s_Tinv[9] * mean[1] + s_Tinv[10] * mean[2] + s_Tinv[11];
                    const double rx = lx - sx, ry = ly - sy, rz = lz - sz;
                    double M[9];
                    mahalanobis_local(s_Rinv, C, M);
                    const double r2 = rx * rx + ry * ry + rz * rz;
                    const double den = prm.th + r2;
                    const double w = (prm.th * prm.th) / (den * den) * 0.8 + 0.2;  // reg.cpp:121
                    acc_mahalanobis(acc, M, sx, sy, sz, rx, ry, rz, w);
                    // point-to-plane fitness term (reg.cpp:94-95,128-131)
                    double nx = s_Rinv[0] * nrm[0] + s_Rinv[1] * nrm[1] + s_Rinv[2] * nrm[2];
                    double ny = s_Rinv[3] * nrm[0] + s_Rinv[4] * nrm[1] + s_Rinv[5] * nrm[2];
                    double nz = s_Rinv[6] * nrm[0] + s_Rinv[7] * nrm[1] + s_Rinv[8] * nrm[2];
                    const double nn = nx * nx + ny * ny + nz * nz;
                    if (nn > 0.0) { const double il = 1.0 / sqrt(nn); nx *= il; ny *= il; nz *= il; }
                    acc[27] += fabs(rx * nx + ry * ny + rz * nz);
                    acc[28] += 1.0;
                }
            }
        }
        __syncthreads();  // every warp is done with s_tile[buf] before it is refilled
    }
    block_reduce_store<NACC, METHOD == 0>(acc, s_red, partials, prm.n);
}

// ======================================================================================================================
// VGICP / AVGICP
// ======================================================================================================================
template <int METHOD>
__global__ void __launch_bounds__(kIcpThreads, 2)
icp_voxels_kernel(MapView map, const float* __restrict__ scan, IcpParams prm, const IcpState* __restrict__ st, double* __restrict__ partials) {
    __shared__ double s_T[16], s_Tinv[16], s_Rinv[9];
    __shared__ double s_red[kIcpWarps][kAcc];
    if (st->done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) { s_T[tid] = st->T[tid]; s_Tinv[tid] = st->Tinv[tid]; }
    if (tid < 9) s_Rinv[tid] = st->Rinv[tid];
    __syncthreads();

    double acc[29];
#pragma unroll
    for (int k = 0; k < 29; ++k) acc[k] = 0.0;

    // one (scan point, voxel) pair -> accumulators   (reg.cpp:171-208)
    auto linearize_pair = [&](double sx, double sy, double sz, int slot, double mx, double my, double mz) {
        double C[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (slot >= 0) {
            const double2* r = reinterpret_cast<const double2*>(map.vcov + static_cast<size_t>(slot) * 12);
            const double2 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
            const double c8 = __ldg(map.vcov + static_cast<size_t>(slot) * 12 + 8);
            C[0] = r0.x; C[1] = r0.y; C[2] = r1.x; C[3] = r1.y; C[4] = r2.x; C[5] = r2.y; C[6] = r3.x; C[7] = r3.y; C[8] = c8;
        }
        const double lx = s_Tinv[0] * mx + s_Tinv[1] * my + s_Tinv[2] * mz + s_Tinv[3];
        const double ly = s_Tinv[4] * mx + s_Tinv[5] * my + s_Tinv[6] * mz + s_Tinv[7];
        const double lz = s_Tinv[8] * mx + s_Tinv[9] * my + s_Tinv[10] * mz + s_Tinv[11];
        const double rx = lx - sx, ry = ly - sy, rz = lz - sz;
        const double r2 = rx * rx + ry * ry + rz * rz;
        const double den = prm.th + r2;
        const double w = (prm.th * prm.th) / (den * den);  // reg.cpp:199
        acc[28] += 1.0;                                      // the pair counts in the denominator either way (Q7)
        if (w < 0.01) return;                                // reg.cpp:201
        double M[9];
        mahalanobis_local(s_Rinv, C, M);
        acc_mahalanobis(acc, M, sx, sy, sz, rx, ry, rz, w);
        acc[27] += sqrt(r2);  // reg.cpp:207
    };

    if (METHOD == 2) {
        // VGICP: one warp searches 32 consecutive scan points, lane q then linearises point q
        const int nbatch = (prm.n + 31) / 32;
        for (int b = blockIdx.x * kIcpWarps + warp; b < nbatch; b += gridDim.x * kIcpWarps) {
            const int i = b * 32 + lane;
            double sx = 0, sy = 0, sz = 0, px = 0, py = 0, pz = 0;
            int kx = 0, ky = 0, kz = 0;
            if (i < prm.n) {
                sx = scan[3 * static_cast<size_t>(i)]; sy = scan[3 * static_cast<size_t>(i) + 1]; sz = scan[3 * static_cast<size_t>(i) + 2];
                px = row_apply_exact(s_T, 0, sx, sy, sz); py = row_apply_exact(s_T, 1, sx, sy, sz); pz = row_apply_exact(s_T, 2, sx, sy, sz);
                kx = voxel_floor(px, map.voxel_size); ky = voxel_floor(py, map.voxel_size); kz = voxel_floor(pz, map.voxel_size);
            }
            const int nq = min(32, prm.n - b * 32);
            int my_slot = -1;
            for (int q = 0; q < nq; ++q) {
                const double qx = __shfl_sync(kFull, px, q), qy = __shfl_sync(kFull, py, q), qz = __shfl_sync(kFull, pz, q);
                const int w = nearest_mean_27(map, qx, qy, qz, __shfl_sync(kFull, kx, q), __shfl_sync(kFull, ky, q), __shfl_sync(kFull, kz, q), lane);
                if (lane == q) my_slot = w;
            }
            if (i < prm.n) {
                double mx = 0.0, my = 0.0, mz = 0.0;  // default CovStruct (I, 0)  (Q2, vhm.cpp:104)
                if (my_slot >= 0) { const double4 vm = map.vslots[my_slot]; mx = vm.y; my = vm.z; mz = vm.w; }
                if (sq3_exact(mx - px, my - py, mz - pz) < prm.max_dist2) linearize_pair(sx, sy, sz, my_slot, mx, my, mz);  // vhm.cpp:129
            }
        }
    } else {
        // AVGICP: 8 lanes per scan point; lane j < 7 owns voxel j of {c, +x, -x, +y, -y, +z, -z} (vhm.cpp:224-230)
        const long long total = static_cast<long long>(prm.n) * 8;
        for (long long t = static_cast<long long>(blockIdx.x) * kIcpThreads + tid; t < total; t += static_cast<long long>(gridDim.x) * kIcpThreads) {
            const int i = static_cast<int>(t >> 3), j = static_cast<int>(t & 7);
            if (j == 7) continue;
            const double sx = scan[3 * static_cast<size_t>(i)], sy = scan[3 * static_cast<size_t>(i) + 1], sz = scan[3 * static_cast<size_t>(i) + 2];
            const double px = row_apply_exact(s_T, 0, sx, sy, sz), py = row_apply_exact(s_T, 1, sx, sy, sz), pz = row_apply_exact(s_T, 2, sx, sy, sz);
            int kx = voxel_floor(px, map.voxel_size), ky = voxel_floor(py, map.voxel_size), kz = voxel_floor(pz, map.voxel_size);
            kx += (j == 1) - (j == 2); ky += (j == 3) - (j == 4); kz += (j == 5) - (j == 6);
            if (!(key_ok(kx) && key_ok(ky) && key_ok(kz))) continue;
            uint32_t s0, c0;
            const int slot = probe(map.slots, map.mask, pack_key(kx, ky, kz), s0, c0);
            if (slot < 0) continue;
            const double4 vm = map.vslots[slot];
            if (sq3_exact(vm.y - px, vm.z - py, vm.w - pz) < prm.max_dist2) linearize_pair(sx, sy, sz, slot, vm.y, vm.z, vm.w);  // vhm.cpp:183
        }
    }
    block_reduce_store<29, false>(acc, s_red, partials, prm.n);
}

// ======================================================================================================================
// small fixed-size kernels
// ======================================================================================================================
namespace {

// general 4x4 inverse by cofactors (stands in for Matrix4d::inverse(), reg.cpp:24)
__device__ void inverse4(const double* m, double* o) {
    const double s0 = m[0] * m[5] - m[4] * m[1], s1 = m[0] * m[6] - m[4] * m[2], s2 = m[0] * m[7] - m[4] * m[3];
    const double s3 = m[1] * m[6] - m[5] * m[2], s4 = m[1] * m[7] - m[5] * m[3], s5 = m[2] * m[7] - m[6] * m[3];
    const double c5 = m[10] * m[15] - m[14] * m[11], c4 = m[9] * m[15] - m[13] * m[11], c3 = m[9] * m[14] - m[13] * m[10];
    const double c2 = m[8] * m[15] - m[12] * m[11], c1 = m[8] * m[14] - m[12] * m[10], c0 = m[8] * m[13] - m[12] * m[9];
    const double inv = 1.0 / (s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0);
    o[0] = (m[5] * c5 - m[6] * c4 + m[7] * c3) * inv;
    o[1] = (-m[1] * c5 + m[2] * c4 - m[3] * c3) * inv;
    o[2] = (m[13] * s5 - m[14] * s4 + m[15] * s3) * inv;
    o[3] = (-m[9] * s5 + m[10] * s4 - m[11] * s3) * inv;
    o[4] = (-m[4] * c5 + m[6] * c2 - m[7] * c1) * inv;
    o[5] = (m[0] * c5 - m[2] * c2 + m[3] * c1) * inv;
    o[6] = (-m[12] * s5 + m[14] * s2 - m[15] * s1) * inv;
    o[7] = (m[8] * s5 - m[10] * s2 + m[11] * s1) * inv;
    o[8] = (m[4] * c4 - m[5] * c2 + m[7] * c0) * inv;
    o[9] = (-m[0] * c4 + m[1] * c2 - m[3] * c0) * inv;
    o[10] = (m[12] * s4 - m[13] * s2 + m[15] * s0) * inv;
    o[11] = (-m[8] * s4 + m[9] * s2 - m[11] * s0) * inv;
    o[12] = (-m[4] * c3 + m[5] * c1 - m[6] * c0) * inv;
    o[13] = (m[0] * c3 - m[1] * c1 + m[2] * c0) * inv;
    o[14] = (-m[12] * s3 + m[13] * s1 - m[14] * s0) * inv;
    o[15] = (m[8] * s3 - m[9] * s1 + m[10] * s0) * inv;
}
__device__ void inverse3(const double* a, double* o) {
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[2] * a[7] - a[1] * a[8], c02 = a[1] * a[5] - a[2] * a[4];
    const double c10 = a[5] * a[6] - a[3] * a[8], c11 = a[0] * a[8] - a[2] * a[6], c12 = a[2] * a[3] - a[0] * a[5];
    const double c20 = a[3] * a[7] - a[4] * a[6], c21 = a[1] * a[6] - a[0] * a[7], c22 = a[0] * a[4] - a[1] * a[3];
    const double inv = 1.0 / (a[0] * c00 + a[1] * c10 + a[2] * c20);
    o[0] = c00 * inv; o[1] = c01 * inv; o[2] = c02 * inv; o[3] = c10 * inv; o[4] = c11 * inv; o[5] = c12 * inv;
    o[6] = c20 * inv; o[7] = c21 * inv; o[8] = c22 * inv;
}
__device__ void refresh_inverses(IcpState* st) {
    inverse4(st->T, st->Tinv);
    const double R[9] = {st->T[0], st->T[1], st->T[2], st->T[4], st->T[5], st->T[6], st->T[8], st->T[9], st->T[10]};
    inverse3(R, st->Rinv);
}

// Symmetric 6x6: pivoted LDL^T (largest remaining diagonal first) with pseudo-inverse of D, like Eigen's
// ldlt().solve() (reg.cpp:56,138,214).  Optionally also the full inverse (GICP local_cov, reg.cpp:141).
__device__ void ldlt6(const double* Ain, const double* b, double* x, double* inv_out) {
    double A[6][6];
    int perm[6];
    for (int i = 0; i < 6; ++i) { perm[i] = i; for (int j = 0; j < 6; ++j) A[i][j] = Ain[6 * i + j]; }
    for (int k = 0; k < 6; ++k) {
        int p = k;
        for (int i = k + 1; i < 6; ++i) if (fabs(A[i][i]) > fabs(A[p][p])) p = i;
        if (p != k) {
            for (int j = 0; j < 6; ++j) { const double t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
            for (int i = 0; i < 6; ++i) { const double t = A[i][k]; A[i][k] = A[i][p]; A[i][p] = t; }
            const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        const double d = A[k][k];
        if (d == 0.0) { for (int i = k + 1; i < 6; ++i) A[i][k] = 0.0; continue; }
        for (int i = k + 1; i < 6; ++i) A[i][k] /= d;
        for (int i = k + 1; i < 6; ++i)
            for (int j = k + 1; j <= i; ++j) { A[i][j] -= A[i][k] * d * A[j][k]; A[j][i] = A[i][j]; }
    }
    const double tol = 1.0 / 1.7976931348623157e308;
    const int nrhs = inv_out ? 7 : 1;
    for (int r = 0; r < nrhs; ++r) {
        double y[6];
        for (int i = 0; i < 6; ++i) y[i] = (r == 0) ? b[perm[i]] : ((perm[i] == r - 1) ? 1.0 : 0.0);
        for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
        for (int i = 0; i < 6; ++i) y[i] = (fabs(A[i][i]) > tol) ? y[i] / A[i][i] : 0.0;
        for (int i = 5; i >= 0; --i) for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
        if (r == 0) { for (int i = 0; i < 6; ++i) x[perm[i]] = y[i]; }
        else { for (int i = 0; i < 6; ++i) inv_out[6 * perm[i] + (r - 1)] = y[i]; }
    }
}

// rotation matrix -> rotation angle, through the quaternion as Eigen's AngleAxisd(Matrix3d) does (reg.cpp:381-382)
__device__ double rotation_angle(const double* R) {
    double qw, qx, qy, qz;
    double t = R[0] + R[4] + R[8];
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        qw = 0.5 * t; t = 0.5 / t;
        qx = (R[7] - R[5]) * t; qy = (R[2] - R[6]) * t; qz = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        double q[3];
        q[i] = 0.5 * t; t = 0.5 / t;
        qw = (R[3 * k + j] - R[3 * j + k]) * t;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
        qx = q[0]; qy = q[1]; qz = q[2];
    }
    const double n = sqrt(qx * qx + qy * qy + qz * qz);
    return (n != 0.0) ? 2.0 * atan2(n, fabs(qw)) : 0.0;
}

}  // namespace

__global__ void icp_begin_kernel(IcpState* st, Pose16 T0) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) st->T[i] = T0.m[i];
        refresh_inverses(st);
        for (int i = 0; i < 36; ++i) st->local_cov[i] = (i % 7 == 0) ? 1.0 : 0.0;  // reg.cpp:280
        st->iterations = 0; st->done = 0; st->overlap_fail = 0;
    }
}

// partials[nblocks][kAcc] -> st->acc, fixed summation order (bit-reproducible run to run)
__global__ void __launch_bounds__(256) icp_reduce_kernel(IcpState* st, const double* __restrict__ partials, int nblocks) {
    __shared__ double s[8][kAcc];
    if (st->done) return;
    const int k = threadIdx.x & 31, g = threadIdx.x >> 5;
    double v = 0.0;
    for (int b = g; b < nblocks; b += 8) v += partials[b * kAcc + k];
    s[g][k] = v;
    __syncthreads();
    if (threadIdx.x < kAcc) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x];
        st->acc[threadIdx.x] = t;
    }
}

// One AlignClouds* tail + the RunRegister bookkeeping around it.
__global__ void icp_solve_kernel(IcpState* st, IcpParams prm) {
    if (threadIdx.x != 0 || st->done) return;
    const double* a = st->acc;
    const double n_corr = a[kIdxNcorr], n_total = a[kIdxNtotal];
    // corres_ratio = (float)i_source_corr_num / i_source_total_num   (reg.cpp:351)
    const float ratio = static_cast<float>(n_corr) / static_cast<float>(n_total);
    if (static_cast<double>(ratio) < prm.min_overlap) {  // reg.cpp:352-356: fail, keep the pose of before this iteration
        st->overlap_fail = 1; st->done = 1;
        return;
    }
    double JTJ[36], JTr[6];
    int k = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { JTJ[6 * i + j] = a[k]; JTJ[6 * j + i] = a[k]; ++k; }
    for (int i = 0; i < 6; ++i) JTr[i] = a[kIdxJtr + i];
    for (int i = 0; i < 36; ++i) st->JTJ[i] = JTJ[i];
    for (int i = 0; i < 6; ++i) st->JTr[i] = JTr[i];
    st->residual_sum = a[kIdxRes];
    st->n_corr = n_corr;
    st->fitness = a[kIdxRes] / n_corr;  // reg.cpp:53,134,210
    double A[36], x[6];
    for (int i = 0; i < 36; ++i) A[i] = JTJ[i];
    for (int i = 0; i < 6; ++i) A[7 * i] = JTJ[7 * i] + prm.lm_lambda * JTJ[7 * i];  // JTJ + lambda diag(JTJ) (Q9)
    ldlt6(A, JTr, x, (prm.method == 1) ? st->local_cov : nullptr);                    // reg.cpp:137-142
    // AngleAxisd(|w|, w/|w|).toRotationMatrix()   (reg.cpp:58-62)
    const double wn2 = x[3] * x[3] + x[4] * x[4] + x[5] * x[5];
    const double angle = sqrt(wn2);
    double ax = x[3], ay = x[4], az = x[5];
    if (wn2 > 0.0) { ax /= angle; ay /= angle; az /= angle; }
    const double s = sin(angle), c = cos(angle), c1 = 1.0 - c;
    double D[16];
    D[0] = c1 * ax * ax + c;       D[1] = c1 * ax * ay - s * az;  D[2] = c1 * ax * az + s * ay;  D[3] = x[0];
    D[4] = c1 * ax * ay + s * az;  D[5] = c1 * ay * ay + c;       D[6] = c1 * ay * az - s * ax;  D[7] = x[1];
    D[8] = c1 * ax * az - s * ay;  D[9] = c1 * ay * az + s * ax;  D[10] = c1 * az * az + c;      D[11] = x[2];
    D[12] = 0.0; D[13] = 0.0; D[14] = 0.0; D[15] = 1.0;
    double Tn[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            Tn[4 * i + j] = st->T[4 * i] * D[j] + st->T[4 * i + 1] * D[4 + j] + st->T[4 * i + 2] * D[8 + j] + st->T[4 * i + 3] * D[12 + j];
    for (int i = 0; i < 16; ++i) st->T[i] = Tn[i];  // last_icp_pose * estimation_local (reg.cpp:378)
    st->iterations += 1;
    const double R[9] = {D[0], D[1], D[2], D[4], D[5], D[6], D[8], D[9], D[10]};
    const double tn = rotation_angle(R) + sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);  // reg.cpp:381-384
    if (tn < prm.term_thr) { st->done = 1; return; }                                      // reg.cpp:385-387
    refresh_inverses(st);
}

// ---- correspondence dump (test hook) -----------------------------------------------------------------------------
__global__ void __launch_bounds__(kIcpThreads)
icp_match_kernel(MapView map, const float* __restrict__ scan, int n, Pose16 T, int method, double max_dist2, int* __restrict__ count, double* __restrict__ target) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = gw; i < n; i += nw) {
        const double sx = scan[3 * static_cast<size_t>(i)], sy = scan[3 * static_cast<size_t>(i) + 1], sz = scan[3 * static_cast<size_t>(i) + 2];
        const double px = row_apply_exact(T.m, 0, sx, sy, sz), py = row_apply_exact(T.m, 1, sx, sy, sz), pz = row_apply_exact(T.m, 2, sx, sy, sz);
        const int kx = voxel_floor(px, map.voxel_size), ky = voxel_floor(py, map.voxel_size), kz = voxel_floor(pz, map.voxel_size);
        if (method == 0 || method == 1) {
            const int w = nearest_point_27(map, px, py, pz, kx, ky, kz, lane);
            double tx = 0, ty = 0, tz = 0;
            if (w >= 0) { const float4 t = map.pts[w]; tx = t.x; ty = t.y; tz = t.z; }
            const bool ok = sq3_exact(tx - px, ty - py, tz - pz) < max_dist2;
            if (lane == 0) {
                count[i] = ok ? 1 : 0;
                target[3 * static_cast<size_t>(i)] = ok ? tx : 0.0; target[3 * static_cast<size_t>(i) + 1] = ok ? ty : 0.0; target[3 * static_cast<size_t>(i) + 2] = ok ? tz : 0.0;
            }
        } else if (method == 2) {
            const int w = nearest_mean_27(map, px, py, pz, kx, ky, kz, lane);
            double mx = 0, my = 0, mz = 0;
            if (w >= 0) { const double4 vm = map.vslots[w]; mx = vm.y; my = vm.z; mz = vm.w; }
            const bool ok = sq3_exact(mx - px, my - py, mz - pz) < max_dist2;
            if (lane == 0) {
                count[i] = ok ? 1 : 0;
                target[3 * static_cast<size_t>(i)] = ok ? mx : 0.0; target[3 * static_cast<size_t>(i) + 1] = ok ? my : 0.0; target[3 * static_cast<size_t>(i) + 2] = ok ? mz : 0.0;
            }
        } else {
            bool ok = false;
            double mx = 0, my = 0, mz = 0;
            if (lane < 7) {
                const int x = kx + (lane == 1) - (lane == 2), y = ky + (lane == 3) - (lane == 4), z = kz + (lane == 5) - (lane == 6);
                if (key_ok(x) && key_ok(y) && key_ok(z)) {
                    uint32_t s0, c0;
                    const int slot = probe(map.slots, map.mask, pack_key(x, y, z), s0, c0);
                    if (slot >= 0) {
                        const double4 vm = map.vslots[slot];
                        mx = vm.y; my = vm.z; mz = vm.w;
                        ok = sq3_exact(mx - px, my - py, mz - pz) < max_dist2;
                    }
                }
            }
            const uint32_t okmask = __ballot_sync(kFull, ok);
            const int pos = __popc(okmask & ((1u << lane) - 1));  // emission order = voxel order
            double* t = target + static_cast<size_t>(i) * 21;
            if (lane < 7) { t[3 * lane] = 0.0; t[3 * lane + 1] = 0.0; t[3 * lane + 2] = 0.0; }
            __syncwarp();
            if (ok) { t[3 * pos] = mx; t[3 * pos + 1] = my; t[3 * pos + 2] = mz; }
            if (lane == 0) count[i] = __popc(okmask);
        }
    }
}

// ======================================================================================================================
// launch wrappers
// ======================================================================================================================
int icp_linearize_grid(const IcpParams& prm, int num_sms) {
    int blocks;
    if (prm.method == 0 || prm.method == 1) {
        const int tile_pts = kIcpWarps * prm.queries_per_warp;
        blocks = (prm.n + tile_pts - 1) / tile_pts;
    } else if (prm.method == 2) {
        blocks = ((prm.n + 31) / 32 + kIcpWarps - 1) / kIcpWarps;
    } else {
        blocks = static_cast<int>((static_cast<long long>(prm.n) * 8 + kIcpThreads - 1) / kIcpThreads);
    }
    const int cap = 2 * num_sms;
    return blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
}

cudaError_t launch_icp_begin(IcpState* st, const double T0[16], cudaStream_t s) {
    Pose16 p;
    for (int i = 0; i < 16; ++i) p.m[i] = T0[i];
    icp_begin_kernel<<<1, 32, 0, s>>>(st, p);
    return cudaGetLastError();
}

cudaError_t launch_icp_linearize(const MapView& map, const float* scan, const IcpParams& prm, const IcpState* st, double* partials, int grid, cudaStream_t s) {
    switch (prm.method) {
        case 0: icp_points_kernel<0><<<grid, kIcpThreads, 0, s>>>(map, scan, prm, st, partials); break;
        case 1: icp_points_kernel<1><<<grid, kIcpThreads, 0, s>>>(map, scan, prm, st, partials); break;
        case 2: icp_voxels_kernel<2><<<grid, kIcpThreads, 0, s>>>(map, scan, prm, st, partials); break;
        default: icp_voxels_kernel<3><<<grid, kIcpThreads, 0, s>>>(map, scan, prm, st, partials); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_icp_reduce(IcpState* st, const double* partials, int nblocks, cudaStream_t s) {
    icp_reduce_kernel<<<1, 256, 0, s>>>(st, partials, nblocks);
    return cudaGetLastError();
}

cudaError_t launch_icp_solve(IcpState* st, const IcpParams& prm, cudaStream_t s) {
    icp_solve_kernel<<<1, 32, 0, s>>>(st, prm);
    return cudaGetLastError();
}

cudaError_t launch_icp_match(const MapView& map, const float* scan, int n, const double T[16], int method, double max_dist2, int* count, double* target, int num_sms, cudaStream_t s) {
    Pose16 p;
    for (int i = 0; i < 16; ++i) p.m[i] = T[i];
    int blocks = (n + kIcpWarps - 1) / kIcpWarps;
    blocks = blocks < 1 ? 1 : (blocks > 8 * num_sms ? 8 * num_sms : blocks);
    icp_match_kernel<<<blocks, kIcpThreads, 0, s>>>(map, scan, n, p, method, max_dist2, count, target);
    return cudaGetLastError();
}

}  // namespace elm
