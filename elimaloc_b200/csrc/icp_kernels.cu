// Hand-written sm_100a kernels of the ICP hot loop.  One ICP iteration = two launches:
//
//   icp_search_points_kernel   P2P/GICP  TransformPoints + GetCorrespondencePoints     reg.hpp:136-148, vhm.cpp:31-88
//        first iteration of a call ("cold"): directory lookup, home voxel, then the voxels that cannot be excluded
//   icp_search_warm_kernel     P2P/GICP  the same search from the second iteration on ("warm"): the previous iteration's
//        match bounds the distance, so only the OCTANTS (half-voxel cells) within that bound are read — same result
//   icp_search_means_kernel    VGICP     TransformPoints + GetCorrespondencesCov       vhm.cpp:90-151
//        -> match[n] (index of the winning map point / voxel slot), win[n] (the matched point itself, streamed by the
//           accumulation), memo[n] (warm start of the next search) — never the 168-B structs
//   icp_accumulate_kernel<M>   AlignCloudsLocal{,PointCov,VoxelCov} accumulation        reg.cpp:28-51 / 85-132 / 171-208
//        block tree reduction, and in the LAST block to finish: fixed-order reduction of all partials + the solve/update step
//   icp_avgicp_kernel          AVGICP    GetCorrespondencesAllCov + AlignCloudsLocalVoxelCov in one kernel (vhm.cpp:153-206)
//        (multi-GPU: before the solve the last block all-reduces the 32 sums over the ranks through peer-memory mailboxes)
//   icp_solve_kernel           overlap gate, LM-damped LDLT solve, exp map, pose update, termination test
//        (reg.cpp:349-356, 53-65, 136-151, 378-387) — separate launch only in the NCCL mode, after the ncclAllReduce
//   icp_begin_kernel           state <- initial guess (+ inverses)                      reg.cpp:298,24,79
//   icp_export_kernel          match[] -> (count, target) dump for the parity tests
// All searches start from ONE lookup in the neighbourhood directory (icp_device.cuh) instead of 27 / 7 probes of a voxel table.
//
// Exactness: the transformed scan point, its voxel key and every DECIDING candidate distance are computed in fp64 with
// explicit round-to-nearest mul/add (never contracted into FMA) in the same association order as the CPU reference, so the
// nearest-neighbour choice — including its first-in-visit-order tie-break — is bit-identical to the reference.  Two
// shortcuts never change that choice: voxels whose bounding box is provably farther than the best candidate found so far are
// skipped (exact pruning; `prune = 0` visits all 27 voxels like the reference does), and candidates are pre-filtered with fp32
// distances inside a proved error band (visit_points) — anything that could win or tie is still decided in fp64.
// The accumulation that follows is plain fp64 (FMA allowed) and is compared with a tolerance.
#include "icp_kernels.cuh"
#include "voxel_key.hpp"

#include <cstdlib>
#include <utility>

namespace elm {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr double kDblMax = 1.7976931348623157e308;

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

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// Every kernel of the ICP loop is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its blocks may become
// resident while the previous kernel of the stream is still finishing (for the search: while the last block of the
// accumulation reduces and solves), run their prologue — barrier init, TMA of the first scan tile — and then block in
// pdl_wait() until the previous grid has completed and its memory is visible.  NOTHING written by an earlier kernel may
// be read (and nothing it reads may be written) before pdl_wait().  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- exact helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double sq3_exact(double dx, double dy, double dz) {  // (dx^2 + dy^2) + dz^2, no FMA
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
// row r of  T * (x, y, z, 1):  ((T0 x + T1 y) + T2 z) + T3, no FMA  (reg.hpp:142-145)
__device__ __forceinline__ double row_apply_exact(const double* T, int r, double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4 * r], x), __dmul_rn(T[4 * r + 1], y)), __dmul_rn(T[4 * r + 2], z)), T[4 * r + 3]);
}
// PointToVoxel (vhm.hpp:176-180): floor(p / vs); also returns the in-cell fraction (for the pruning bound)
__device__ __forceinline__ int voxel_floor(double p, const MapView& map, float* frac = nullptr) {
    // a power-of-two voxel size (the default 1.0) divides exactly by multiplication: skips the software DDIV sequence
    const double q = (map.inv_voxel_size != 0.0) ? __dmul_rn(p, map.inv_voxel_size) : __ddiv_rn(p, map.voxel_size);
    const double f = floor(q);
    if (frac) *frac = static_cast<float>(q - f);
    // saturate far outside the table's key range instead of the reference's undefined int overflow
    return (f >= 2.0e9) ? 2000000000 : ((f <= -2.0e9) ? -2000000000 : static_cast<int>(f));
}
// Neighbourhood-directory lookup of the centre key: two independent 32-byte bucket loads (2-choice cuckoo, 2 slots per
// bucket) — one round trip, no probe chain, lanes of a warp never wait for each other's collisions.  Returns the slot
// index (== row index) or -1 when no voxel of the 27-neighbourhood holds a point; `centre` = descriptor of the centre
// z-column {first point, n(z-1) | n(z) << 10 | n(z+1) << 20}.
__device__ __forceinline__ int dir_lookup(const MapView& map, int kx, int ky, int kz, uint2& centre) {
    if (!(key_in_range(kx) && key_in_range(ky) && key_in_range(kz))) return -1;
    const uint64_t key = pack_key(kx, ky, kz);
    uint32_t b1, b2;
    dir_buckets(key, map.bmask, b1, b2);
    // one 32-byte load per bucket (sm_100 LDG.E.256): half the L1TEX requests of four 16-byte loads
    uint4 s0, s1, s2, s3;
#ifndef ELM_DIR_LD
#define ELM_DIR_LD "ld.global.nc.v8.u32"
#endif
    asm(ELM_DIR_LD " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(s0.x), "=r"(s0.y), "=r"(s0.z), "=r"(s0.w), "=r"(s1.x), "=r"(s1.y), "=r"(s1.z), "=r"(s1.w) : "l"(map.dslots + 2 * static_cast<size_t>(b1)));
    asm(ELM_DIR_LD " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(s2.x), "=r"(s2.y), "=r"(s2.z), "=r"(s2.w), "=r"(s3.x), "=r"(s3.y), "=r"(s3.z), "=r"(s3.w) : "l"(map.dslots + 2 * static_cast<size_t>(b2)));
    const uint32_t klo = static_cast<uint32_t>(key), khi = static_cast<uint32_t>(key >> 32);
    if (s0.x == klo && s0.y == khi) { centre = make_uint2(s0.z, s0.w); return static_cast<int>(2 * b1); }
    if (s1.x == klo && s1.y == khi) { centre = make_uint2(s1.z, s1.w); return static_cast<int>(2 * b1 + 1); }
    if (s2.x == klo && s2.y == khi) { centre = make_uint2(s2.z, s2.w); return static_cast<int>(2 * b2); }
    if (s3.x == klo && s3.y == khi) { centre = make_uint2(s3.z, s3.w); return static_cast<int>(2 * b2 + 1); }
    return -1;
}
// descriptor of column c (= 3 (dx + 1) + (dy + 1)) of the row that belongs to directory slot `si`
__device__ __forceinline__ uint2 dir_column(const MapView& map, int si, int c) {
    return __ldg(reinterpret_cast<const uint2*>(map.drows + row_col_word(static_cast<size_t>(si), c)));
}
// The contiguous run of `pts` that covers the voxels of `zmask` (bit 0: z-1, bit 1: z, bit 2: z+1; != 0) of a column.
// (A mask with a gap also covers the voxel in between: visiting more candidates never changes the exact result.)
__device__ __forceinline__ void column_run(uint2 d, uint32_t zmask, uint32_t& start, uint32_t& len) {
    const uint32_t n0 = d.y & kDirCountMask, n1 = (d.y >> kDirCountBits) & kDirCountMask, n2 = (d.y >> (2 * kDirCountBits)) & kDirCountMask;
    const uint32_t lo = (zmask & 1u) ? 0u : ((zmask & 2u) ? n0 : n0 + n1);
    const uint32_t hi = (zmask & 4u) ? n0 + n1 + n2 : ((zmask & 2u) ? n0 + n1 : n0);
    start = d.x + lo;
    len = hi - lo;
}

// Running best of one search: smallest (d2, rank).  The canonical order — voxels sorted by (x, y, z), insertion order
// inside — IS the reference's visit order (voxels x-outer / y / z-inner, vhm.cpp:234-240; strict <, vhm.cpp:45): among
// equal distances the reference keeps the candidate with the smallest canonical index.  On the device the points of a
// voxel are stored sorted by octant, so that index travels with the point (pts[i].w) as its rank.
struct Best {
    double d2 = kDblMax;
    uint32_t idx = 0xffffffffu;
    uint32_t rank = 0xffffffffu;  // canonical index of the candidate: the tie-break among equal distances
};
constexpr uint32_t kNone = 0xffffffffu;
__device__ __forceinline__ bool closer(double d2, uint32_t rank, const Best& b) { return d2 < b.d2 || (d2 == b.d2 && rank < b.rank); }
__device__ __forceinline__ unsigned long long rank_idx(uint32_t rank, uint32_t idx) { return (static_cast<unsigned long long>(rank) << 32) | idx; }
// 32-byte (two stored points) read-only load: sm_100 LDG.E.256.  With one lane per query every lane touches a different
// 128-byte line, so the L1TEX tag stage — one line per cycle — bounds the search (measured: ~14.5 B/cycle/SM with 16-byte
// loads, profiles/r01b_*); a 32-byte load moves twice the points per tag lookup.
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
#ifndef ELM_PTS_LD
#define ELM_PTS_LD "ld.global.nc.v8.f32"
#endif
    asm(ELM_PTS_LD " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
// Stream `n` consecutive stored points starting at index idx0 and fold them into `b`, exactly (smaller distance, then
// smaller rank).  The loads of one batch are issued TOGETHER, before any of them is consumed: addresses
// past the run are clamped to its last element instead of being guarded by the loop condition, which would make every
// load wait for the previous iteration's exit test (one load in flight per thread — what the compiler produced from the
// plain `for (o < n)` loop whatever the unroll factor).  `pts` is padded by one element so that an aligned pair that
// straddles the end of the array can be read.
#ifndef ELM_BATCH
#define ELM_BATCH 3
#endif
// Exact position of a candidate record.  Stored map points: the fp32 coordinates ARE the exact position (note A of the
// survey: map points are float32 values).  Voxel means (VGICP): the record holds the mean rounded to fp32 for the
// pre-filter plus the voxel's slot; the exact fp64 mean is read from the voxel table.
struct ExactPoint {
    __device__ __forceinline__ void operator()(const float4& q, double& x, double& y, double& z) const { x = q.x; y = q.y; z = q.z; }
    __device__ __forceinline__ uint32_t rank(const float4& q, uint32_t) const { return __float_as_uint(q.w); }
};
template <class Fetch>
__device__ __forceinline__ void fold_exact(const Fetch& fetch, const float4& q, uint32_t idx, double px, double py, double pz, Best& b) {
    double x, y, z;
    fetch(q, x, y, z);
    const double d2 = sq3_exact(x - px, y - py, z - pz);
    const uint32_t r = fetch.rank(q, idx);
    if (closer(d2, r, b)) { b.d2 = d2; b.idx = idx; b.rank = r; }
}
template <class Fetch>
__device__ __forceinline__ void visit_points_exact(const float4* __restrict__ pts, uint32_t idx0, uint32_t n, double px, double py, double pz, Best& b,
                                                   const Fetch& fetch) {
    if (n == 0) return;
    const uint32_t end = idx0 + n;
    constexpr uint32_t K = ELM_BATCH;  // 32-byte pairs per batch
    const uint32_t last_pair = (end - 1) & ~1u;
    for (uint32_t i = idx0 & ~1u; i < end; i += 2 * K) {
        float4 q0[K], q1[K];
#pragma unroll
        for (uint32_t u = 0; u < K; ++u) ldg256(pts + min(i + 2 * u, last_pair), q0[u], q1[u]);
#pragma unroll
        for (uint32_t u = 0; u < K; ++u) {
            const uint32_t pi = i + 2 * u;
            if (pi >= idx0 && pi < end) fold_exact(fetch, q0[u], pi, px, py, pz, b);
            if (pi + 1 < end) fold_exact(fetch, q1[u], pi + 1, px, py, pz, b);  // (pi + 1 >= idx0 always)
        }
    }
}

// The query as the streaming loop needs it: exact fp64 position, its fp32 rounding and the width of the fp32 error band.
struct Query {
    double px, py, pz;
    float fx, fy, fz;
    float band;  // 2^-20 (|x| + |y| + |z|): bounds twice the error the fp32 rounding of the query adds to a distance
    __device__ __forceinline__ Query(double x, double y, double z)
        : px(x), py(y), pz(z), fx(static_cast<float>(x)), fy(static_cast<float>(y)), fz(static_cast<float>(z)) {
        band = (fabsf(fx) + fabsf(fy) + fabsf(fz)) * 9.5367431640625e-07f;
    }
};

// Same result as visit_points_exact — bit for bit — at about half the instructions: the run is scanned with fp32
// distances (no F2F/DADD/DMUL per candidate), keeping the fp32 minimum m, its index and the SECOND smallest value s.
// With r the true distance, u the fp32 difference vector and d the fp32 sum of squares:
//   | |u| - r | <= 2^-24 (1 + 2^-24) |p| + 2^-24 r      (rounding of the query to fp32 + of the three subtractions)
//   sqrt(d) in |u| [1 - 2^-23, 1 + 2^-23]               (FMUL + 2 FFMA, all terms >= 0)
// hence a candidate j can only be at least as close as the fp32 argmin if sqrt(d_j) <= sqrt(m) (1 + 2^-21) + 2^-22 |p|.
// If s lies outside that band (widened 2-4x below to absorb the float evaluation of the bound itself) the fp32 argmin is
// the unique exact nearest point of the run and ONE exact fp64 distance is computed for it; otherwise (a near tie,
// ~1e-3 of the runs on a 100 m map) the run is re-scanned exactly.
// (Voxel means: the candidate's fp32 rounding adds 2^-24 |mean| <= 2^-24 (|p| + r) per candidate; the band needed
// becomes sqrt(m) (1 + 2^-21.2) + 2^-22 |p|, still inside the one used.)
template <class Fetch = ExactPoint>
__device__ __forceinline__ void visit_points(const float4* __restrict__ pts, uint32_t idx0, uint32_t n, const Query& Q, Best& b,
                                             const Fetch& fetch = Fetch()) {
#ifdef ELM_EXACT_SCAN
    visit_points_exact(pts, idx0, n, Q.px, Q.py, Q.pz, b, fetch);
#else
    if (n == 0) return;
    const uint32_t end = idx0 + n;
    constexpr uint32_t K = ELM_BATCH;
    const uint32_t last_pair = (end - 1) & ~1u;
    const float kInf = __int_as_float(0x7f800000);
    float m = kInf, s2 = kInf;
    uint32_t mi = idx0;
    auto fold32 = [&](const float4& q, uint32_t pi, bool valid) {
        const float dx = q.x - Q.fx, dy = q.y - Q.fy, dz = q.z - Q.fz;
        float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        d = valid ? d : kInf;
        s2 = fminf(s2, fmaxf(d, m));
        mi = (d < m) ? pi : mi;
        m = fminf(m, d);
    };
    for (uint32_t i = idx0 & ~1u; i < end; i += 2 * K) {
        float4 q0[K], q1[K];
#pragma unroll
        for (uint32_t u = 0; u < K; ++u) ldg256(pts + min(i + 2 * u, last_pair), q0[u], q1[u]);
#pragma unroll
        for (uint32_t u = 0; u < K; ++u) {
            const uint32_t pi = i + 2 * u;
            fold32(q0[u], pi, pi >= idx0 && pi < end);
            fold32(q1[u], pi + 1, pi + 1 < end);
        }
    }
    auto band_around = [&](float d) {
        const float sd = fmaf(sqrtf(d), 1.00000095367431640625f, Q.band);          // sqrt(d) (1 + 2^-20) + band
        return fmaf(sd * sd, 1.000003814697265625f, 1e-30f);                        // (1 + 2^-18), + underflow slack
    };
    const float T = band_around(m);
    if (s2 > T) {  // (false for NaN / inf: those take the exact path)
        const float4 q = __ldg(pts + mi);
        double x, y, z;
        fetch(q, x, y, z);
        const double d2 = sq3_exact(x - Q.px, y - Q.py, z - Q.pz);
        const uint32_t r = fetch.rank(q, mi);
        if (closer(d2, r, b)) { b.d2 = d2; b.idx = mi; b.rank = r; }
    } else {
        Best e;
        visit_points_exact(pts, idx0, n, Q.px, Q.py, Q.pz, e, fetch);
        if (e.idx != kNone && closer(e.d2, e.rank, b)) b = e;
    }
#endif
}
// folds a candidate found elsewhere (smaller distance wins, then smaller rank)
__device__ __forceinline__ void fold_best(Best& b, const Best& o) {
    if (o.idx != kNone && closer(o.d2, o.rank, b)) b = o;
}

// Squared-distance lower bound helper (fp32, units of voxel_size, deliberately under-estimated): gap along one axis
// between the query and any point STORED under key (kq + o).  Insert keys truncate toward zero (vhm.cpp:275), so along
// one axis the voxel with key c holds p/vs in [c, c+1) for c > 0, (c-1, c] for c < 0 and (-1, 1) for c == 0.
__device__ __forceinline__ float axis_gap2(int kq, int o, float f) {
    const int c = kq + o;
    const float lo = static_cast<float>(o - (c <= 0 ? 1 : 0));
    const float hi = static_cast<float>(o + (c >= 0 ? 1 : 0));
    const float g = fmaxf(fmaxf(fmaxf(lo - f, f - hi), 0.0f) - 1e-5f, 0.0f);
    return g * g;
}
// Bit L set <=> voxel L (outside the mask `skip` of already visited ones) may still hold a point at least as close as
// `best_d2`: a voxel is dropped only when its box is PROVABLY farther, so dropping can never change the result.
__device__ __forceinline__ uint32_t voxels_to_visit(int kx, int ky, int kz, float fx, float fy, float fz, double best_d2, float inv_vs2_up,
                                                    uint32_t skip) {
    const float bound = __double2float_ru(best_d2) * inv_vs2_up;  // best distance so far, voxel units, rounded up
    const float gx[3] = {axis_gap2(kx, -1, fx), axis_gap2(kx, 0, fx), axis_gap2(kx, 1, fx)};
    const float gy[3] = {axis_gap2(ky, -1, fy), axis_gap2(ky, 0, fy), axis_gap2(ky, 1, fy)};
    const float gz[3] = {axis_gap2(kz, -1, fz), axis_gap2(kz, 0, fz), axis_gap2(kz, 1, fz)};
    uint32_t need = 0;
#pragma unroll
    for (int L = 0; L < 27; ++L) {
        const float lb = (gx[L / 9] + gy[(L / 3) % 3] + gz[L % 3]) * 0.9999f;
        if (!(lb > bound)) need |= 1u << L;
    }
    return need & ~skip;
}
// exact fp64 mean of voxel v (first sector of its 128-byte record)
__device__ __forceinline__ void voxel_mean(const MapView& map, uint32_t v, double& mx, double& my, double& mz) {
    const double2* r = reinterpret_cast<const double2*>(map.vrec + static_cast<size_t>(v) * 16);
    const double2 a = __ldg(r), b = __ldg(r + 1);
    mx = a.x; my = a.y; mz = b.x;
}
// Nearest voxel MEAN of the 27 voxels (VGICP, vhm.cpp:92-115), one thread per query: ONE directory lookup, then the entry's
// candidate list — the non-empty voxels of the neighbourhood in the reference's visit order, 8 bytes each: the mean relative
// to the entry's key in 13-bit fixed point + the voxel index (voxel_key.hpp) — is scanned with fp32 distances in the entry's
// own frame (voxel units: no large coordinates, so the only error is the quantisation, <= 4.3e-4 per candidate).  A candidate
// can only be at least as close as the fp32 argmin if its fp32 distance is within 2 x that error of the minimum: if the
// second-smallest value lies outside that band (1.2e-3 voxel sizes, against gaps of order 0.1-1 between voxel means) the
// argmin is the unique exact winner and NOTHING else is read; otherwise (rare) every candidate is decided with its exact
// fp64 mean in visit order (strict <: the first of equals, vhm.cpp:113).  Returns the winning voxel's index or -1.
__device__ __forceinline__ int nearest_mean_27(const MapView& map, double px, double py, double pz) {
    float fx, fy, fz;
    const int kx = voxel_floor(px, map, &fx), ky = voxel_floor(py, map, &fy), kz = voxel_floor(pz, map, &fz);
    uint2 centre;
    const int row = dir_lookup(map, kx, ky, kz, centre);
    if (row < 0) return -1;
    const uint2 d = __ldg(reinterpret_cast<const uint2*>(map.drows + row_word(static_cast<size_t>(row), kRowCandFirst)));  // {first candidate, count}
    if (d.y == 0) return -1;
    const uint32_t first = d.x, end = d.x + d.y;
    const uint32_t last_quad = (end - 1) & ~3u;
    const float kInf = __int_as_float(0x7f800000);
    const float step = 4.0f / static_cast<float>(kVcandAxisMax);
    float m = kInf, s2 = kInf;
    uint32_t mv = 0;
    constexpr uint32_t K = 4;  // 32-byte loads (4 candidates each) in flight
    for (uint32_t i = first & ~3u; i < end; i += 4 * K) {
        uint32_t w[K][8];
#pragma unroll
        for (uint32_t u = 0; u < K; ++u)
            asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]), "=r"(w[u][4]), "=r"(w[u][5]), "=r"(w[u][6]), "=r"(w[u][7])
                : "l"(map.vcand8 + min(i + 4 * u, last_quad)));
#pragma unroll
        for (uint32_t u = 0; u < K; ++u)
#pragma unroll
            for (uint32_t t = 0; t < 4; ++t) {
                const uint32_t idx = i + 4 * u + t, lo = w[u][2 * t], hi = w[u][2 * t + 1];
                const float cx = fmaf(static_cast<float>(lo & kVcandAxisMax), step, -2.0f), cy = fmaf(static_cast<float>((lo >> 13) & kVcandAxisMax), step, -2.0f),
                            cz = fmaf(static_cast<float>(((lo >> 26) | (hi << 6)) & kVcandAxisMax), step, -2.0f);
                const float dx = cx - fx, dy = cy - fy, dz = cz - fz;
                float dd = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                dd = (idx >= first && idx < end) ? dd : kInf;
                s2 = fminf(s2, fmaxf(dd, m));
                mv = (dd < m) ? (hi >> 7) : mv;
                m = fminf(m, dd);
            }
    }
    const float sd = sqrtf(m) + 1.2e-3f;
    if (s2 > sd * sd * 1.00001f) return static_cast<int>(mv);  // (false for NaN / inf: those take the exact path)
    int best = -1;
    double best_d2 = kDblMax;
    for (uint32_t i = first; i < end; ++i) {
        const uint32_t v = static_cast<uint32_t>(__ldg(map.vcand8 + i) >> 39);
        double mx, my, mz;
        voxel_mean(map, v, mx, my, mz);
        const double d2 = sq3_exact(mx - px, my - py, mz - pz);
        if (d2 < best_d2) { best_d2 = d2; best = static_cast<int>(v); }
    }
    return best;
}

}  // namespace

// ======================================================================================================================
// accumulation + reduction (+ solve)
// ======================================================================================================================
namespace {

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

// canonical slot k (upper JtJ row-major, Jtr, residual, count) of the P2P sums
__device__ __forceinline__ double expand_p2p(const double* a, int k) {
    switch (k) {
        case 0: case 6: case 11: return a[0];       // I block
        case 4: return a[3];  case 5: return -a[2];  // -skew(B): (0,4)=bz (0,5)=-by
        case 8: return -a[3]; case 10: return a[1];  // (1,3)=-bz (1,5)=bx
        case 12: return a[2]; case 13: return -a[1]; // (2,3)=by (2,4)=-bx
        case 15: return a[4]; case 16: return a[5]; case 17: return a[6]; case 18: return a[7]; case 19: return a[8]; case 20: return a[9];
        case 21: return a[10]; case 22: return a[11]; case 23: return a[12]; case 24: return a[13]; case 25: return a[14]; case 26: return a[15];
        case 27: return a[16]; case 28: return a[17];
        default: return 0.0;
    }
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

// Symmetric 6x6 in shared memory: pivoted LDL^T (largest remaining diagonal first) with pseudo-inverse of D, like
// Eigen's ldlt().solve() (reg.cpp:56,138,214).  A is destroyed.  Optionally also the full inverse (GICP local_cov,
// reg.cpp:141).  Single thread; everything lives in shared memory (no local-memory indexing).
__device__ void ldlt6(double (*A)[6], const double* b, double* x, double* inv_out, int* perm, double* y) {
    for (int i = 0; i < 6; ++i) perm[i] = i;
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
        const double id = 1.0 / d;
        for (int i = k + 1; i < 6; ++i) A[i][k] *= id;
        for (int i = k + 1; i < 6; ++i)
            for (int j = k + 1; j <= i; ++j) { A[i][j] -= A[i][k] * d * A[j][k]; A[j][i] = A[i][j]; }
    }
    const double tol = 1.0 / kDblMax;
    const int nrhs = inv_out ? 7 : 1;
    for (int r = 0; r < nrhs; ++r) {
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

struct SolveScratch {
    double A[6][6];
    double x[6], y[6];
    int perm[6];
};

// Register-resident LDL^T of a well-conditioned symmetric 6x6 (every loop fully unrolled, static indexing), the common
// case of JtJ + lambda diag(JtJ).  Returns false — without touching x — when a pivot is tiny relative to the largest
// diagonal entry; the caller then takes the pivoted path above, which mirrors Eigen's ldlt() on degenerate input.
__device__ __forceinline__ bool ldlt6_fast(const double (*Ain)[6], const double* b, double* x, double* inv_out) {
    double A[6][6];
    double dmax = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) A[i][j] = Ain[i][j];
        dmax = fmax(dmax, fabs(A[i][i]));
    }
    double dinv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double d = A[k][k];
        if (!(fabs(d) > 1e-13 * dmax)) return false;
        dinv[k] = 1.0 / d;
#pragma unroll
        for (int i = k + 1; i < 6; ++i) A[i][k] *= dinv[k];
#pragma unroll
        for (int i = k + 1; i < 6; ++i)
#pragma unroll
            for (int j = k + 1; j <= i; ++j) A[i][j] -= A[i][k] * d * A[j][k];
    }
    auto solve = [&](double* y) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) y[i] *= dinv[i];
#pragma unroll
        for (int i = 5; i >= 0; --i)
#pragma unroll
            for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
    };
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = b[i];
    solve(y);
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = y[i];
    if (inv_out) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double e[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) e[i] = (i == c) ? 1.0 : 0.0;
            solve(e);
#pragma unroll
            for (int i = 0; i < 6; ++i) inv_out[6 * i + c] = e[i];
        }
    }
    return true;
}

// One AlignClouds* tail + the RunRegister bookkeeping around it.  Runs in ONE thread.
//   a   : the 30 reduced sums (shared memory)       Tcur : current pose rows 0..2 (shared memory)
__device__ void solve_step(IcpState* st, const IcpParams& prm, SolveScratch* sc, const double* a, const double* Tcur) {
    const double n_corr = a[kIdxNcorr], n_total = a[kIdxNtotal];
    // corres_ratio = (float)i_source_corr_num / i_source_total_num   (reg.cpp:351)
    const float ratio = static_cast<float>(n_corr) / static_cast<float>(n_total);
    if (static_cast<double>(ratio) < prm.min_overlap) {  // reg.cpp:352-356: fail, keep the pose of before this iteration
        st->overlap_fail = 1; st->done = 1;
        return;
    }
    int k = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { sc->A[i][j] = a[k]; sc->A[j][i] = a[k]; ++k; }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) st->JTJ[6 * i + j] = sc->A[i][j];
    for (int i = 0; i < 6; ++i) st->JTr[i] = a[kIdxJtr + i];
    st->residual_sum = a[kIdxRes];
    st->n_corr = n_corr;
    st->fitness = a[kIdxRes] / n_corr;  // reg.cpp:53,134,210
    for (int i = 0; i < 6; ++i) sc->A[i][i] += prm.lm_lambda * sc->A[i][i];  // JTJ + lambda diag(JTJ) (Q9)
    double x[6];
    double* cov = (prm.method == 1) ? st->local_cov : nullptr;                // reg.cpp:137-142
    if (!ldlt6_fast(sc->A, a + kIdxJtr, x, cov)) {
        ldlt6(sc->A, a + kIdxJtr, sc->x, cov, sc->perm, sc->y);
        for (int i = 0; i < 6; ++i) x[i] = sc->x[i];
    }
    // AngleAxisd(|w|, w/|w|).toRotationMatrix()   (reg.cpp:58-62)
    const double wn2 = x[3] * x[3] + x[4] * x[4] + x[5] * x[5];
    const double angle = sqrt(wn2);
    double ax = x[3], ay = x[4], az = x[5];
    if (wn2 > 0.0) { const double ia = 1.0 / angle; ax *= ia; ay *= ia; az *= ia; }
    double s, c;
    sincos(angle, &s, &c);
    const double c1 = 1.0 - c;
    const double D0 = c1 * ax * ax + c, D1 = c1 * ax * ay - s * az, D2 = c1 * ax * az + s * ay;
    const double D4 = c1 * ax * ay + s * az, D5 = c1 * ay * ay + c, D6 = c1 * ay * az - s * ax;
    const double D8 = c1 * ax * az - s * ay, D9 = c1 * ay * az + s * ax, D10 = c1 * az * az + c;
    // last_icp_pose * estimation_local (reg.cpp:378); the bottom row of a rigid pose stays (0 0 0 1)
    double Tn[16];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double t0 = Tcur[4 * i], t1 = Tcur[4 * i + 1], t2 = Tcur[4 * i + 2], t3 = Tcur[4 * i + 3];
        Tn[4 * i] = t0 * D0 + t1 * D4 + t2 * D8;
        Tn[4 * i + 1] = t0 * D1 + t1 * D5 + t2 * D9;
        Tn[4 * i + 2] = t0 * D2 + t1 * D6 + t2 * D10;
        Tn[4 * i + 3] = t0 * x[0] + t1 * x[1] + t2 * x[2] + t3;
    }
    {
        const double t0 = st->T[12], t1 = st->T[13], t2 = st->T[14], t3 = st->T[15];
        Tn[12] = t0 * D0 + t1 * D4 + t2 * D8;
        Tn[13] = t0 * D1 + t1 * D5 + t2 * D9;
        Tn[14] = t0 * D2 + t1 * D6 + t2 * D10;
        Tn[15] = t0 * x[0] + t1 * x[1] + t2 * x[2] + t3;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) st->T[i] = Tn[i];
    st->iterations += 1;
    if (prm.term_thr > 0.0) {  // (the metric is >= 0: with a threshold <= 0 the test of reg.cpp:385 can never fire)
        const double R[9] = {D0, D1, D2, D4, D5, D6, D8, D9, D10};
        const double tn = rotation_angle(R) + sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);  // reg.cpp:381-384
        if (tn < prm.term_thr) { st->done = 1; return; }                                      // reg.cpp:385-387
    }
    double Ti[16], Ri[9];
    inverse4(Tn, Ti);
    const double Rn[9] = {Tn[0], Tn[1], Tn[2], Tn[4], Tn[5], Tn[6], Tn[8], Tn[9], Tn[10]};
    inverse3(Rn, Ri);
#pragma unroll
    for (int i = 0; i < 16; ++i) st->Tinv[i] = Ti[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) st->Rinv[i] = Ri[i];
}

// One P2P / GICP correspondence -> accumulators (reg.cpp:28-51 / 85-132).  target = the matched map point (win[i]), m =
// its device index or -1 (GICP reads its covariance record; -1: the reference's default-constructed neighbour at the
// origin, Q2); p = T * s exactly as the search saw it.
template <int METHOD>
__device__ __forceinline__ void linearize_point_pair(const MapView& map, int m, const float4& target, double sx, double sy, double sz, double px,
                                                     double py, double pz, const double* s_Tinv, const double* s_Rinv, double th, double max_dist2,
                                                     double* acc) {
    // target = the matched map point, or the default-constructed neighbour at the origin when nothing was found (Q2, vhm.cpp:37)
    const double tx = target.x, ty = target.y, tz = target.z;
    if (!(sq3_exact(tx - px, ty - py, tz - pz) < max_dist2)) return;  // vhm.cpp:66
    if (METHOD == 0) {
        const double lx = s_Tinv[0] * tx + s_Tinv[1] * ty + s_Tinv[2] * tz + s_Tinv[3];
        const double ly = s_Tinv[4] * tx + s_Tinv[5] * ty + s_Tinv[6] * tz + s_Tinv[7];
        const double lz = s_Tinv[8] * tx + s_Tinv[9] * ty + s_Tinv[10] * tz + s_Tinv[11];
        acc_p2p(acc, sx, sy, sz, lx - sx, ly - sy, lz - sz, th);
    } else {
        double mean[3] = {0.0, 0.0, 0.0}, C[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, nrm[3] = {1.0, 0.0, 0.0};
        if (m >= 0) {
            const double2* r = reinterpret_cast<const double2*>(map.prec + static_cast<size_t>(m) * 16);
            const double2 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3), r4 = __ldg(r + 4), r5 = __ldg(r + 5),
                          r6 = __ldg(r + 6), r7 = __ldg(r + 7);
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
        const double den = th + r2;
        const double w = (th * th) / (den * den) * 0.8 + 0.2;  // reg.cpp:121
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

// mean[3] + cov[9] of voxel v: its 128-byte record, six 16-byte loads of ONE line
__device__ __forceinline__ void load_voxel_record(const MapView& map, uint32_t v, double* rec) {
    const double2* r = reinterpret_cast<const double2*>(map.vrec + static_cast<size_t>(v) * 16);
    const double2 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3), r4 = __ldg(r + 4), r5 = __ldg(r + 5);
    rec[0] = r0.x; rec[1] = r0.y; rec[2] = r1.x; rec[3] = r1.y; rec[4] = r2.x; rec[5] = r2.y;
    rec[6] = r3.x; rec[7] = r3.y; rec[8] = r4.x; rec[9] = r4.y; rec[10] = r5.x; rec[11] = r5.y;
}
// one (scan point, voxel record) pair -> accumulators (reg.cpp:171-208); rec = {mean[3], cov[9]}
__device__ __forceinline__ void linearize_voxel_pair(double* acc, const double* s_Tinv, const double* s_Rinv, double th, double sx, double sy, double sz,
                                                     const double* rec) {
    const double mx = rec[0], my = rec[1], mz = rec[2];
    const double lx = s_Tinv[0] * mx + s_Tinv[1] * my + s_Tinv[2] * mz + s_Tinv[3];
    const double ly = s_Tinv[4] * mx + s_Tinv[5] * my + s_Tinv[6] * mz + s_Tinv[7];
    const double lz = s_Tinv[8] * mx + s_Tinv[9] * my + s_Tinv[10] * mz + s_Tinv[11];
    const double rx = lx - sx, ry = ly - sy, rz = lz - sz;
    const double r2 = rx * rx + ry * ry + rz * rz;
    const double den = th + r2;
    const double w = (th * th) / (den * den);  // reg.cpp:199
    acc[28] += 1.0;                            // the pair counts in the denominator either way (Q7)
    if (w < 0.01) return;                      // reg.cpp:201
    double M[9];
    mahalanobis_local(s_Rinv, rec + 3, M);
    acc_mahalanobis(acc, M, sx, sy, sz, rx, ry, rz, w);
    acc[27] += sqrt(r2);  // reg.cpp:207
}

// Block tree over per-lane accumulators: warp shuffles, then the 8 warps in fixed order; thread k < 29 ADDS the block's
// sum of canonical slot k to s_sum[k].  Every thread of the block must call it.
// The warp stage is a reduce-SCATTER: at the step with lane distance o a lane keeps one half of its values (the lower half if its bit o
// is clear) and hands the other half to its partner, so after the five steps lane l holds the warp's total of accumulator l — 31
// 64-bit exchanges per lane where a butterfly over all NACC values needs 5 NACC (shuffles issue at one warp instruction per clock
// per SM, and every block of the grid reaches this point at the same time).  Each total is formed by the same pairing tree as the
// butterfly's (distance 16, 8, 4, 2, 1; IEEE addition commutes): bit-identical sums (tests/test_block_reduce_model.py; on the B200 the
// pose checksums of both builds are equal, profiles/r02_ab_reduce_scatter_*.txt).  Measured: GICP +11 %, VGICP +3.5 %, AVGICP +2.8 %.
// -DELM_BUTTERFLY_REDUCE keeps the butterfly everywhere, -DELM_P2P_BUTTERFLY for the 18 structured sums of P2P only (A/B switches).
#if defined(ELM_BUTTERFLY_REDUCE)
constexpr bool kButterflyAll = true, kButterflyP2p = true;
#elif defined(ELM_P2P_BUTTERFLY)
constexpr bool kButterflyAll = false, kButterflyP2p = true;
#else
constexpr bool kButterflyAll = false, kButterflyP2p = false;
#endif
// P2P: where lane l's total (accumulator l of the 18 structured sums) goes in the canonical row — expand_p2p read backwards.
// {slot 0, slot 1, slot 2} one byte each (0xff: none), bit 24: the lane writes zeros (the structurally zero slots), bit 25: slot 1 is negated
#define ELM_SC(a, b, c, f) (static_cast<uint32_t>(a) | (static_cast<uint32_t>(b) << 8) | (static_cast<uint32_t>(c) << 16) | (static_cast<uint32_t>(f) << 24))
__constant__ uint32_t kP2pScatter[32] = {
    ELM_SC(0, 6, 11, 0),          // a[0]: the I block (0,0) (1,1) (2,2)
    ELM_SC(10, 13, 0xff, 2),      // a[1]: (1,5) = bx, (2,4) = -bx
    ELM_SC(12, 5, 0xff, 2),       // a[2]: (2,3) = by, (0,5) = -by
    ELM_SC(4, 8, 0xff, 2),        // a[3]: (0,4) = bz, (1,3) = -bz
    ELM_SC(15, 0xff, 0xff, 0), ELM_SC(16, 0xff, 0xff, 0), ELM_SC(17, 0xff, 0xff, 0), ELM_SC(18, 0xff, 0xff, 0), ELM_SC(19, 0xff, 0xff, 0), ELM_SC(20, 0xff, 0xff, 0),  // a[4..9]
    ELM_SC(21, 0xff, 0xff, 0), ELM_SC(22, 0xff, 0xff, 0), ELM_SC(23, 0xff, 0xff, 0), ELM_SC(24, 0xff, 0xff, 0), ELM_SC(25, 0xff, 0xff, 0), ELM_SC(26, 0xff, 0xff, 0),  // a[10..15]
    ELM_SC(27, 0xff, 0xff, 0), ELM_SC(28, 0xff, 0xff, 0),                                                                                                                // a[16], a[17]
    ELM_SC(1, 2, 3, 1), ELM_SC(7, 9, 14, 1),                                                                                                                             // zeros
    ELM_SC(0xff, 0xff, 0xff, 0), ELM_SC(0xff, 0xff, 0xff, 0), ELM_SC(0xff, 0xff, 0xff, 0), ELM_SC(0xff, 0xff, 0xff, 0), ELM_SC(0xff, 0xff, 0xff, 0), ELM_SC(0xff, 0xff, 0xff, 0),
    ELM_SC(0xff, 0xff, 0xff, 0), ELM_SC(0xff, 0xff, 0xff, 0), ELM_SC(0xff, 0xff, 0xff, 0), ELM_SC(0xff, 0xff, 0xff, 0), ELM_SC(0xff, 0xff, 0xff, 0), ELM_SC(0xff, 0xff, 0xff, 0)};
#undef ELM_SC
template <int NACC, bool IS_P2P, int WARPS = kIcpWarps>
__device__ __forceinline__ void block_sum_into(double* acc, double (*s_red)[kAcc], double* s_sum) {
    static_assert(NACC > 16 && NACC <= 32, "the first step pairs accumulator k with k + 16");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (kButterflyAll || (IS_P2P && kButterflyP2p)) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
            double v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            acc[k] = v;
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 29; ++k) s_red[warp][k] = IS_P2P ? expand_p2p(acc, k) : acc[k < NACC ? k : 0];
        }
    } else {
        {
            const bool up = (lane & 16) != 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {  // accumulators k + 16 >= NACC do not exist: zero
                const double hi = (k + 16 < NACC) ? acc[k + 16 < NACC ? k + 16 : 0] : 0.0;
                const double send = up ? acc[k] : hi, keep = up ? hi : acc[k];
                acc[k] = keep + __shfl_xor_sync(kFull, send, 16);
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int k = 0; k < o; ++k) {
                const double send = up ? acc[k] : acc[k + o], keep = up ? acc[k + o] : acc[k];
                acc[k] = keep + __shfl_xor_sync(kFull, send, o);
            }
        }
        // lane l holds the warp's total of accumulator l; the canonical slots of the row (expand_p2p for the 18 structured sums of
        // P2P) are written by the lanes that hold their sources — no run-time switch (it compiled into jump tables: P2P -32 %)
        const double v = acc[0];
        double* const r = s_red[warp];
        if (IS_P2P) {  // (predicated stores from a per-lane table: a chain of `if (lane == ..)` compiles into a jump table as well)
            const uint32_t t = kP2pScatter[lane];
            const uint32_t s0 = t & 0xffu, s1 = (t >> 8) & 0xffu, s2 = (t >> 16) & 0xffu;
            const double w = (t & (1u << 24)) ? 0.0 : v;
            const double w1 = (t & (1u << 25)) ? -w : w;
            if (s0 != 0xffu) r[s0] = w;
            if (s1 != 0xffu) r[s1] = w1;
            if (s2 != 0xffu) r[s2] = w;
        } else if (lane < NACC) {
            r[lane] = v;
        }
    }
    __syncthreads();
    if (tid < 29) {
        double v = 0.0;
        for (int w = 0; w < WARPS; ++w) v += s_red[w][tid];
        s_sum[tid] += v;
    }
    __syncthreads();
}

#ifdef ELM_PHASE_TIMING
#define ELM_ATICK(k) do { const long long t__ = clock64(); if (prm.stats && threadIdx.x == 0) atomicAdd(prm.stats + (k), static_cast<unsigned long long>(t__ - atick)); atick = t__; } while (0)
#else
#define ELM_ATICK(k) do { } while (0)
#endif
// Publish this block's sums and let the LAST block to arrive reduce all partials in a fixed order (bit-reproducible
// whichever block is last) into st->acc and, when `solve_here`, run the solve/update step.
// row0 / nrows / first_row: this grid's blocks write the rows row0 .. row0 + gridDim.x - 1 of `partials`, and the last block
// sums the rows first_row .. first_row + nrows - 1 (default: this grid's own rows).
// (the scan size n_total travels in the row of the grid's first block)
__device__ __forceinline__ void publish_partials(const double* s_sum, const IcpParams& prm, double* __restrict__ partials, int row) {
    const int tid = threadIdx.x;
    if (tid < kAcc) {
        double v = 0.0;
        if (tid < 29) v = s_sum[tid];
        else if (tid == kIdxNtotal) v = (blockIdx.x == 0) ? static_cast<double>(prm.n) : 0.0;
        partials[static_cast<size_t>(row) * kAcc + tid] = v;
    }
}
template <int WARPS = kIcpWarps>
__device__ __forceinline__ void finish_grid(const double* s_sum, double (*s_red)[kAcc], double* s_acc, bool* s_last, SolveScratch* s_solve,
                                            const double* s_T, IcpState* st, const IcpParams& prm, double* __restrict__ partials,
                                            unsigned int* __restrict__ ticket, int solve_here, int row0 = 0, int nrows = -1, int first_row = 0) {
    const int tid = threadIdx.x;
#ifdef ELM_PHASE_TIMING
    long long atick = clock64();
#endif
    publish_partials(s_sum, prm, partials, row0 + static_cast<int>(blockIdx.x));
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        *s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    ELM_ATICK(12);
    if (!*s_last) return;
    __threadfence();
    {
        const int k = tid & 31, g = tid >> 5;
        double v = 0.0;
        const int nb = nrows >= 0 ? nrows : static_cast<int>(gridDim.x);
        const double* const rows = partials + static_cast<size_t>(nrows >= 0 ? first_row : row0) * kAcc;
        // 16 loads in flight per thread, summed in index order; rows past the end contribute +0.0 (the tail used to be one
        // dependent load per row: 9 round trips for 296 rows instead of 3)
        constexpr int kDepth = WARPS == 4 ? 32 : 16;  // (4 warps: 128 rows in ONE pass)
        for (int b = g; b < nb; b += kDepth * WARPS) {
            double t[kDepth];
#pragma unroll
            for (int u = 0; u < kDepth; ++u) {
                const int row = b + u * WARPS;
                t[u] = (row < nb) ? __ldcg(rows + static_cast<size_t>(row) * kAcc + k) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kDepth; ++u) v += t[u];
        }
        s_red[g][k] = v;
    }
    __syncthreads();
    if (tid < kAcc) {
        double t = 0.0;
        for (int i = 0; i < WARPS; ++i) t += s_red[i][tid];
        st->acc[tid] = t;
        s_acc[tid] = t;
    }
    __syncthreads();
    ELM_ATICK(13);
    if (prm.peer.world > 0) {
        // ---- all-reduce over the ranks through peer memory (NVLink stores into every rank's mailbox), fused into this kernel
        const PeerComm& pc = prm.peer;
        PeerMailbox* mine = pc.box[pc.rank];
        __shared__ unsigned long long s_seq;
        __shared__ int s_timeout;
        if (tid == 0) { s_seq = *reinterpret_cast<volatile unsigned long long*>(&mine->seq) + 1; s_timeout = 0; }
        __syncthreads();
        const unsigned long long seq = s_seq;
        const int par = static_cast<int>(seq & 1);
        const unsigned long long tag = (seq & 0xffffffffull) << 32;
        for (int e = tid; e < pc.world * kAcc; e += WARPS * 32) {  // one (destination rank, accumulator) per thread: two self-validating 8-byte stores
            const int p = e / kAcc, k = e % kAcc;
            const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(s_acc[k]));
            volatile unsigned long long* w = pc.box[p]->word[par][pc.rank][k];
            w[0] = (bits & 0xffffffffull) | tag;
            w[1] = (bits >> 32) | tag;
        }
        __syncthreads();  // (s_acc is overwritten below)
        if (tid < kAcc) {
            double t = 0.0;
            const long long t0 = clock64();
            for (int r = 0; r < pc.world; ++r) {  // rank order: identical on every rank
                const volatile unsigned long long* w = mine->word[par][r][tid];
                unsigned long long lo, hi;
                for (;;) {
                    lo = w[0]; hi = w[1];
                    if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
                    if (clock64() - t0 > 6000000000ll) { s_timeout = 1; lo = hi = 0; break; }  // ~3 s: a rank died; do not hang the GPU
                }
                t += __longlong_as_double(static_cast<long long>((lo & 0xffffffffull) | (hi << 32)));
            }
            st->acc[tid] = t;
            s_acc[tid] = t;
        }
        __syncthreads();
        if (tid == 0) {
            *reinterpret_cast<volatile unsigned long long*>(&mine->seq) = seq;
            if (s_timeout) { st->comm_error = 1; st->done = 1; }
        }
        if (s_timeout) { if (tid == 0) *ticket = 0; return; }
    }
    if (tid == 0) {
        *ticket = 0;
        if (solve_here) solve_step(st, prm, s_solve, s_acc, s_T);
    }
    ELM_ATICK(14);
}

}  // namespace

// ======================================================================================================================
// search: P2P / GICP
// ======================================================================================================================
// Block = 256 threads = one tile of 256 scan points, pulled into shared memory by the TMA bulk-copy engine (packed xyz,
// 12 B/point; double-buffered when a block owns several tiles).  TransformPoints is fused (reg.hpp:136-148).
//
// Every query starts with ONE directory lookup of its centre key (two independent 32-byte loads): a miss means the
// 27-neighbourhood is empty; a hit yields the centre z-column's run and the row of the other eight column runs.
//
// COOP = true (default, exact pruning), three phases per tile:
//   A  thread per QUERY : transform, lookup, stream the z-column of the voxel the query falls into, derive which of the
//                         other 24 voxels cannot be excluded and append one work item per COLUMN that still has such
//                         voxels (with their z-mask) to a block-wide list;
//   B  thread per ITEM  : fetch the column's descriptor from the query's row and stream that run (balanced: every lane
//                         busy, instead of each query thread walking its own 0..8 columns while its warp-mates idle);
//   C  merge            : atomicMin on the fp64 distance bits, then on the canonical index among the exact minima,
//                         which reproduces the reference's first-in-visit-order tie-break (vhm.cpp:45).
// COOP = false: every thread streams all 9 columns of its own query — the reference's exhaustive visit.
#ifdef ELM_PHASE_TIMING
#define ELM_USE(cond) do { if (__any_sync(__activemask(), (cond))) asm volatile(""); } while (0)
#define ELM_TICK(k) do { const long long t__ = clock64(); if (prm.stats && (threadIdx.x & 31) == 0) atomicAdd(prm.stats + (k), static_cast<unsigned long long>(t__ - tick)); tick = t__; } while (0)
#else
#define ELM_TICK(k) do { } while (0)
#define ELM_USE(cond) do { } while (0)
#endif
constexpr int kNoItem = 0xffff;
#ifndef ELM_ITEM_CAP
#define ELM_ITEM_CAP 1024
#endif
constexpr int kItemCap = ELM_ITEM_CAP;  // work items per tile kept in shared memory; overflow stays with its owner
// Scope of the work-item list.  Warp scope (default): every warp shares out the items of ITS 32 queries among its own
// lanes and only ever waits for itself (__syncwarp) — the warps of a block drift through the phases independently and hide
// each other's load latency.  Block scope (-DELM_BLOCK_SCOPE): one list per tile and __syncthreads between the phases
// (better balance, but every warp waits for the slowest one: 19 % of the stall samples were barrier waits).
#ifdef ELM_BLOCK_SCOPE
constexpr int kItemGroups = 1;
#else
constexpr int kItemGroups = kIcpWarps;
#endif
constexpr int kGroupCap = kItemCap / kItemGroups;
constexpr int kGroupThreads = kIcpThreads / kItemGroups;

// FUSE = 0 (P2P) / 1 (GICP): the tile's threads go straight on to linearise their correspondence (AlignCloudsLocal /
// AlignCloudsLocalPointCov accumulation), the block tree-reduces, and the last block of the grid reduces all partials and
// solves — ONE launch per ICP iteration, the matched point still hot in L1/L2.  FUSE = -1: search only (match[] out).
template <bool COOP, int FUSE>
#ifndef ELM_MINBLOCKS
#define ELM_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(kIcpThreads, FUSE == 1 ? 3 : (FUSE == 0 ? 4 : ELM_MINBLOCKS))
icp_search_points_kernel(MapView map, const float* __restrict__ scan, const int* __restrict__ orig, IcpParams prm, IcpState* __restrict__ st,
                         IcpWork wk, int solve_here) {
    int* const __restrict__ match = wk.match;
    double* const __restrict__ partials = wk.partials;
    unsigned int* const __restrict__ ticket = wk.ticket;
    constexpr bool kFuse = FUSE >= 0;
    constexpr int NACC = AccSize<FUSE == 0 ? 0 : 1>::value;
    __shared__ double s_Tinv[kFuse ? 12 : 1], s_Rinv[kFuse ? 9 : 1];
    __shared__ double s_red[kFuse ? kIcpWarps : 1][kAcc], s_sum[kAcc], s_acc[kAcc];
    __shared__ SolveScratch s_solve;
    __shared__ bool s_last;
    __shared__ __align__(16) float s_tile[2][kIcpThreads * 3];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_T[12];
    __shared__ double s_px[COOP ? kIcpThreads : 1], s_py[COOP ? kIcpThreads : 1], s_pz[COOP ? kIcpThreads : 1];
    __shared__ unsigned long long s_best[COOP ? kIcpThreads : 1];
    __shared__ unsigned long long s_win[COOP ? kIcpThreads : 1];      // rank << 32 | index of the winner among the exact minima
    __shared__ unsigned long long s_item_d2[COOP ? kItemCap : 1];
    __shared__ unsigned long long s_item_win[COOP ? kItemCap : 1];
    __shared__ uint16_t s_items[COOP ? kItemCap : 1];
    __shared__ int s_nitems[kItemGroups];

    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int grp = tid / kGroupThreads, gtid = tid % kGroupThreads;  // item-list group of this thread and its rank in it
    uint16_t* const g_items = s_items + (COOP ? grp * kGroupCap : 0);
    unsigned long long* const g_item_d2 = s_item_d2 + (COOP ? grp * kGroupCap : 0);
    unsigned long long* const g_item_win = s_item_win + (COOP ? grp * kGroupCap : 0);
    auto group_sync = [&]() { if (kItemGroups == 1) __syncthreads(); else __syncwarp(); };

    // ---- prologue that does not depend on the previous kernel: barriers + the TMA of the first scan tile
    constexpr int tile_pts = kIcpThreads;
    const int ntiles = (prm.n + tile_pts - 1) / tile_pts;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(scan) & 15) == 0;
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
    __syncthreads();

    auto tile_count = [&](int t) { return min(tile_pts, prm.n - t * tile_pts); };
    auto tile_tma_ok = [&](int t) { return base_aligned && ((tile_count(t) * 12) & 15) == 0; };
    auto issue = [&](int t, int buf) {
        const uint32_t bytes = tile_count(t) * 12;
        mbar_expect_tx(&s_bar[buf], bytes);
        tma_load_1d(s_tile[buf], scan + static_cast<size_t>(t) * tile_pts * 3, bytes, &s_bar[buf]);
    };
    const float inv_vs2_up = static_cast<float>(1.0 / (map.voxel_size * map.voxel_size)) * 1.00001f;

    uint32_t visited = 0, searched = 0;
    int tile = blockIdx.x, buf = 0;
    uint32_t phase[2] = {0, 0};
    const bool first_by_tma = tile < ntiles && tile_tma_ok(tile);
    if (first_by_tma && tid == 0) issue(tile, 0);
    // ---- from here on the previous kernel's results (pose, done flag) are read, and match[] is written
    pdl_wait();
    if (st->done) {  // loop already left (termination / overlap failure)
        if (first_by_tma) mbar_wait(&s_bar[0], 0);  // the bulk copy must land before the block's shared memory is released
        return;
    }
    if (tid < 12) s_T[tid] = st->T[tid];
    if (kFuse) {
        if (tid < 12) s_Tinv[tid] = st->Tinv[tid];
        if (tid < 9) s_Rinv[tid] = st->Rinv[tid];
        if (tid < kAcc) s_sum[tid] = 0.0;
    }
    for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const int next = tile + gridDim.x;
        if (tid == 0 && next < ntiles && tile_tma_ok(next)) issue(next, buf ^ 1);  // prefetch the next tile
#ifdef ELM_PHASE_TIMING
        long long tick = clock64();
        if (prm.stats && (tid & 31) == 0) atomicAdd(prm.stats + 8, 1ull);
#endif
        if (gtid == 0) s_nitems[grp] = 0;
        if (tile_tma_ok(tile)) {
            mbar_wait(&s_bar[buf], phase[buf]);
            phase[buf] ^= 1;
        } else {  // ragged last tile / unaligned base: plain cooperative copy
            const int nf = tile_count(tile) * 3;
            for (int i = tid; i < nf; i += kIcpThreads) s_tile[buf][i] = scan[static_cast<size_t>(tile) * tile_pts * 3 + i];
        }
        __syncthreads();
        const int cnt = tile_count(tile);
        const bool mine = tid < cnt;
        ELM_TICK(2);
        // ---- phase A: one thread per query
        Best b;
        double px = 0, py = 0, pz = 0, sx = 0, sy = 0, sz = 0;
        uint32_t own_cols = 0;  // voxel mask (bit 3 c + iz) of the columns this thread has to walk itself
        int row = -1;
        int my_match = -1;
        uint32_t qkey_lo = kNone, qkey_hi = kNone;  // packed floor key of the query (all ones: outside the key range)
        if (mine) {
            const float* sp = &s_tile[buf][tid * 3];
            sx = sp[0]; sy = sp[1]; sz = sp[2];
            px = row_apply_exact(s_T, 0, sx, sy, sz);
            py = row_apply_exact(s_T, 1, sx, sy, sz);
            pz = row_apply_exact(s_T, 2, sx, sy, sz);
            float fx, fy, fz;
            const int kx = voxel_floor(px, map, &fx), ky = voxel_floor(py, map, &fy), kz = voxel_floor(pz, map, &fz);
            ++searched;
#ifdef ELM_PHASE_TIMING
            long long ftick = clock64();
            ELM_USE(kx + ky + kz == 0x7fffffff);
            { const long long t__ = clock64(); if (prm.stats && (tid & 31) == 0) atomicAdd(prm.stats + 16, static_cast<unsigned long long>(t__ - tick)); ftick = t__; }
#endif
            uint2 centre;
            row = dir_lookup(map, kx, ky, kz, centre);
            if (key_in_range(kx) && key_in_range(ky) && key_in_range(kz)) {
                const uint64_t qk = pack_key(kx, ky, kz);
                qkey_lo = static_cast<uint32_t>(qk); qkey_hi = static_cast<uint32_t>(qk >> 32);
            }
#ifdef ELM_PHASE_TIMING
            ELM_USE(row == 0x7fffffff);
            { const long long t__ = clock64(); if (prm.stats && (tid & 31) == 0) atomicAdd(prm.stats + 17, static_cast<unsigned long long>(t__ - ftick)); ftick = t__; }
#endif
            const Query Q(px, py, pz);
            ELM_USE(row == 0x7fffffff);
            ELM_TICK(3);
            if (COOP) {
                s_win[tid] = ~0ull;
                if (row >= 0) {
                    // the z-column holding the voxel whose STORED-key cell contains the query: insert keys truncate toward
                    // zero (vhm.cpp:275), so on a negative axis that cell is the floor key + 1
                    const int hx = (kx < 0) ? 1 : 0, hy = (ky < 0) ? 1 : 0;
                    const int ch = 3 * (hx + 1) + (hy + 1);
                    const uint2 hd = (ch == 4) ? centre : dir_column(map, row, ch);
                    uint32_t rs, rl;
                    // phase A streams that voxel only (-DELM_HOME_COLUMN: its whole z-column — more candidates, fewer items;
                    // measured 36.7 vs 35.1 us)
#ifdef ELM_HOME_COLUMN
                    const uint32_t hz = 7u;
#else
                    const uint32_t hz = (kz < 0) ? 4u : 2u;
#endif
                    column_run(hd, hz, rs, rl);
                    visit_points(map.pts, rs, rl, Q, b);
                    visited += rl;
                    ELM_USE(b.d2 < 0.0);
                    ELM_TICK(4);
                    own_cols = voxels_to_visit(kx, ky, kz, fx, fy, fz, b.d2, inv_vs2_up, hz << (3 * ch));
                    s_px[tid] = px; s_py[tid] = py; s_pz[tid] = pz;
                    int k = 0;
#pragma unroll
                    for (int c = 0; c < 9; ++c) k += ((own_cols >> (3 * c)) & 7u) ? 1 : 0;
                    if (k) {
                        const int pos = atomicAdd(&s_nitems[grp], k);
                        if (pos + k <= kGroupCap) {  // hand the columns to the group; otherwise they stay with this thread
                            int w = pos;
#pragma unroll
                            for (int c = 0; c < 9; ++c) {
                                const uint32_t zm = (own_cols >> (3 * c)) & 7u;
                                if (zm) {
                                    // the item's column descriptor travels with it: cp.async (8 bytes, no registers) into the slot
                                    // that later receives the item's result, so phase B starts streaming at once
                                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(g_item_d2 + w)),
                                                 "l"(map.drows + row_col_word(static_cast<size_t>(row), c)) : "memory");
                                    g_items[w++] = static_cast<uint16_t>((tid << 7) | (c << 3) | zm);
                                }
                            }
                            own_cols = 0;
                        } else {
                            for (int w = pos; w < kGroupCap; ++w) g_items[w] = kNoItem;  // the tail of the list this claim straddles
                        }
                    }
                    asm volatile("cp.async.commit_group;\n\tcp.async.wait_all;" ::: "memory");
                }
                s_best[tid] = static_cast<unsigned long long>(__double_as_longlong(b.d2));  // what phase A found (DBL_MAX: nothing)
            } else if (row >= 0) {
#pragma unroll 1
                for (int c = 0; c < 9; ++c) {  // ascending column = ascending canonical index = the reference's visit order
                    const uint2 d = (c == 4) ? centre : dir_column(map, row, c);
                    uint32_t rs, rl;
                    column_run(d, 7u, rs, rl);
                    visit_points(map.pts, rs, rl, Q, b);
                    visited += rl;
                }
                my_match = (b.idx == 0xffffffffu) ? -1 : static_cast<int>(b.idx);
            }
        }
        if (COOP) {
            group_sync();
            ELM_TICK(5);
            // ---- phase B: one thread per (query, column) item
            const int nitems = min(s_nitems[grp], kGroupCap);  // (items beyond the cap were never written: their owners kept them)
            for (int j = gtid; j < nitems; j += kGroupThreads) {
                const int it = g_items[j];
                if (it == kNoItem) continue;
                const int q = it >> 7;
                uint32_t rs, rl;
                const unsigned long long dbits = g_item_d2[j];
                column_run(make_uint2(static_cast<uint32_t>(dbits), static_cast<uint32_t>(dbits >> 32)), static_cast<uint32_t>(it & 7), rs, rl);
                Best ib;
                visit_points(map.pts, rs, rl, Query(s_px[q], s_py[q], s_pz[q]), ib);
                visited += rl;
                g_item_d2[j] = static_cast<unsigned long long>(__double_as_longlong(ib.d2));
                g_item_win[j] = rank_idx(ib.rank, ib.idx);
                if (ib.idx != 0xffffffffu) atomicMin(&s_best[q], static_cast<unsigned long long>(__double_as_longlong(ib.d2)));
            }
            if (mine) {
                if (own_cols) {
#pragma unroll 1
                    for (int c = 0; c < 9; ++c) {
                        const uint32_t zm = (own_cols >> (3 * c)) & 7u;
                        if (!zm) continue;
                        uint32_t rs, rl;
                        column_run(dir_column(map, row, c), zm, rs, rl);
                        Best ob;
                        visit_points(map.pts, rs, rl, Query(px, py, pz), ob);
                        visited += rl;
                        fold_best(b, ob);
                    }
                }
                if (b.idx != 0xffffffffu) atomicMin(&s_best[tid], static_cast<unsigned long long>(__double_as_longlong(b.d2)));
            }
            group_sync();
            ELM_TICK(6);
            // ---- phase C: among the exact minima the smallest canonical index wins
            for (int j = gtid; j < nitems; j += kGroupThreads) {
                const int it = g_items[j];
                if (it == kNoItem) continue;
                const int q = it >> 7;
                if (g_item_win[j] != ~0ull && g_item_d2[j] == s_best[q]) atomicMin(&s_win[q], g_item_win[j]);
            }
            if (mine && b.idx != 0xffffffffu && static_cast<unsigned long long>(__double_as_longlong(b.d2)) == s_best[tid])
                atomicMin(&s_win[tid], rank_idx(b.rank, b.idx));
            group_sync();
            ELM_TICK(7);
            if (mine) my_match = (s_win[tid] == ~0ull) ? -1 : static_cast<int>(static_cast<uint32_t>(s_win[tid]));
        }
        // the matched point itself (streamed by the accumulation; {0, 0, 0} = the reference's default neighbour, Q2) and the
        // warm start of the next iteration's search
        float4 wpt = make_float4(0.f, 0.f, 0.f, __uint_as_float(kNone));
        if (mine) {
            if (my_match >= 0) { wpt = __ldg(map.pts + my_match); wpt.w = __uint_as_float(static_cast<uint32_t>(my_match)); }
            const size_t gi = static_cast<size_t>(tile) * tile_pts + tid;
            const size_t oi = orig ? static_cast<size_t>(orig[gi]) : gi;
            if (match) match[oi] = my_match;
            if (wk.win) wk.win[oi] = wpt;
            if (wk.memo) {  // (no candidate list yet: the first warm search refreshes)
                wk.memo[oi] = make_uint4(static_cast<uint32_t>(row), qkey_lo, qkey_hi, static_cast<uint32_t>(my_match));
                wk.ncand[oi] = kNone;
            }
        }
        if (kFuse) {  // linearise this tile's correspondences and fold them into the block's running sums
            double acc[NACC];
#pragma unroll
            for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
            if (mine) linearize_point_pair<FUSE == 1 ? 1 : 0>(map, my_match, wpt, sx, sy, sz, px, py, pz, s_Tinv, s_Rinv, prm.th, prm.max_dist2, acc);
            block_sum_into<NACC, FUSE == 0>(acc, s_red, s_sum);
        } else if (next < ntiles) {
            __syncthreads();  // everyone is done with the tile's shared state before it is refilled (block-uniform)
        }
    }
    if (prm.stats) {
        // warp-aggregate, one atomic pair per warp
        for (int o = 16; o > 0; o >>= 1) { visited += __shfl_xor_sync(kFull, visited, o); searched += __shfl_xor_sync(kFull, searched, o); }
        if ((tid & 31) == 0 && searched) {
            atomicAdd(prm.stats, static_cast<unsigned long long>(visited));
            atomicAdd(prm.stats + 1, static_cast<unsigned long long>(searched));
        }
    }
    if (kFuse) finish_grid(s_sum, s_red, s_acc, &s_last, &s_solve, s_T, st, prm, partials, ticket, solve_here);
}

// ======================================================================================================================
// search: P2P / GICP, warm-started (second iteration of a call onwards)
// ======================================================================================================================
// The search of iteration k+1 asks the same question as iteration k from a slightly moved pose.  The point matched last
// time (win[i], still a member of the query's 27 voxels whenever the query's key did not change — memo says so without a
// directory lookup) gives an upper bound d on the nearest distance BEFORE anything of the map is read, so the search only
// has to look at map points that can be at least as close: the OCTANTS (half-voxel cells; the points of a voxel are stored
// sorted by octant, the row of the directory entry carries every voxel's octant offsets) whose box lies within d.
//
// Two paths per query, both exact:
//   REFRESH  the octants within R = d + margin of the query q0 are searched, and every stored point really within R of q0
//            (a handful) is COPIED into the query's own candidate list in HBM (cand[j][i], query-minor: the lanes of a warp
//            read consecutive addresses); q0 and R go into the memo;
//   REUSE    while the query keeps its key and  d' + |q - q0| <= R  (d' = distance from the new position q to the previous
//            match), every map point of the 27 voxels within d' of q is within R of q0, i.e. in the candidate list: the
//            thread streams its list and is done — no directory, no column records, no random access to the map at all.
//            In a converging ICP loop the pose moves by less than the margin, so nearly every query of nearly every
//            iteration takes this path.
// Why the copy: B200 reads RANDOM 32-byte sectors from HBM at 1.25 TB/s (measured, profiles/micro/dep_gather.cu: the rate
// of DRAM row activations, whatever the number of loads in flight) against 6.5 TB/s for streams, and the candidate lists
// of a 131 072-point scan (~40 MB) stay in the 126 MB L2 from one iteration to the next, which the scattered runs of the
// map (one 128-byte line per 1-2 useful sectors) do not.
// Exactness: every point of the 27 voxels at distance <= d' is in the scanned set, the decision is the exact fp64 one of
// visit_points (fp32 pre-filter inside a proved band, exact re-decision on near ties, smallest rank among equals) — so
// the result (nearest point, first in visit order among equals, vhm.cpp:45) is the one the full visit finds.  A query
// whose key changed looks its row up again and keeps the bound only if the old match is still inside its 27 voxels;
// otherwise it visits everything (and its list stays unusable until the next refresh with a finite bound).
// squared gap (voxel units, fp32, deliberately under-estimated) between the in-cell coordinate f of the query (relative to
// its floor key) and the interval [a, b]
__device__ __forceinline__ float interval_gap2(float f, float a, float b) {
    const float g = fmaxf(fmaxf(fmaxf(a - f, f - b), 0.0f) - 1e-5f, 0.0f);
    return g * g;
}
// Along one axis: bit 2 (o + 1) + h set <=> half h of the voxel at offset o (stored key kq + o, span per the truncation
// rule of axis_gap2, halves split at the middle of the span — voxel_key.hpp) is within `bound` of the query on this axis
// alone.  gv[o + 1] = gap to the whole span of that voxel.
__device__ __forceinline__ uint32_t axis_halves(int kq, float f, float bound, float gv[3]) {
    uint32_t m = 0;
#pragma unroll
    for (int o = -1; o <= 1; ++o) {
        const int c = kq + o;
        const float lo = static_cast<float>(o - (c <= 0 ? 1 : 0)), hi = static_cast<float>(o + (c >= 0 ? 1 : 0));
        const float mid = 0.5f * (lo + hi);
        const float g0 = interval_gap2(f, lo, mid), g1 = interval_gap2(f, mid, hi);
        gv[o + 1] = fminf(g0, g1);
        if (!(g0 * 0.9999f > bound)) m |= 1u << (2 * (o + 1));
        if (!(g1 * 0.9999f > bound)) m |= 2u << (2 * (o + 1));
    }
    return m;
}
// stored (insert) key of a map coordinate: static_cast<int>(p / voxel_size), truncation toward zero (vhm.cpp:275)
__device__ __forceinline__ int stored_key(float p, const MapView& map) {
    const double q = (map.inv_voxel_size != 0.0) ? __dmul_rn(static_cast<double>(p), map.inv_voxel_size) : __ddiv_rn(static_cast<double>(p), map.voxel_size);
    return static_cast<int>(q);
}
// Octant o = hz << 2 | hy << 1 | hx, so "x half 0" = octants 0x55, "y half 0" = 0x33, "z half 0" = 0x0f.
__device__ __forceinline__ uint32_t half_pattern(uint32_t h, uint32_t p0, uint32_t p1) { return ((h & 1u) ? p0 : 0u) | ((h & 2u) ? p1 : 0u); }
// The runs of stored points that hold every point of the query's 27 voxels whose octant box is within `bound` (voxel
// units, squared) of the query, for the z-columns of `columns` (bit c = 3 (dx + 1) + (dy + 1)): per column with a needed
// voxel one 32-byte record, per needed voxel its octants lowest..highest needed; runs of z-neighbours that touch (or
// nearly) are merged.  emit(first, end) per run.
template <class F>
__device__ __forceinline__ void octant_runs(const MapView& map, int row, int kx, int ky, int kz, float fx, float fy, float fz, float bound,
                                            uint32_t columns, F&& emit) {
    float gvx[3], gvy[3], gvz[3];
    const uint32_t hx = axis_halves(kx, fx, bound, gvx), hy = axis_halves(ky, fy, bound, gvy), hz = axis_halves(kz, fz, bound, gvz);
    const float gzmin = fminf(fminf(gvz[0], gvz[1]), gvz[2]);
    uint32_t colmask = 0;
#pragma unroll
    for (int c = 0; c < 9; ++c)
        if (((hx >> (2 * (c / 3))) & 3u) && ((hy >> (2 * (c % 3))) & 3u) && !((gvx[c / 3] + gvy[c % 3] + gzmin) * 0.9999f > bound)) colmask |= 1u << c;
    colmask &= columns;
    while (colmask) {
        const int c = __ffs(colmask) - 1;
        colmask &= colmask - 1;
        const int ox = c / 3, oy = c - 3 * ox;
        uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
        asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "l"(map.drows + row_col_word(static_cast<size_t>(row), c)));
        const uint32_t xy = half_pattern((hx >> (2 * ox)) & 3u, 0x55u, 0xaau) & half_pattern((hy >> (2 * oy)) & 3u, 0x33u, 0xccu);
        const float gxy = (ox == 0 ? gvx[0] : (ox == 1 ? gvx[1] : gvx[2])) + (oy == 0 ? gvy[0] : (oy == 1 ? gvy[1] : gvy[2]));
        uint32_t run_s = 0, run_e = 0, vfirst = r0;
#pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
            const uint32_t n = (r1 >> (kDirCountBits * dz)) & kDirCountMask;
            const uint32_t zh = (hz >> (2 * dz)) & 3u;
            if (n && zh && !((gxy + gvz[dz]) * 0.9999f > bound)) {
                const unsigned long long ow = (static_cast<unsigned long long>(dz == 0 ? r3 : (dz == 1 ? r5 : r7)) << 32) | (dz == 0 ? r2 : (dz == 1 ? r4 : r6));
                uint32_t a = 0, e = n;
                if (ow >> 56) {  // octant words valid (byte 7 = n for cap <= 255, 0 otherwise)
                    const uint32_t need = xy & half_pattern(zh, 0x0fu, 0xf0u);
                    const int lo = __ffs(need) - 1, hi = 32 - __clz(need);  // octants lo .. hi - 1
                    a = lo ? static_cast<uint32_t>(ow >> (8 * (lo - 1))) & 0xffu : 0u;
                    e = hi < 8 ? static_cast<uint32_t>(ow >> (8 * (hi - 1))) & 0xffu : n;
                }
                const uint32_t rs = vfirst + a, re = vfirst + e;
                if (re > rs) {
                    if (run_e > run_s && rs <= run_e + 2) run_e = re;  // touches (or nearly) the run of the voxel below: one run
                    else { if (run_e > run_s) emit(run_s, run_e); run_s = rs; run_e = re; }
                }
            }
            vfirst += n;
        }
        if (run_e > run_s) emit(run_s, run_e);
    }
}

#ifdef ELM_PHASE_TIMING
#define ELM_WTICK(k) do { const long long t__ = clock64(); if (prm.stats && lane == 0) atomicAdd(prm.stats + (k), static_cast<unsigned long long>(t__ - wtick)); wtick = t__; } while (0)
#else
#define ELM_WTICK(k) do { } while (0)
#endif
// REFRESH of one query (a query that changed its voxel, used up its margin, or has no usable list yet): look the row up
// again if the key changed, search the octants within R = d + margin exactly, and copy every stored point in them into the
// query's candidate list.  Lives in the refresh kernel: the REUSE path — nearly every query of nearly every iteration —
// must not pay for this path's registers (with both in one kernel it spilled 232 bytes per thread = 58 MB of
// local-memory stores per launch) nor wait for its stragglers.
struct WarmRefresh {
    Best b;           // the exact nearest neighbour (none: idx == kNone)
    float4 wpt;       // the matched point itself
    int row;          // directory row of the query's key or -1
    uint32_t n_new;   // length of the new candidate list; kNone: unusable
    float Rf;         // every point of the 27 voxels within Rf of the query's fp32 rounding is in the list
};
__device__ __forceinline__ void warm_refresh(const MapView& map, float4* out_cand, uint32_t ccap, double warm_margin, uint4 m0, float4 prev,
                                             bool same_key, double px, double py, double pz, WarmRefresh* out) {
    const float kInf = __int_as_float(0x7f800000);
    const float inv_vs2_up = static_cast<float>(1.0 / (map.voxel_size * map.voxel_size)) * 1.00001f;
    constexpr size_t cstride = kIcpThreads;
    float fx, fy, fz;
    const int kx = voxel_floor(px, map, &fx), ky = voxel_floor(py, map, &fy), kz = voxel_floor(pz, map, &fz);
    Best b;
    float4 wpt = make_float4(0.f, 0.f, 0.f, __uint_as_float(kNone));
    int row;
    bool prev_ok = m0.w != kNone;
    if (same_key) {
        row = static_cast<int>(m0.x);  // same voxel as last time: same row, and the old match is one of its candidates
    } else {
        uint2 centre;
        row = dir_lookup(map, kx, ky, kz, centre);
        if (prev_ok) {  // still inside the 27 voxels of the new key?
            const int cx = stored_key(prev.x, map) - kx, cy = stored_key(prev.y, map) - ky, cz = stored_key(prev.z, map) - kz;
            prev_ok = cx >= -1 && cx <= 1 && cy >= -1 && cy <= 1 && cz >= -1 && cz <= 1;
        }
    }
    uint32_t n_new = kNone;
    double R = 0.0;
    if (row >= 0) {
        float bound = kInf;
        if (prev_ok) {
            b.d2 = sq3_exact(static_cast<double>(prev.x) - px, static_cast<double>(prev.y) - py, static_cast<double>(prev.z) - pz);
            b.idx = m0.w; b.rank = __float_as_uint(__ldg(map.pts + m0.w).w);
            R = (sqrt(b.d2) + warm_margin) * (1.0 + 1e-9);
            bound = __double2float_ru(R * R) * inv_vs2_up;
            n_new = 0;  // (the points within a finite bound are worth remembering)
        }
        // the octants bound WHICH points have to be looked at; the list keeps only those really within R (a handful).  One
        // pass per run, two sectors in flight, every distance exact (a run is a few points: a pre-filter would not pay)
        const double R2 = R * R;
        float4 bpt = prev;  // coordinates of the best so far (starts as the previous match)
        auto fold = [&](const float4& c, uint32_t p) {
            const double d2 = sq3_exact(static_cast<double>(c.x) - px, static_cast<double>(c.y) - py, static_cast<double>(c.z) - pz);
            if (closer(d2, __float_as_uint(c.w), b)) { b.d2 = d2; b.rank = __float_as_uint(c.w); b.idx = p; bpt = c; }
            if (d2 <= R2 && n_new != kNone) {
                if (n_new >= ccap) n_new = kNone;  // does not fit: no list
                else { out_cand[static_cast<size_t>(n_new) * cstride] = make_float4(c.x, c.y, c.z, __uint_as_float(p)); ++n_new; }
            }
        };
        octant_runs(map, row, kx, ky, kz, fx, fy, fz, bound, 0x1ffu, [&](uint32_t rs, uint32_t re) {
            const uint32_t last_pair = (re - 1) & ~1u;
            for (uint32_t i = rs & ~1u; i < re; i += 4) {
                float4 q0[2], q1[2];
#pragma unroll
                for (uint32_t u = 0; u < 2; ++u) ldg256(map.pts + min(i + 2 * u, last_pair), q0[u], q1[u]);
#pragma unroll
                for (uint32_t u = 0; u < 2; ++u) {
                    const uint32_t pi = i + 2 * u;
                    if (pi >= rs && pi < re) fold(q0[u], pi);
                    if (pi + 1 >= rs && pi + 1 < re) fold(q1[u], pi + 1);
                }
            }
        });
        if (b.idx != kNone) wpt = make_float4(bpt.x, bpt.y, bpt.z, __uint_as_float(b.idx));
    }
    // what the list is good for: every point of the 27 voxels within R of the query; the memo keeps the query rounded to fp32
    // (q0), so R shrinks by that rounding
    const double ex = px - static_cast<double>(static_cast<float>(px)), ey = py - static_cast<double>(static_cast<float>(py)),
                 ez = pz - static_cast<double>(static_cast<float>(pz));
    out->b = b; out->wpt = wpt; out->row = row; out->n_new = n_new;
    out->Rf = __double2float_rd(R - sqrt(ex * ex + ey * ey + ez * ez) * (1.0 + 1e-9));
}

// REFRESH of ONE query by the whole warp (a straggler: a query that crossed into another voxel while its 31 warp-mates
// reuse their lists).  The kernel is a single wave, so it lasts as long as its slowest warp; a lone thread walking
// lookup -> records -> runs -> points is ~15 dependent memory round trips, the warp does it in three: every lane computes the
// query's bound and masks (same inputs, same results), lanes 0..8 take one z-column each (record + runs), the points of all
// runs are then spread over the lanes, and the results meet in shuffles.  Same arithmetic as warm_refresh.
// q = the owner's lane; every lane passes ITS OWN values, the owner's are broadcast.  s_run: 64 words of the warp.
__device__ __forceinline__ double shfl_double(double v, int src) {
    const long long b = __double_as_longlong(v);
    const int lo = __shfl_sync(kFull, static_cast<int>(b), src), hi = __shfl_sync(kFull, static_cast<int>(b >> 32), src);
    return __longlong_as_double((static_cast<long long>(hi) << 32) | static_cast<unsigned int>(lo));
}
__device__ __forceinline__ void warm_refresh_warp(const MapView& map, float4* cand, uint32_t ccap, double warm_margin, int q, double my_px, double my_py,
                                                  double my_pz, uint4 my_m0, float4 my_prev, bool my_same_key, size_t my_cbase, unsigned int* s_run,
                                                  WarmRefresh* out) {
    const int lane = threadIdx.x & 31;
    const float kInf = __int_as_float(0x7f800000);
    const float inv_vs2_up = static_cast<float>(1.0 / (map.voxel_size * map.voxel_size)) * 1.00001f;
    constexpr size_t cstride = kIcpThreads;
    // the owner's query, in every lane
    const double px = shfl_double(my_px, q), py = shfl_double(my_py, q), pz = shfl_double(my_pz, q);
    uint4 m0; float4 prev;
    m0.x = __shfl_sync(kFull, my_m0.x, q); m0.y = __shfl_sync(kFull, my_m0.y, q); m0.z = __shfl_sync(kFull, my_m0.z, q); m0.w = __shfl_sync(kFull, my_m0.w, q);
    prev.x = __shfl_sync(kFull, my_prev.x, q); prev.y = __shfl_sync(kFull, my_prev.y, q); prev.z = __shfl_sync(kFull, my_prev.z, q);
    prev.w = __shfl_sync(kFull, my_prev.w, q);
    const bool same_key = __shfl_sync(kFull, my_same_key ? 1 : 0, q) != 0;
    const size_t cbase = (static_cast<size_t>(__shfl_sync(kFull, static_cast<unsigned int>(my_cbase >> 32), q)) << 32) |
                         __shfl_sync(kFull, static_cast<unsigned int>(my_cbase), q);
    float fx, fy, fz;
    const int kx = voxel_floor(px, map, &fx), ky = voxel_floor(py, map, &fy), kz = voxel_floor(pz, map, &fz);
    int row;
    bool prev_ok = m0.w != kNone;
    if (same_key) {
        row = static_cast<int>(m0.x);
    } else {
        uint2 centre;
        row = dir_lookup(map, kx, ky, kz, centre);  // (all lanes: the same two sectors)
        if (prev_ok) {
            const int cx = stored_key(prev.x, map) - kx, cy = stored_key(prev.y, map) - ky, cz = stored_key(prev.z, map) - kz;
            prev_ok = cx >= -1 && cx <= 1 && cy >= -1 && cy <= 1 && cz >= -1 && cz <= 1;
        }
    }
    Best b;
    float4 wpt = make_float4(0.f, 0.f, 0.f, __uint_as_float(kNone));
    uint32_t n_new = kNone;
    double R = 0.0;
    if (row >= 0) {
        float bound = kInf;
        if (prev_ok) {
            b.d2 = sq3_exact(static_cast<double>(prev.x) - px, static_cast<double>(prev.y) - py, static_cast<double>(prev.z) - pz);
            b.idx = m0.w; b.rank = __float_as_uint(__ldg(map.pts + m0.w).w);
            R = (sqrt(b.d2) + warm_margin) * (1.0 + 1e-9);
            bound = __double2float_ru(R * R) * inv_vs2_up;
            n_new = 0;
        }
        const double R2 = R * R;
        // lanes 0..8: the runs of one z-column each (at most three per column)
        uint32_t rs3[3] = {0, 0, 0}, rl3[3] = {0, 0, 0};
        int nr = 0;
        if (lane < 9)
            octant_runs(map, row, kx, ky, kz, fx, fy, fz, bound, 1u << lane, [&](uint32_t rs, uint32_t re) {
                if (nr == 0) { rs3[0] = rs; rl3[0] = re - rs; } else if (nr == 1) { rs3[1] = rs; rl3[1] = re - rs; } else { rs3[2] = rs; rl3[2] = re - rs; }
                ++nr;
            });
        // all runs of the query in the warp's table {first point, first flattened index}, and the total number of points
        uint32_t mine_pts = rl3[0] + rl3[1] + rl3[2], off = mine_pts;
        int slot = nr;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(kFull, off, o);
            const int w = __shfl_up_sync(kFull, slot, o);
            if (lane >= o) { off += v; slot += w; }
        }
        const uint32_t total = __shfl_sync(kFull, off, 31);
        const int nruns = __shfl_sync(kFull, slot, 31);
        off -= mine_pts; slot -= nr;
        for (int k = 0; k < nr; ++k) {
            const uint32_t st = k == 0 ? rs3[0] : (k == 1 ? rs3[1] : rs3[2]);
            s_run[2 * (slot + k)] = st;
            s_run[2 * (slot + k) + 1] = off;
            off += k == 0 ? rl3[0] : (k == 1 ? rl3[1] : rl3[2]);
        }
        if (lane == 0) { s_run[2 * nruns] = 0; s_run[2 * nruns + 1] = total; }  // sentinel
        __syncwarp();
        // the points of all runs spread over the lanes: exact distances, the list of those within R, the nearest
        Best lb;            // this lane's nearest
        float4 lpt = wpt;
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t j = base + lane;
            bool within = false;
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t p = 0;
            if (j < total) {
                int r = 0;
                while (j >= s_run[2 * (r + 1) + 1]) ++r;
                p = s_run[2 * r] + (j - s_run[2 * r + 1]);
                c = __ldg(map.pts + p);
                const double d2 = sq3_exact(static_cast<double>(c.x) - px, static_cast<double>(c.y) - py, static_cast<double>(c.z) - pz);
                if (closer(d2, __float_as_uint(c.w), lb)) { lb.d2 = d2; lb.rank = __float_as_uint(c.w); lb.idx = p; lpt = c; }
                within = d2 <= R2;
            }
            const uint32_t wmask = __ballot_sync(kFull, within);
            if (n_new != kNone) {
                const uint32_t cnt = __popc(wmask);
                if (n_new + cnt > ccap) n_new = kNone;  // does not fit: no list
                else {
                    if (within) {
                        const uint32_t pos = n_new + __popc(wmask & ((1u << lane) - 1u));
                        cand[cbase + static_cast<size_t>(pos) * cstride] = make_float4(c.x, c.y, c.z, __uint_as_float(p));
                    }
                    n_new += cnt;
                }
            }
        }
        __syncwarp();
        // the warp's nearest: smaller distance, then smaller rank; the previous match competes through b
        if (lb.idx != kNone && closer(lb.d2, lb.rank, b)) b = lb;
        bool have_pt = lb.idx != kNone && b.idx == lb.idx;  // this lane holds the coordinates of its candidate
        if (!have_pt) lpt = wpt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Best ob;
            ob.d2 = shfl_double(b.d2, lane ^ o);
            ob.idx = __shfl_xor_sync(kFull, b.idx, o);
            ob.rank = __shfl_xor_sync(kFull, b.rank, o);
            float4 opt;
            opt.x = __shfl_xor_sync(kFull, lpt.x, o); opt.y = __shfl_xor_sync(kFull, lpt.y, o); opt.z = __shfl_xor_sync(kFull, lpt.z, o); opt.w = __shfl_xor_sync(kFull, lpt.w, o);
            const bool ohave = __shfl_xor_sync(kFull, have_pt ? 1 : 0, o) != 0;
            if (ob.idx != kNone && (closer(ob.d2, ob.rank, b) || (ob.idx == b.idx && ohave && !have_pt))) { b = ob; lpt = opt; have_pt = ohave; }
        }
        if (b.idx != kNone) {
            wpt = have_pt ? lpt : __ldg(map.pts + b.idx);  // (the previous match won: its coordinates are not in a lane)
            wpt.w = __uint_as_float(b.idx);
        }
    }
    const double ex = px - static_cast<double>(static_cast<float>(px)), ey = py - static_cast<double>(static_cast<float>(py)),
                 ez = pz - static_cast<double>(static_cast<float>(pz));
    if (lane == q) {
        out->b = b; out->wpt = wpt; out->row = row; out->n_new = n_new;
        out->Rf = __double2float_rd(R - sqrt(ex * ex + ey * ey + ez * ez) * (1.0 + 1e-9));
    }
}

// A warm iteration is TWO launches:
//   icp_warm_reuse_kernel    one thread per query, nothing but the REUSE path: memo, transform, test, stream the candidate
//                            list, decide, linearise the correspondence (AlignCloudsLocal / ...PointCov accumulation fused:
//                            the matched point is in registers), block tree -> one row of partial sums per block.  A query
//                            that cannot reuse appends itself to a work list and contributes nothing here.  No call, no
//                            straggler: the kernel is one wave of warps and lasts as long as one warp's instruction stream.
//   icp_warm_refresh_kernel  the work list: a handful of queries per iteration in a converging loop (one warp per query,
//                            warm_refresh_warp), all of them in the first warm iteration of a call (one thread per query,
//                            warm_refresh); their correspondences are linearised here, and the LAST block sums the rows of
//                            both kernels in a fixed order, all-reduces over the ranks (multi-GPU) and solves.
// Can the query's candidate list answer this iteration's question?  The list holds every stored point of the 27 voxels around
// the memo's key k0 within R of q0 (memo1).  With d' = distance from the new position q to the previous match:
//   same key:  d' + |q - q0| <= R  =>  every point of the neighbourhood within d' of q is in the list;
//   the query MOVED INTO ANOTHER VOXEL k1: still true, and the answer is still the reference's, if R < one voxel and both
//   neighbourhoods lie where stored keys are floor keys (every key of k0 - 1 .. k0 + 1 and k1 - 1 .. k1 + 1 >= 1: insert keys
//   truncate toward zero, vhm.cpp:275, so only there does "within one voxel size of q" imply "inside the 27 voxels around
//   floor(q)").  Then the ball of radius d' < vs around q lies inside BOTH neighbourhoods: the previous match is a candidate of
//   GetCorrespondencePoints at k1, so its nearest point is within d' of q, hence within R of q0, hence in the list — and every
//   list point within d' of q belongs to k1's 27 voxels.  Ranks are canonical indices, the same in both visit orders.
//   (The memo keeps k0 and its row: they describe the LIST; the next refresh looks the new key up.)
__device__ __forceinline__ bool warm_list_answers(const MapView& map, const uint4& m0, const uint4& m1, uint32_t nc, uint32_t ccap, const float4& prev,
                                                  double px, double py, double pz, int kx, int ky, int kz, bool in_range, bool same_key) {
    if (m0.w == kNone || nc > ccap) return false;
    const float Rf = __uint_as_float(m1.w);
    if (!same_key) {
        if (!in_range || (m0.y & m0.z) == kNone) return false;
        int ox, oy, oz;
        unpack_key((static_cast<uint64_t>(m0.z) << 32) | m0.y, ox, oy, oz);
        if (!(min(min(kx, ky), kz) >= 2 && min(min(ox, oy), oz) >= 2 && Rf < 0.99f * static_cast<float>(map.voxel_size))) return false;
    }
    const double d2_prev = sq3_exact(static_cast<double>(prev.x) - px, static_cast<double>(prev.y) - py, static_cast<double>(prev.z) - pz);
    const float dx = static_cast<float>(px - static_cast<double>(__uint_as_float(m1.x))), dy = static_cast<float>(py - static_cast<double>(__uint_as_float(m1.y))),
                dz = static_cast<float>(pz - static_cast<double>(__uint_as_float(m1.z)));
    const float delta = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx))) * 1.000001f + 1e-30f;
    const float room = (Rf - delta) * 0.999999f;  // what is left of R for d' (each side rounded against us)
    return room > 0.f && d2_prev <= static_cast<double>(room) * static_cast<double>(room);
}

#ifdef ELM_TRACE
// developer trace (profiles/trace_async.py): wall-clock marks of one iteration in the stats slots; FIRST = first writer, LAST = latest
__device__ __forceinline__ unsigned long long trace_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ELM_TRACE_FIRST(k) do { if (prm.stats && threadIdx.x == 0) atomicCAS(prm.stats + (k), 0ull, trace_now()); } while (0)
#define ELM_TRACE_LAST(k) do { if (prm.stats && threadIdx.x == 0) atomicMax(prm.stats + (k), trace_now()); } while (0)
#else
#define ELM_TRACE_FIRST(k) do { } while (0)
#define ELM_TRACE_LAST(k) do { } while (0)
#endif
// release / acquire on a 64-bit flag in global memory (tile work lists handed from the reuse kernel to the concurrent refresh kernel)
__device__ __forceinline__ void flag_store_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long flag_load_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int METHOD>
__global__ void __launch_bounds__(kIcpThreads, METHOD == 1 ? 3 : 4)
icp_warm_reuse_kernel(MapView map, const float* __restrict__ scan, IcpParams prm, const IcpState* __restrict__ st, IcpWork wk) {
    constexpr int NACC = AccSize<METHOD>::value;
    __shared__ double s_T[12], s_Tinv[12], s_Rinv[9];
    __shared__ double s_red[kIcpWarps][kAcc], s_sum[kAcc];
    __shared__ int s_done;
    __shared__ unsigned int s_wcnt[kIcpWarps];

    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31;
    const float kInf = __int_as_float(0x7f800000);
    uint4* const memo0 = wk.memo;                   // {row, key_lo, key_hi, index of the match}
    uint4* const memo1 = wk.memo + wk.memo_stride;  // {q0.x, q0.y, q0.z, R} as floats: the list holds every point of the 27 voxels within R of q0
    const uint32_t ccap = static_cast<uint32_t>(wk.cand_cap);
    uint32_t visited = 0, searched = 0, refreshed = 0;
    // the scan is never written inside the loop: this thread's first point may be read before the previous kernel has finished
    int gi = blockIdx.x * kIcpThreads + tid;
    float sxf = 0.f, syf = 0.f, szf = 0.f;
    if (gi < prm.n) { sxf = scan[3 * static_cast<size_t>(gi)]; syf = scan[3 * static_cast<size_t>(gi) + 1]; szf = scan[3 * static_cast<size_t>(gi) + 2]; }
    pdl_wait();
    // the first query's memo is requested BEFORE the pose is staged in shared memory: both loads share one round trip
    uint4 m0 = make_uint4(kNone, kNone, kNone, kNone), m1 = make_uint4(0, 0, 0, 0);
    uint32_t nc = kNone;
    float4 prev = make_float4(0.f, 0.f, 0.f, __uint_as_float(kNone));
    if (gi < prm.n) { m0 = memo0[gi]; m1 = memo1[gi]; nc = wk.ncand[gi]; prev = wk.win[gi]; }
    ELM_TRACE_FIRST(2);
    if (tid < 12) { s_T[tid] = st->T[tid]; s_Tinv[tid] = st->Tinv[tid]; }
    if (tid < 9) s_Rinv[tid] = st->Rinv[tid];
    if (tid < kAcc) s_sum[tid] = 0.0;
    if (tid == 0) s_done = st->done;
    __syncthreads();
    if (s_done) {  // loop already left (termination / overlap failure): tell the concurrent refresh kernel, which waits for this block's tiles
        if (wk.epoch && tid == 0)
            for (int t = blockIdx.x; t * kIcpThreads < prm.n; t += gridDim.x)
                flag_store_release(wk.tile_flag + t, (static_cast<unsigned long long>(wk.epoch) << 32) | 0xffffffffull);
        return;
    }
    for (bool first = true; gi - tid < prm.n; gi += gridDim.x * kIcpThreads, first = false) {
        const bool mine = gi < prm.n;
        const int tile = (gi - tid) / kIcpThreads;
        if (mine && !first) {
            m0 = memo0[gi]; m1 = memo1[gi]; nc = wk.ncand[gi]; prev = wk.win[gi];
            sxf = scan[3 * static_cast<size_t>(gi)]; syf = scan[3 * static_cast<size_t>(gi) + 1]; szf = scan[3 * static_cast<size_t>(gi) + 2];
        }
        // candidate j of this query: tile-interleaved (icp_device.cuh) — the 256 queries of a tile keep their lists in one
        // contiguous block, candidate-major inside it, so a warp reads 512 consecutive bytes per candidate
        const float4* const my_cand = wk.cand + static_cast<size_t>(tile) * (static_cast<size_t>(ccap) * kIcpThreads) + static_cast<size_t>(tid);
        constexpr size_t cstride = kIcpThreads;
        const double sx = sxf, sy = syf, sz = szf;
        const double px = row_apply_exact(s_T, 0, sx, sy, sz), py = row_apply_exact(s_T, 1, sx, sy, sz), pz = row_apply_exact(s_T, 2, sx, sy, sz);
        // ---- decision first: REUSE (same key and d' + |q - q0| <= R with d' = distance to the previous match, each side rounded
        // against us), nothing to find (same voxel, empty neighbourhood), or REFRESH — so that the tile's work list can leave
        // for the refresh kernel before the lists are streamed
        uint32_t qkey_lo = kNone, qkey_hi = kNone;
        bool reuse = false, refresh = false;
        if (mine) {
            const int kx = voxel_floor(px, map), ky = voxel_floor(py, map), kz = voxel_floor(pz, map);
            ++searched;
            const bool in_range = key_in_range(kx) && key_in_range(ky) && key_in_range(kz);
            if (in_range) { const uint64_t qk = pack_key(kx, ky, kz); qkey_lo = static_cast<uint32_t>(qk); qkey_hi = static_cast<uint32_t>(qk >> 32); }
            const bool same_key = in_range && m0.y == qkey_lo && m0.z == qkey_hi;
            reuse = warm_list_answers(map, m0, m1, nc, ccap, prev, px, py, pz, kx, ky, kz, in_range, same_key);
            if (!same_key) { qkey_lo = m0.y; qkey_hi = m0.z; }  // (a reused list keeps the key it was built for; a refresh writes the new one)
            // (same voxel as last time and its 27-neighbourhood holds no point: still nothing to find, Q2 applies below)
            refresh = !reuse && !(same_key && static_cast<int>(m0.x) < 0);
            if (refresh) { ++refreshed; if (prm.stats) atomicAdd(prm.stats + (same_key ? 29 : 28), 1ull); }
        }
        // the work list of the refresh kernel: one segment per tile, filled in thread order (no atomics: the refresh kernel's
        // summation order, hence every bit of the result, is the same from run to run)
        const uint32_t rmask = __ballot_sync(kFull, refresh);
        if (lane == 0) s_wcnt[tid >> 5] = __popc(rmask);
        __syncthreads();
        uint32_t base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kIcpWarps; ++w) { if (w < (tid >> 5)) base += s_wcnt[w]; total += s_wcnt[w]; }
        if (refresh) {
            wk.refresh_list[static_cast<size_t>(tile) * kIcpThreads + base + __popc(rmask & ((1u << lane) - 1u))] = static_cast<uint32_t>(gi);
            if (wk.epoch) __threadfence();
        }
        if (wk.epoch) __syncthreads();  // (block-uniform) every entry of the list is written and fenced before the flag goes out
        if (tid == 0) {
            wk.refresh_count[tile] = total;
            if (wk.epoch) flag_store_release(wk.tile_flag + tile, (static_cast<unsigned long long>(wk.epoch) << 32) | total);
        }
        // ---- REUSE path
        double acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
        if (mine && !refresh) {
            int my_match = -1;
            float4 wpt = make_float4(0.f, 0.f, 0.f, __uint_as_float(kNone));  // the matched point {x, y, z, index}
            if (reuse) {
                // one fp32 pass over the list (coalesced: candidate j of the warp's queries is 512 consecutive bytes), four loads
                // in flight, then ONE decision (the band of visit_points): the fp32 argmin is the unique exact winner, or (a near
                // tie) the list is decided again exactly
                const Query Q(px, py, pz);
                float m = kInf, s2 = kInf;
                uint32_t mj = 0;
                for (uint32_t j = 0; j < nc; j += 4) {
                    float4 c[4];  // (addresses past the list are clamped, not predicated: a predicated load sent c[] to local memory)
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u) c[u] = my_cand[static_cast<size_t>(min(j + u, nc - 1)) * cstride];
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u) {
                        const float dx = c[u].x - Q.fx, dy = c[u].y - Q.fy, dz = c[u].z - Q.fz;
                        float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        d = (j + u < nc) ? d : kInf;
                        s2 = fminf(s2, fmaxf(d, m)); mj = (d < m) ? j + u : mj; m = fminf(m, d);
                    }
                }
                visited += nc;
                const float sd = fmaf(sqrtf(m), 1.00000095367431640625f, Q.band);
                const float T = fmaf(sd * sd, 1.000003814697265625f, 1e-30f);
                if (s2 > T) {  // (the exact distance of the unique winner is not needed: nothing is left to compare it with)
                    wpt = my_cand[static_cast<size_t>(mj) * cstride];
                } else {       // near tie (or an empty list): every candidate exactly, smallest rank (read from the map) among equals
                    Best b;
                    for (uint32_t j = 0; j < nc; ++j) {
                        const float4 c = my_cand[static_cast<size_t>(j) * cstride];
                        const double d2 = sq3_exact(static_cast<double>(c.x) - px, static_cast<double>(c.y) - py, static_cast<double>(c.z) - pz);
                        const uint32_t rank = __float_as_uint(__ldg(map.pts + __float_as_uint(c.w)).w);
                        if (closer(d2, rank, b)) { b.d2 = d2; b.rank = rank; b.idx = __float_as_uint(c.w); wpt = c; }
                    }
                }
                my_match = static_cast<int>(__float_as_uint(wpt.w));  // (kNone = -1: an empty list)
            }
            if (wk.match) wk.match[gi] = my_match;
            wk.win[gi] = wpt;
            memo0[gi] = make_uint4(m0.x, qkey_lo, qkey_hi, static_cast<uint32_t>(my_match));
            linearize_point_pair<METHOD>(map, my_match, wpt, sx, sy, sz, px, py, pz, s_Tinv, s_Rinv, prm.th, prm.max_dist2, acc);
        }
        block_sum_into<NACC, METHOD == 0>(acc, s_red, s_sum);  // (two block barriers: s_wcnt may be rewritten by the next tile)
    }
    if (prm.stats) {
        for (int o = 16; o > 0; o >>= 1) {
            visited += __shfl_xor_sync(kFull, visited, o); searched += __shfl_xor_sync(kFull, searched, o); refreshed += __shfl_xor_sync(kFull, refreshed, o);
        }
        if (lane == 0 && searched) {
            atomicAdd(prm.stats, static_cast<unsigned long long>(visited));
            atomicAdd(prm.stats + 1, static_cast<unsigned long long>(searched));
            atomicAdd(prm.stats + 20, static_cast<unsigned long long>(refreshed));  // warm searches that could not reuse their candidate list
            atomicAdd(prm.stats + 21, static_cast<unsigned long long>(searched));   // warm searches
        }
    }
    publish_partials(s_sum, prm, wk.partials, static_cast<int>(blockIdx.x));
    ELM_TRACE_LAST(3);
}

// ---- ONE kernel per warm iteration (every warm iteration but the first of a call) ------------------------------------
// The pair above costs two launches and two grid drains per iteration, and its second kernel is a pure latency chain (wait
// for the first grid, state, work-list counts, a handful of stragglers, fold, ticket, final reduction, solve: ~20 us for
// ~0.1 % of the queries).  Here the warp that owns a straggler refreshes it on the spot, cooperatively (warm_refresh_warp,
// three dependent round trips; an out-of-line call would keep the spills out of the REUSE path, but ptxas 12.9 crashes on it); the block then
// reduces and the last block of THIS grid sums the rows, all-reduces over the ranks and solves.
__device__ __forceinline__ void warm_refresh_warp_call(const MapView& map, float4* cand, uint32_t ccap, double warm_margin, int q, double px, double py, double pz,
                                                    uint4 m0, float4 prev, bool same_key, size_t cbase, unsigned int* s_run, WarmRefresh* out) {
    warm_refresh_warp(map, cand, ccap, warm_margin, q, px, py, pz, m0, prev, same_key, cbase, s_run, out);
}
__device__ __forceinline__ void warm_refresh_call(const MapView& map, float4* out_cand, uint32_t ccap, double warm_margin, uint4 m0, float4 prev, bool same_key,
                                               double px, double py, double pz, WarmRefresh* out) {
    warm_refresh(map, out_cand, ccap, warm_margin, m0, prev, same_key, px, py, pz, out);
}

template <int METHOD>
__global__ void __launch_bounds__(kIcpThreads, METHOD == 1 ? 3 : 4)
icp_warm_kernel(MapView map, const float* __restrict__ scan, IcpParams prm, IcpState* __restrict__ st, IcpWork wk, int solve_here) {
    constexpr int NACC = AccSize<METHOD>::value;
    __shared__ double s_T[12], s_Tinv[12], s_Rinv[9];
    __shared__ double s_red[kIcpWarps][kAcc], s_sum[kAcc], s_acc[kAcc];
    __shared__ SolveScratch s_solve;
    __shared__ bool s_last;
    __shared__ int s_done;
    __shared__ unsigned int s_run[kIcpWarps][64];

    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float kInf = __int_as_float(0x7f800000);
    uint4* const memo0 = wk.memo;
    uint4* const memo1 = wk.memo + wk.memo_stride;
    const uint32_t ccap = static_cast<uint32_t>(wk.cand_cap);
    uint32_t visited = 0, searched = 0, refreshed = 0;
    int gi = blockIdx.x * kIcpThreads + tid;
    float sxf = 0.f, syf = 0.f, szf = 0.f;
    if (gi < prm.n) { sxf = scan[3 * static_cast<size_t>(gi)]; syf = scan[3 * static_cast<size_t>(gi) + 1]; szf = scan[3 * static_cast<size_t>(gi) + 2]; }
    pdl_wait();
    uint4 m0 = make_uint4(kNone, kNone, kNone, kNone), m1 = make_uint4(0, 0, 0, 0);
    uint32_t nc = kNone;
    float4 prev = make_float4(0.f, 0.f, 0.f, __uint_as_float(kNone));
    if (gi < prm.n) { m0 = memo0[gi]; m1 = memo1[gi]; nc = wk.ncand[gi]; prev = wk.win[gi]; }
    if (tid < 12) { s_T[tid] = st->T[tid]; s_Tinv[tid] = st->Tinv[tid]; }
    if (tid < 9) s_Rinv[tid] = st->Rinv[tid];
    if (tid < kAcc) s_sum[tid] = 0.0;
    if (tid == 0) s_done = st->done;
    __syncthreads();
    if (s_done) return;  // loop already left (termination / overlap failure)
    for (bool first = true; gi - tid < prm.n; gi += gridDim.x * kIcpThreads, first = false) {
        const bool mine = gi < prm.n;
        if (mine && !first) {
            m0 = memo0[gi]; m1 = memo1[gi]; nc = wk.ncand[gi]; prev = wk.win[gi];
            sxf = scan[3 * static_cast<size_t>(gi)]; syf = scan[3 * static_cast<size_t>(gi) + 1]; szf = scan[3 * static_cast<size_t>(gi) + 2];
        }
        const size_t cbase = static_cast<size_t>(gi / kIcpThreads) * (static_cast<size_t>(ccap) * kIcpThreads) + static_cast<size_t>(gi % kIcpThreads);
        constexpr size_t cstride = kIcpThreads;
        const double sx = sxf, sy = syf, sz = szf;
        const double px = row_apply_exact(s_T, 0, sx, sy, sz), py = row_apply_exact(s_T, 1, sx, sy, sz), pz = row_apply_exact(s_T, 2, sx, sy, sz);
        uint32_t qkey_lo = kNone, qkey_hi = kNone;
        bool same_key = false, refresh = false;
        int my_match = -1;
        float4 wpt = make_float4(0.f, 0.f, 0.f, __uint_as_float(kNone));  // the matched point {x, y, z, index}
        if (mine) {
            const int kx = voxel_floor(px, map), ky = voxel_floor(py, map), kz = voxel_floor(pz, map);
            ++searched;
            const bool in_range = key_in_range(kx) && key_in_range(ky) && key_in_range(kz);
            if (in_range) { const uint64_t qk = pack_key(kx, ky, kz); qkey_lo = static_cast<uint32_t>(qk); qkey_hi = static_cast<uint32_t>(qk >> 32); }
            same_key = in_range && m0.y == qkey_lo && m0.z == qkey_hi;
            const bool reuse = warm_list_answers(map, m0, m1, nc, ccap, prev, px, py, pz, kx, ky, kz, in_range, same_key);  // (see icp_warm_reuse_kernel)
            if (reuse && !same_key) { qkey_lo = m0.y; qkey_hi = m0.z; }  // (a reused list keeps the key it was built for)
            if (reuse) {
                const float4* const my_cand = wk.cand + cbase;
                const Query Q(px, py, pz);
                float m = kInf, s2 = kInf;
                uint32_t mj = 0;
                for (uint32_t j = 0; j < nc; j += 4) {
                    float4 c[4];
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u) c[u] = my_cand[static_cast<size_t>(min(j + u, nc - 1)) * cstride];
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u) {
                        const float dx = c[u].x - Q.fx, dy = c[u].y - Q.fy, dz = c[u].z - Q.fz;
                        float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        d = (j + u < nc) ? d : kInf;
                        s2 = fminf(s2, fmaxf(d, m)); mj = (d < m) ? j + u : mj; m = fminf(m, d);
                    }
                }
                visited += nc;
                const float sd = fmaf(sqrtf(m), 1.00000095367431640625f, Q.band);
                const float T = fmaf(sd * sd, 1.000003814697265625f, 1e-30f);
                if (s2 > T) {
                    wpt = my_cand[static_cast<size_t>(mj) * cstride];
                } else {
                    Best b;
                    for (uint32_t j = 0; j < nc; ++j) {
                        const float4 c = my_cand[static_cast<size_t>(j) * cstride];
                        const double d2 = sq3_exact(static_cast<double>(c.x) - px, static_cast<double>(c.y) - py, static_cast<double>(c.z) - pz);
                        const uint32_t rank = __float_as_uint(__ldg(map.pts + __float_as_uint(c.w)).w);
                        if (closer(d2, rank, b)) { b.d2 = d2; b.rank = rank; b.idx = __float_as_uint(c.w); wpt = c; }
                    }
                }
                my_match = static_cast<int>(__float_as_uint(wpt.w));
            } else if (same_key && static_cast<int>(m0.x) < 0) {
                // same voxel as last time and its 27-neighbourhood holds no point: still nothing to find (Q2 applies below)
            } else {
                refresh = true;
                ++refreshed;
            }
        }
        // stragglers: a few per warp -> one after the other by the whole warp; many (the pose jumped) -> every lane its own
        uint32_t row_new = m0.x;
        uint32_t rmask = __ballot_sync(kFull, refresh);
        if (rmask) {
            WarmRefresh r;
            bool have = false;
            if (__popc(rmask) <= 6) {
                while (rmask) {
                    const int q = __ffs(rmask) - 1;
                    rmask &= rmask - 1;
                    WarmRefresh t;
                    warm_refresh_warp_call(map, wk.cand, ccap, prm.warm_margin, q, px, py, pz, m0, prev, same_key, cbase, s_run[warp], &t);
                    if (lane == q) { r = t; have = true; }
                    __syncwarp();
                }
            } else if (refresh) {
                warm_refresh_call(map, wk.cand + cbase, ccap, prm.warm_margin, m0, prev, same_key, px, py, pz, &r);
                have = true;
            }
            if (have) {
                my_match = r.b.idx != kNone ? static_cast<int>(r.b.idx) : -1;
                wpt = r.wpt;
                row_new = static_cast<uint32_t>(r.row);
                memo1[gi] = make_uint4(__float_as_uint(static_cast<float>(px)), __float_as_uint(static_cast<float>(py)), __float_as_uint(static_cast<float>(pz)), __float_as_uint(r.Rf));
                wk.ncand[gi] = r.n_new;
            }
        }
        double acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
        if (mine) {
            if (wk.match) wk.match[gi] = my_match;
            wk.win[gi] = wpt;
            memo0[gi] = make_uint4(row_new, qkey_lo, qkey_hi, static_cast<uint32_t>(my_match));
            linearize_point_pair<METHOD>(map, my_match, wpt, sx, sy, sz, px, py, pz, s_Tinv, s_Rinv, prm.th, prm.max_dist2, acc);
        }
        block_sum_into<NACC, METHOD == 0>(acc, s_red, s_sum);
    }
    if (prm.stats) {
        for (int o = 16; o > 0; o >>= 1) {
            visited += __shfl_xor_sync(kFull, visited, o); searched += __shfl_xor_sync(kFull, searched, o); refreshed += __shfl_xor_sync(kFull, refreshed, o);
        }
        if (lane == 0 && searched) {
            atomicAdd(prm.stats, static_cast<unsigned long long>(visited));
            atomicAdd(prm.stats + 1, static_cast<unsigned long long>(searched));
            atomicAdd(prm.stats + 20, static_cast<unsigned long long>(refreshed));
            atomicAdd(prm.stats + 21, static_cast<unsigned long long>(searched));
        }
    }
    finish_grid(s_sum, s_red, s_acc, &s_last, &s_solve, s_T, st, prm, wk.partials, wk.ticket, solve_here);
}

// ---- the refresh kernel BESIDE the reuse kernel of the same iteration -------------------------------------------------------
// icp_warm_refresh_kernel starts its work when the reuse grid has completed; everything it does then — state, work-list counts, a
// handful of stragglers (five dependent round trips each), fold, ticket, final reduction, solve — is a serial latency chain of
// ~17 us per iteration.  Here the refresh blocks are small (128 threads: they fit beside three resident reuse blocks of an SM),
// become resident while the reuse kernel runs (programmatic dependent launch) and take the tiles' work lists AS THE REUSE BLOCKS
// PUBLISH THEM (tile_flag: release / acquire), so the stragglers are refreshed in the shadow of the reuse kernel.  Only then do
// they wait for the reuse grid (griddepcontrol.wait), fold its rows and the tile rows in a FIXED block <- row mapping, and the last
// block reduces, exchanges and solves.  What remains on the critical path after the reuse kernel: one fold, the ticket, the final
// reduction over <= 148 rows and the solve.
//   tiles are handed out dynamically (tile_ticket; blocks that are not resident yet simply find nothing left), but every sum is
//   formed in a fixed order — a tile's stragglers in list order into tile_rows[tile], rows folded by block (index mod grid) — so
//   results stay bit-reproducible from run to run.
//   Before a flag of THIS iteration's epoch has been seen nothing an earlier kernel wrote may be read: the flag is written by a
//   reuse block after ITS griddepcontrol.wait, i.e. after the previous iteration's refresh kernel has completed.
constexpr int kAsyncWarps = 4;
constexpr int kAsyncThreads = kAsyncWarps * 32;
constexpr int kChunkTiles = 16;  // tiles handed out per ticket: their flags are polled side by side, their stragglers form ONE list
template <int METHOD>
__global__ void __launch_bounds__(kAsyncThreads, 4)
icp_warm_refresh_async_kernel(MapView map, const float* __restrict__ scan, IcpParams prm, IcpState* __restrict__ st, IcpWork wk, int rows_before, int solve_here) {
    constexpr int NACC = AccSize<METHOD>::value;
    __shared__ double s_T[12], s_Tinv[12], s_Rinv[9];
    __shared__ double s_red[kAsyncWarps][kAcc], s_sum[kAcc], s_acc[kAcc];
    __shared__ SolveScratch s_solve;
    __shared__ bool s_last;
    __shared__ unsigned int s_run[kAsyncWarps][64];
    __shared__ unsigned long long s_word;
    __shared__ unsigned int s_cnt[kChunkTiles], s_off[kChunkTiles + 1];
    pdl_launch_dependents();
    ELM_TRACE_FIRST(4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ccap = static_cast<uint32_t>(wk.cand_cap);
    uint4* const memo0 = wk.memo;
    uint4* const memo1 = wk.memo + wk.memo_stride;
    const int ntiles = (prm.n + kIcpThreads - 1) / kIcpThreads;
    const int nchunks = (ntiles + kChunkTiles - 1) / kChunkTiles;
    bool have_state = false, loop_left = false;
    double acc[NACC];
    auto load_state = [&]() {
        if (tid < 12) { s_T[tid] = st->T[tid]; s_Tinv[tid] = st->Tinv[tid]; }
        if (tid < 9) s_Rinv[tid] = st->Rinv[tid];
        have_state = true;
        __syncthreads();
    };
    auto commit = [&](uint32_t gi, const WarmRefresh& r, double sx, double sy, double sz, double px, double py, double pz, uint32_t qkey_lo, uint32_t qkey_hi) {
        const int my_match = r.b.idx != kNone ? static_cast<int>(r.b.idx) : -1;
        if (wk.match) wk.match[gi] = my_match;
        wk.win[gi] = r.wpt;
        memo0[gi] = make_uint4(static_cast<uint32_t>(r.row), qkey_lo, qkey_hi, static_cast<uint32_t>(my_match));
        memo1[gi] = make_uint4(__float_as_uint(static_cast<float>(px)), __float_as_uint(static_cast<float>(py)), __float_as_uint(static_cast<float>(pz)), __float_as_uint(r.Rf));
        wk.ncand[gi] = r.n_new;
        linearize_point_pair<METHOD>(map, my_match, r.wpt, sx, sy, sz, px, py, pz, s_Tinv, s_Rinv, prm.th, prm.max_dist2, acc);
    };
    auto load_query = [&](uint32_t gi, double& sx, double& sy, double& sz, double& px, double& py, double& pz, uint4& m0, float4& prev, bool& same_key,
                          uint32_t& qkey_lo, uint32_t& qkey_hi, size_t& cbase) {
        sx = scan[3 * static_cast<size_t>(gi)]; sy = scan[3 * static_cast<size_t>(gi) + 1]; sz = scan[3 * static_cast<size_t>(gi) + 2];
        px = row_apply_exact(s_T, 0, sx, sy, sz); py = row_apply_exact(s_T, 1, sx, sy, sz); pz = row_apply_exact(s_T, 2, sx, sy, sz);
        m0 = memo0[gi]; prev = wk.win[gi];
        const int kx = voxel_floor(px, map), ky = voxel_floor(py, map), kz = voxel_floor(pz, map);
        qkey_lo = qkey_hi = kNone;
        const bool in_range = key_in_range(kx) && key_in_range(ky) && key_in_range(kz);
        if (in_range) { const uint64_t qk = pack_key(kx, ky, kz); qkey_lo = static_cast<uint32_t>(qk); qkey_hi = static_cast<uint32_t>(qk >> 32); }
        same_key = in_range && m0.y == qkey_lo && m0.z == qkey_hi;
        cbase = static_cast<size_t>(gi / kIcpThreads) * (static_cast<size_t>(ccap) * kIcpThreads) + static_cast<size_t>(gi % kIcpThreads);
    };
    // entry `it` of the chunk's combined straggler list (tiles in order, each tile's list in order)
    auto chunk_entry = [&](int tile0, uint32_t it) {
        int j = 0;
        while (it >= s_off[j + 1]) ++j;
        return wk.refresh_list[static_cast<size_t>(tile0 + j) * kIcpThreads + (it - s_off[j])];
    };
    // ---- chunks of tiles, as their work lists arrive
    while (wk.epoch) {
        __syncthreads();  // (s_word / s_cnt of the previous round have been read by everyone)
        if (tid == 0) s_word = atomicAdd(wk.tile_ticket, 1ull);
        __syncthreads();
        const unsigned long long c64 = s_word;
        if (c64 >= static_cast<unsigned long long>(nchunks)) break;
        const int chunk = static_cast<int>(c64), tile0 = chunk * kChunkTiles;
        if (tid < kChunkTiles) {
            uint32_t cnt = 0;
            if (tile0 + tid < ntiles) {  // one thread per tile of the chunk: the flags are awaited side by side
                const long long t0 = clock64();
                for (;;) {
                    const unsigned long long f = flag_load_acquire(wk.tile_flag + tile0 + tid);
                    if (static_cast<unsigned int>(f >> 32) == wk.epoch) { cnt = static_cast<uint32_t>(f); break; }
                    if (clock64() - t0 > 4000000000ll) { cnt = 0xfffffffeu; break; }  // ~2 s: never hang the GPU
                    __nanosleep(32);
                }
            }
            s_cnt[tid] = cnt;
        }
        __syncthreads();
        ELM_TRACE_FIRST(5);
        uint32_t total = 0, marker = 0;
#pragma unroll
        for (int j = 0; j < kChunkTiles; ++j) { const uint32_t c = s_cnt[j]; marker |= (c >= 0xfffffffeu) ? c : 0u; total += (c >= 0xfffffffeu) ? 0u : c; }
        if (marker) {  // the loop was left in an earlier iteration (or a flag never came: report it instead of hanging)
            loop_left = true;
            if (marker == 0xfffffffeu && tid == 0) { st->comm_error = 2; st->done = 1; }
            break;
        }
        double* const crow = wk.tile_rows + static_cast<size_t>(chunk) * kAcc;
        if (total) {
            if (tid == 0) { uint32_t o = 0; for (int j = 0; j < kChunkTiles; ++j) { s_off[j] = o; o += s_cnt[j]; } s_off[kChunkTiles] = o; }
            if (!have_state) load_state(); else __syncthreads();
#pragma unroll
            for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
            if (total <= 4 * kAsyncWarps) {  // a handful: one warp per query
                for (uint32_t it = warp; it < total; it += kAsyncWarps) {
                    const uint32_t gi = chunk_entry(tile0, it);
                    double sx, sy, sz, px, py, pz; uint4 m0; float4 prev; bool same_key; uint32_t qkey_lo, qkey_hi; size_t cbase;
                    load_query(gi, sx, sy, sz, px, py, pz, m0, prev, same_key, qkey_lo, qkey_hi, cbase);  // (every lane the same query)
                    WarmRefresh r;
                    warm_refresh_warp(map, wk.cand, ccap, prm.warm_margin, 0, px, py, pz, m0, prev, same_key, cbase, s_run[warp], &r);
                    if (lane == 0) commit(gi, r, sx, sy, sz, px, py, pz, qkey_lo, qkey_hi);
                }
            } else {                         // many (the pose jumped): one thread per query
                for (uint32_t it = tid; it < total; it += kAsyncThreads) {
                    const uint32_t gi = chunk_entry(tile0, it);
                    double sx, sy, sz, px, py, pz; uint4 m0; float4 prev; bool same_key; uint32_t qkey_lo, qkey_hi; size_t cbase;
                    load_query(gi, sx, sy, sz, px, py, pz, m0, prev, same_key, qkey_lo, qkey_hi, cbase);
                    WarmRefresh r;
                    warm_refresh(map, wk.cand + cbase, ccap, prm.warm_margin, m0, prev, same_key, px, py, pz, &r);
                    commit(gi, r, sx, sy, sz, px, py, pz, qkey_lo, qkey_hi);
                }
            }
            __syncthreads();
            if (tid < kAcc) s_sum[tid] = 0.0;
            __syncthreads();
            block_sum_into<NACC, METHOD == 0, kAsyncWarps>(acc, s_red, s_sum);
            if (tid < kAcc) crow[tid] = tid < 29 ? s_sum[tid] : 0.0;
        } else if (tid < kAcc) {
            crow[tid] = 0.0;
        }
        // the chunk's row is complete: count it (the fold below waits for all chunks of this iteration)
        if (tid < kAcc) __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(wk.tile_ticket + 1, 1ull);
    }
    // the pose of this iteration, fetched in the shadow of the reuse kernel: any flag of this epoch proves that the previous
    // iteration's solve is complete and visible (blocks that handled a chunk have seen one already)
    if (wk.epoch && !have_state && !loop_left && ntiles > 0) {
        __syncthreads();  // (everyone has read the last ticket from s_word)
        if (tid == 0) {
            const long long t0 = clock64();
            unsigned long long f;
            for (;;) {
                f = flag_load_acquire(wk.tile_flag);
                if (static_cast<unsigned int>(f >> 32) == wk.epoch || clock64() - t0 > 4000000000ll) break;
                __nanosleep(64);
            }
            s_word = f;
        }
        __syncthreads();
        if (static_cast<unsigned int>(s_word >> 32) == wk.epoch && static_cast<uint32_t>(s_word) < 0xfffffffeu) load_state();
    }
    ELM_TRACE_LAST(6);
    // ---- from here on the reuse kernel's rows are complete
    pdl_wait();
    ELM_TRACE_LAST(7);
    if (loop_left) return;
    // speculative loads of the fold (rows of the reuse grid) share their round trip with the done flag
    __shared__ int s_flagdone;
    constexpr int kFold = 8;  // (592 reuse rows over 80 refresh blocks: one round trip)
    double fold[kFold];
#pragma unroll
    for (int u = 0; u < kFold; ++u) fold[u] = 0.0;
    if (tid < 29) {
#pragma unroll
        for (int u = 0; u < kFold; ++u) {
            const int row = static_cast<int>(blockIdx.x) + u * static_cast<int>(gridDim.x);
            if (row < rows_before) fold[u] = __ldcg(wk.partials + static_cast<size_t>(row) * kAcc + tid);
        }
    }
    if (tid == 0) {
        // the done flag and the chunk counter are requested together (they used to be two dependent round trips)
        int d = *reinterpret_cast<volatile int*>(&st->done);
        unsigned long long c = ~0ull;
        if (wk.epoch) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(c) : "l"(wk.tile_ticket + 1) : "memory");
        if (!d && wk.epoch) {  // every chunk row of this iteration must be complete (written by OTHER blocks of this grid)
            const long long t0 = clock64();
            while (c < static_cast<unsigned long long>(nchunks)) {
                if (clock64() - t0 > 4000000000ll) { st->comm_error = 2; st->done = 1; d = 1; break; }
                __nanosleep(32);
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(c) : "l"(wk.tile_ticket + 1) : "memory");
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");  // (acquire: the chunk rows read below)
        }
        s_flagdone = d;
    }
    __syncthreads();
    if (s_flagdone) return;
    if (!have_state) load_state();
    // fold: block b takes the reuse rows b, b + G, ... and the chunk rows b, b + G, ... in that order
    if (tid < kAcc) {
        double v = 0.0;
        if (tid < 29) {
            v = fold[0];  // rows b, b + G, b + 2 G, ... summed in that order (rows past the end contribute +0.0)
#pragma unroll
            for (int u = 1; u < kFold; ++u) v += fold[u];
            for (int row = static_cast<int>(blockIdx.x) + kFold * static_cast<int>(gridDim.x); row < rows_before; row += static_cast<int>(gridDim.x))
                v += __ldcg(wk.partials + static_cast<size_t>(row) * kAcc + tid);
            if (wk.epoch)
                for (int c = static_cast<int>(blockIdx.x); c < nchunks; c += static_cast<int>(gridDim.x)) v += __ldcg(wk.tile_rows + static_cast<size_t>(c) * kAcc + tid);
        }
        s_sum[tid] = v;
    }
    __syncthreads();
    ELM_TRACE_LAST(9);
    finish_grid<kAsyncWarps>(s_sum, s_red, s_acc, &s_last, &s_solve, s_T, st, prm, wk.partials, wk.ticket, solve_here, rows_before, -1, rows_before);
    ELM_TRACE_LAST(10);
}

// rows_before = rows of `partials` the reuse kernel published (its grid size)
template <int METHOD>
__global__ void __launch_bounds__(kIcpThreads, 2)
icp_warm_refresh_kernel(MapView map, const float* __restrict__ scan, IcpParams prm, IcpState* __restrict__ st, IcpWork wk, int rows_before, int solve_here) {
    constexpr int NACC = AccSize<METHOD>::value;
    __shared__ double s_T[12], s_Tinv[12], s_Rinv[9];
    __shared__ double s_red[kIcpWarps][kAcc], s_sum[kAcc], s_acc[kAcc];
    __shared__ SolveScratch s_solve;
    __shared__ bool s_last;
    __shared__ unsigned int s_run[kIcpWarps][64];
    pdl_launch_dependents();
    pdl_wait();
    if (st->done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 12) { s_T[tid] = st->T[tid]; s_Tinv[tid] = st->Tinv[tid]; }
    if (tid < 9) s_Rinv[tid] = st->Rinv[tid];
    if (tid < kAcc) s_sum[tid] = 0.0;
    __syncthreads();
    const uint32_t ccap = static_cast<uint32_t>(wk.cand_cap);
    uint4* const memo0 = wk.memo;
    uint4* const memo1 = wk.memo + wk.memo_stride;
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    // what one refreshed query leaves behind (executed by the thread that owns the result)
    auto commit = [&](uint32_t gi, const WarmRefresh& r, double sx, double sy, double sz, double px, double py, double pz, uint32_t qkey_lo, uint32_t qkey_hi) {
        const int my_match = r.b.idx != kNone ? static_cast<int>(r.b.idx) : -1;
        if (wk.match) wk.match[gi] = my_match;
        wk.win[gi] = r.wpt;
        memo0[gi] = make_uint4(static_cast<uint32_t>(r.row), qkey_lo, qkey_hi, static_cast<uint32_t>(my_match));
        memo1[gi] = make_uint4(__float_as_uint(static_cast<float>(px)), __float_as_uint(static_cast<float>(py)), __float_as_uint(static_cast<float>(pz)), __float_as_uint(r.Rf));
        wk.ncand[gi] = r.n_new;
        linearize_point_pair<METHOD>(map, my_match, r.wpt, sx, sy, sz, px, py, pz, s_Tinv, s_Rinv, prm.th, prm.max_dist2, acc);
    };
    auto load_query = [&](uint32_t gi, double& sx, double& sy, double& sz, double& px, double& py, double& pz, uint4& m0, float4& prev, bool& same_key,
                          uint32_t& qkey_lo, uint32_t& qkey_hi, size_t& cbase) {
        sx = scan[3 * static_cast<size_t>(gi)]; sy = scan[3 * static_cast<size_t>(gi) + 1]; sz = scan[3 * static_cast<size_t>(gi) + 2];
        px = row_apply_exact(s_T, 0, sx, sy, sz); py = row_apply_exact(s_T, 1, sx, sy, sz); pz = row_apply_exact(s_T, 2, sx, sy, sz);
        m0 = memo0[gi]; prev = wk.win[gi];
        const int kx = voxel_floor(px, map), ky = voxel_floor(py, map), kz = voxel_floor(pz, map);
        qkey_lo = qkey_hi = kNone;
        const bool in_range = key_in_range(kx) && key_in_range(ky) && key_in_range(kz);
        if (in_range) { const uint64_t qk = pack_key(kx, ky, kz); qkey_lo = static_cast<uint32_t>(qk); qkey_hi = static_cast<uint32_t>(qk >> 32); }
        same_key = in_range && m0.y == qkey_lo && m0.z == qkey_hi;
        cbase = static_cast<size_t>(gi / kIcpThreads) * (static_cast<size_t>(ccap) * kIcpThreads) + static_cast<size_t>(gi % kIcpThreads);
    };
    // this block's tiles: blockIdx.x, blockIdx.x + gridDim.x, ... two at a time, so that their few stragglers (a converging
    // loop: 0-2 per tile) share the block's eight warps instead of queueing tile after tile
    const int ntiles = (prm.n + kIcpThreads - 1) / kIcpThreads;
    // rows_before < 0: EVERY query is refreshed (the first warm iteration of a call, whose lists do not exist yet) — no reuse kernel
    // ran, there is no work list: a tile's entries are its queries
    const bool all_queries = rows_before < 0;
    if (all_queries) rows_before = 0;
    auto tile_queries = [&](int t) { return static_cast<uint32_t>(min(kIcpThreads, prm.n - t * kIcpThreads)); };
    for (int tile0 = static_cast<int>(blockIdx.x); tile0 < ntiles; tile0 += 2 * static_cast<int>(gridDim.x)) {
        const int tile1 = tile0 + static_cast<int>(gridDim.x);
        const uint32_t c0 = all_queries ? tile_queries(tile0) : wk.refresh_count[tile0],
                       c1 = tile1 < ntiles ? (all_queries ? tile_queries(tile1) : wk.refresh_count[tile1]) : 0u;
        const uint32_t* const list0 = wk.refresh_list + static_cast<size_t>(tile0) * kIcpThreads;
        const uint32_t* const list1 = wk.refresh_list + static_cast<size_t>(tile1 < ntiles ? tile1 : tile0) * kIcpThreads;
        if (c0 + c1 <= 4 * kIcpWarps) {
            // a handful: one warp per query
            for (uint32_t it = warp; it < c0 + c1; it += kIcpWarps) {
                const uint32_t gi = all_queries ? static_cast<uint32_t>((it < c0 ? tile0 : tile1) * kIcpThreads) + (it < c0 ? it : it - c0)
                                                : (it < c0 ? list0[it] : list1[it - c0]);
                double sx, sy, sz, px, py, pz; uint4 m0; float4 prev; bool same_key; uint32_t qkey_lo, qkey_hi; size_t cbase;
                load_query(gi, sx, sy, sz, px, py, pz, m0, prev, same_key, qkey_lo, qkey_hi, cbase);  // (every lane the same query)
                WarmRefresh r;
                warm_refresh_warp(map, wk.cand, ccap, prm.warm_margin, 0, px, py, pz, m0, prev, same_key, cbase, s_run[warp], &r);
                if (lane == 0) commit(gi, r, sx, sy, sz, px, py, pz, qkey_lo, qkey_hi);
            }
        } else {
            // many (the first warm iteration of a call): one thread per query, tile after tile
            for (int h = 0; h < 2; ++h) {
                const uint32_t cnt = h ? c1 : c0;
                if (static_cast<uint32_t>(tid) < cnt) {
                    const uint32_t gi = all_queries ? static_cast<uint32_t>((h ? tile1 : tile0) * kIcpThreads + tid) : (h ? list1 : list0)[tid];
                    double sx, sy, sz, px, py, pz; uint4 m0; float4 prev; bool same_key; uint32_t qkey_lo, qkey_hi; size_t cbase;
                    load_query(gi, sx, sy, sz, px, py, pz, m0, prev, same_key, qkey_lo, qkey_hi, cbase);
                    WarmRefresh r;
                    warm_refresh(map, wk.cand + cbase, ccap, prm.warm_margin, m0, prev, same_key, px, py, pz, &r);
                    commit(gi, r, sx, sy, sz, px, py, pz, qkey_lo, qkey_hi);
                }
            }
        }
    }
    if (all_queries && prm.stats && tid == 0) {  // (the counters the reuse kernel would have kept)
        unsigned long long q = 0;
        for (int t = static_cast<int>(blockIdx.x); t < ntiles; t += static_cast<int>(gridDim.x)) q += tile_queries(t);
        atomicAdd(prm.stats + 1, q); atomicAdd(prm.stats + 20, q); atomicAdd(prm.stats + 21, q);
    }
    block_sum_into<NACC, METHOD == 0>(acc, s_red, s_sum);
    // the rows of the reuse kernel are folded into this grid's rows in parallel (block b takes the rows b, b + gridDim.x, ...
    // in that order), so that the last block's serial reduction is over gridDim.x rows, not over both grids
    if (tid < 29) {
        double v = s_sum[tid];
        for (int row = static_cast<int>(blockIdx.x); row < rows_before; row += static_cast<int>(gridDim.x)) v += __ldcg(wk.partials + static_cast<size_t>(row) * kAcc + tid);
        s_sum[tid] = v;
    }
    __syncthreads();
    finish_grid(s_sum, s_red, s_acc, &s_last, &s_solve, s_T, st, prm, wk.partials, wk.ticket, solve_here, rows_before, -1, rows_before);
}

// ======================================================================================================================
// search: VGICP
// ======================================================================================================================
__global__ void __launch_bounds__(kIcpThreads, 4)
icp_search_means_kernel(MapView map, const float* __restrict__ scan, const int* __restrict__ orig, IcpParams prm, const IcpState* __restrict__ st,
                        int* __restrict__ match) {
    __shared__ double s_T[12];
    pdl_launch_dependents();
    pdl_wait();
    if (st->done) return;
    const int tid = threadIdx.x;
    if (tid < 12) s_T[tid] = st->T[tid];
    __syncthreads();
    for (int i = blockIdx.x * kIcpThreads + tid; i < prm.n; i += gridDim.x * kIcpThreads) {
        const double sx = scan[3 * static_cast<size_t>(i)], sy = scan[3 * static_cast<size_t>(i) + 1], sz = scan[3 * static_cast<size_t>(i) + 2];
        const double px = row_apply_exact(s_T, 0, sx, sy, sz), py = row_apply_exact(s_T, 1, sx, sy, sz), pz = row_apply_exact(s_T, 2, sx, sy, sz);
        match[orig ? orig[i] : i] = nearest_mean_27(map, px, py, pz);
    }
}


// One thread per scan point.  partials[gridDim.x][kAcc]; the last block to finish
// sums them in a fixed order into st->acc and, when `solve_here`, runs the solve/update step.
template <int METHOD>
__global__ void __launch_bounds__(kIcpThreads, 2)
icp_accumulate_kernel(MapView map, const float* __restrict__ scan, IcpParams prm, IcpState* __restrict__ st, IcpWork wk, int solve_here) {
    constexpr int NACC = AccSize<METHOD>::value;
    const int* const __restrict__ match = wk.match;
    const float4* const __restrict__ win = wk.win;
    double* const __restrict__ partials = wk.partials;
    unsigned int* const __restrict__ ticket = wk.ticket;
    __shared__ double s_T[12], s_Tinv[12], s_Rinv[9];
    __shared__ double s_red[kIcpWarps][kAcc];
    __shared__ SolveScratch s_solve;
    __shared__ bool s_last;
    pdl_launch_dependents();
    pdl_wait();
    if (st->done) return;
    const int tid = threadIdx.x;
#ifdef ELM_PHASE_TIMING
    long long atick = clock64();
    if (prm.stats && tid == 0) atomicAdd(prm.stats + 15, 1ull);
#endif
    if (tid < 12) { s_T[tid] = st->T[tid]; s_Tinv[tid] = st->Tinv[tid]; }
    if (tid < 9) s_Rinv[tid] = st->Rinv[tid];
    __syncthreads();
    ELM_ATICK(9);

    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;

    static_assert(METHOD != 3, "AVGICP has its own kernel (icp_avgicp_kernel)");
    for (int i = blockIdx.x * kIcpThreads + tid; i < prm.n; i += gridDim.x * kIcpThreads) {
        const double sx = scan[3 * static_cast<size_t>(i)], sy = scan[3 * static_cast<size_t>(i) + 1], sz = scan[3 * static_cast<size_t>(i) + 2];
        const double px = row_apply_exact(s_T, 0, sx, sy, sz), py = row_apply_exact(s_T, 1, sx, sy, sz), pz = row_apply_exact(s_T, 2, sx, sy, sz);
        const int m = (METHOD == 0) ? 0 : match[i];  // (P2P needs only the streamed target)
        if (METHOD == 2) {
            double rec[12] = {0.0, 0.0, 0.0, 1, 0, 0, 0, 1, 0, 0, 0, 1};  // default CovStruct (mean 0, cov I)  (Q2, vhm.cpp:104)
            if (m >= 0) load_voxel_record(map, static_cast<uint32_t>(m), rec);
            if (sq3_exact(rec[0] - px, rec[1] - py, rec[2] - pz) < prm.max_dist2)  // vhm.cpp:129
                linearize_voxel_pair(acc, s_Tinv, s_Rinv, prm.th, sx, sy, sz, rec);
        } else {
            const float4 target = __ldg(win + i);  // coalesced: the search wrote the matched point itself
            linearize_point_pair<METHOD == 1 ? 1 : 0>(map, m, target, sx, sy, sz, px, py, pz, s_Tinv, s_Rinv, prm.th, prm.max_dist2, acc);
        }
    }

    __shared__ double s_sum[kAcc], s_acc[kAcc];
    if (tid < kAcc) s_sum[tid] = 0.0;
    __syncthreads();
    ELM_ATICK(10);
    block_sum_into<NACC, METHOD == 0>(acc, s_red, s_sum);
    ELM_ATICK(11);
    finish_grid(s_sum, s_red, s_acc, &s_last, &s_solve, s_T, st, prm, partials, ticket, solve_here);
}

// ======================================================================================================================
// AVGICP: GetCorrespondencesAllCov (vhm.cpp:153-206) + AlignCloudsLocalVoxelCov (reg.cpp:154-225) in one kernel
// ======================================================================================================================
// A warp takes 32 scan points at a time.
//   1  lane = POINT: transform, key, ONE directory lookup (two 32-byte buckets), the entry's 32-byte dir7 row = the voxel indices
//      of {centre, +x, -x, +y, -y, +z, -z} (vhm.cpp:224-230); the point (exact fp64 position + source coordinates) goes to
//      shared memory, its non-empty voxels are appended to the warp's pair list (ballot-free prefix over the lanes);
//   2  lane = PAIR: the warp's (point, voxel) pairs — 4.8 per point on the bench map, up to 224 — are shared out round-robin,
//      so every lane linearises the same number of pairs (before: 8 lanes per point, each repeating the transform and the
//      lookup, 7 of 8 lanes loading, idle lanes for every empty voxel).  A pair reads ONE 128-byte line (mean + covariance),
//      applies the distance gate on the exact mean (vhm.cpp:183) and accumulates.
// The emission order of the reference (voxel order per point) only matters for the summation order, which is compared with
// a tolerance; the per-lane assignment is fixed, so results are bit-reproducible from run to run.
__global__ void __launch_bounds__(kIcpThreads, 2)
icp_avgicp_kernel(MapView map, const float* __restrict__ scan, IcpParams prm, IcpState* __restrict__ st, IcpWork wk, int solve_here) {
    constexpr int NACC = 29;
    __shared__ double s_T[12], s_Tinv[12], s_Rinv[9];
    __shared__ double s_red[kIcpWarps][kAcc], s_sum[kAcc], s_acc[kAcc];
    __shared__ SolveScratch s_solve;
    __shared__ bool s_last;
    __shared__ double s_p[kIcpWarps][32][3];
    __shared__ float s_s[kIcpWarps][32][3];
    __shared__ uint32_t s_pair[kIcpWarps][7 * 32];  // lane << 25 | voxel index
    pdl_launch_dependents();
    pdl_wait();
    if (st->done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 12) { s_T[tid] = st->T[tid]; s_Tinv[tid] = st->Tinv[tid]; }
    if (tid < 9) s_Rinv[tid] = st->Rinv[tid];
    if (tid < kAcc) s_sum[tid] = 0.0;
    __syncthreads();
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    for (long long base = (static_cast<long long>(blockIdx.x) * kIcpWarps + warp) * 32; base < prm.n; base += static_cast<long long>(gridDim.x) * kIcpThreads) {
        const long long i = base + lane;
        int vox[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
        if (i < prm.n) {
            const float sxf = scan[3 * static_cast<size_t>(i)], syf = scan[3 * static_cast<size_t>(i) + 1], szf = scan[3 * static_cast<size_t>(i) + 2];
            const double sx = sxf, sy = syf, sz = szf;
            const double px = row_apply_exact(s_T, 0, sx, sy, sz), py = row_apply_exact(s_T, 1, sx, sy, sz), pz = row_apply_exact(s_T, 2, sx, sy, sz);
            s_p[warp][lane][0] = px; s_p[warp][lane][1] = py; s_p[warp][lane][2] = pz;
            s_s[warp][lane][0] = sxf; s_s[warp][lane][1] = syf; s_s[warp][lane][2] = szf;
            const int kx = voxel_floor(px, map), ky = voxel_floor(py, map), kz = voxel_floor(pz, map);
            uint2 centre;
            const int row = dir_lookup(map, kx, ky, kz, centre);
            if (row >= 0)
                asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(vox[0]), "=r"(vox[1]), "=r"(vox[2]), "=r"(vox[3]), "=r"(vox[4]), "=r"(vox[5]), "=r"(vox[6]), "=r"(vox[7])
                    : "l"(map.dir7 + static_cast<size_t>(row) * 8));
        }
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) cnt += vox[j] >= 0 ? 1 : 0;
        int off = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, off, o);
            if (lane >= o) off += v;
        }
        const int total = __shfl_sync(kFull, off, 31);
        off -= cnt;
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (vox[j] >= 0) s_pair[warp][off++] = (static_cast<uint32_t>(lane) << 25) | static_cast<uint32_t>(vox[j]);
        __syncwarp();
        for (int t = lane; t < total; t += 32) {
            const uint32_t pr = s_pair[warp][t];
            const int l = static_cast<int>(pr >> 25);
            double rec[12];
            load_voxel_record(map, pr & 0x1ffffffu, rec);
            const double px = s_p[warp][l][0], py = s_p[warp][l][1], pz = s_p[warp][l][2];
            if (sq3_exact(rec[0] - px, rec[1] - py, rec[2] - pz) < prm.max_dist2)  // vhm.cpp:183
                linearize_voxel_pair(acc, s_Tinv, s_Rinv, prm.th, static_cast<double>(s_s[warp][l][0]), static_cast<double>(s_s[warp][l][1]),
                                     static_cast<double>(s_s[warp][l][2]), rec);
        }
        __syncwarp();
    }
    block_sum_into<NACC, false>(acc, s_red, s_sum);
    finish_grid(s_sum, s_red, s_acc, &s_last, &s_solve, s_T, st, prm, wk.partials, wk.ticket, solve_here);
}

// ======================================================================================================================
// small fixed-size kernels
// ======================================================================================================================
__global__ void icp_begin_kernel(IcpState* st, Pose16 T0, unsigned int* ticket, unsigned long long* tile_ticket) {
    // every concurrent-refresh iteration of the call owns ITS pair of counters {chunks handed out, chunk rows completed}: kernels of
    // several iterations can be resident at once (programmatic dependent launch), a shared counter would mix their draws
    if (tile_ticket) for (int i = threadIdx.x; i < 2 * kMaxAsyncIterations; i += blockDim.x) tile_ticket[i] = 0ull;
    if (threadIdx.x == 0) {
        // (the tile-ticket counters are reset by every lane below)
        for (int i = 0; i < 16; ++i) st->T[i] = T0.m[i];
        refresh_inverses(st);
        for (int i = 0; i < 36; ++i) st->local_cov[i] = (i % 7 == 0) ? 1.0 : 0.0;  // reg.cpp:280
        st->iterations = 0; st->done = 0; st->overlap_fail = 0; st->comm_error = 0;
        *ticket = 0;
    }
}

__global__ void icp_solve_kernel(IcpState* st, IcpParams prm) {
    __shared__ SolveScratch s_solve;
    __shared__ double s_acc[kAcc], s_T[12];
    pdl_launch_dependents();
    pdl_wait();
    if (st->done) return;
    if (threadIdx.x < kAcc) s_acc[threadIdx.x] = st->acc[threadIdx.x];
    if (threadIdx.x < 12) s_T[threadIdx.x] = st->T[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) solve_step(st, prm, &s_solve, s_acc, s_T);
}

// ---- correspondence dump (test hook): turns the production search's match[] into (count, target) ---------------------
__global__ void __launch_bounds__(kIcpThreads)
icp_export_kernel(MapView map, const float* __restrict__ scan, const int* __restrict__ match, int n, const IcpState* __restrict__ st, int method,
                  double max_dist2, int* __restrict__ count, double* __restrict__ target) {
    const double* T = st->T;
    if (method == 3) {
        // AVGICP: every voxel of {c, +x, -x, +y, -y, +z, -z} whose mean is in range, emitted in that order (vhm.cpp:153-206)
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        const double sx = scan[3 * static_cast<size_t>(i)], sy = scan[3 * static_cast<size_t>(i) + 1], sz = scan[3 * static_cast<size_t>(i) + 2];
        const double px = row_apply_exact(T, 0, sx, sy, sz), py = row_apply_exact(T, 1, sx, sy, sz), pz = row_apply_exact(T, 2, sx, sy, sz);
        const int kx = voxel_floor(px, map), ky = voxel_floor(py, map), kz = voxel_floor(pz, map);
        double* t = target + static_cast<size_t>(i) * 21;
        int c = 0;
        uint2 centre;
        const int row = dir_lookup(map, kx, ky, kz, centre);
        for (int j = 0; j < 7 && row >= 0; ++j) {
            const int slot = __ldg(map.dir7 + static_cast<size_t>(row) * 8 + j);
            if (slot < 0) continue;
            double mx, my, mz;
            voxel_mean(map, static_cast<uint32_t>(slot), mx, my, mz);
            if (sq3_exact(mx - px, my - py, mz - pz) < max_dist2) { t[3 * c] = mx; t[3 * c + 1] = my; t[3 * c + 2] = mz; ++c; }
        }
        count[i] = c;
        for (; c < 7; ++c) { t[3 * c] = 0.0; t[3 * c + 1] = 0.0; t[3 * c + 2] = 0.0; }
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double sx = scan[3 * static_cast<size_t>(i)], sy = scan[3 * static_cast<size_t>(i) + 1], sz = scan[3 * static_cast<size_t>(i) + 2];
    const double px = row_apply_exact(T, 0, sx, sy, sz), py = row_apply_exact(T, 1, sx, sy, sz), pz = row_apply_exact(T, 2, sx, sy, sz);
    const int m = match[i];
    double tx = 0.0, ty = 0.0, tz = 0.0;  // Q2 default
    if (m >= 0) {
        if (method == 2) voxel_mean(map, static_cast<uint32_t>(m), tx, ty, tz);
        else { const float4 t = map.pts[m]; tx = t.x; ty = t.y; tz = t.z; }
    }
    const bool ok = sq3_exact(tx - px, ty - py, tz - pz) < max_dist2;
    count[i] = ok ? 1 : 0;
    target[3 * static_cast<size_t>(i)] = ok ? tx : 0.0;
    target[3 * static_cast<size_t>(i) + 1] = ok ? ty : 0.0;
    target[3 * static_cast<size_t>(i) + 2] = ok ? tz : 0.0;
}

// ======================================================================================================================
// launch wrappers
// ======================================================================================================================
// developer switch: ELM_NO_PDL=1 launches every kernel fully serialised
static const bool g_use_pdl = [] { const char* e = getenv("ELM_NO_PDL"); return !(e && e[0] == '1'); }();

// blocks of the warm reuse kernel: one query per thread, grid-stride loop beyond one resident wave
int icp_warm_grid(const IcpParams& prm, int num_sms) {
    const int blocks = (prm.n + kIcpThreads - 1) / kIcpThreads;
    const int cap = 4 * num_sms;
    return blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
}
// blocks of the warm refresh kernel (two resident per SM): enough threads for a bulk refresh of the whole scan, not more than
// one warp per query otherwise... a handful of queries leaves most of them idle, which costs nothing but their launch
int icp_warm_refresh_grid(const IcpParams& prm, int num_sms) {
    const int blocks = (prm.n + kIcpThreads - 1) / kIcpThreads;
    const int cap = 2 * num_sms;
    return blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
}

int icp_search_grid(const IcpParams& prm, int num_sms) {
    int blocks;
    if (prm.method <= 1) {
        blocks = (prm.n + kIcpThreads - 1) / kIcpThreads;
    } else {
        blocks = (prm.n + kIcpThreads - 1) / kIcpThreads;
    }
    const int cap = 4 * num_sms;
    return blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
}

int icp_accumulate_grid(const IcpParams& prm, int num_sms) {
    const long long threads = prm.n;
    long long blocks = (threads + kIcpThreads - 1) / kIcpThreads;
    const int cap = 2 * num_sms;
    return blocks < 1 ? 1 : (blocks > cap ? cap : static_cast<int>(blocks));
}

namespace {
// launch with programmatic stream serialization (see pdl_wait above)
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, int dyn_smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(static_cast<unsigned>(block));
    cfg.dynamicSmemBytes = static_cast<size_t>(dyn_smem);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace

cudaError_t launch_icp_begin(IcpState* st, const double T0[16], unsigned int* ticket, unsigned long long* tile_ticket, cudaStream_t s) {
    Pose16 p;
    for (int i = 0; i < 16; ++i) p.m[i] = T0[i];
    icp_begin_kernel<<<1, 32, 0, s>>>(st, p, ticket, tile_ticket);
    return cudaGetLastError();
}

// fuse: linearise + reduce (+ solve when solve_here) inside the search kernel (P2P / GICP only); then wk.match may be NULL.
cudaError_t launch_icp_search(const MapView& map, const float* scan, const int* orig, const IcpParams& prm, IcpState* st, const IcpWork& wk,
                              int grid, int prune, int fuse, int solve_here, cudaStream_t s) {
    cudaError_t e = cudaSuccess;
    if (prm.method <= 1) {
#define ELM_LAUNCH(C, F) e = launch_pdl(icp_search_points_kernel<C, F>, grid, kIcpThreads, 0, s, map, scan, orig, prm, st, wk, solve_here)
        if (!fuse) { if (prune) ELM_LAUNCH(true, -1); else ELM_LAUNCH(false, -1); }
        else if (prm.method == 0) { if (prune) ELM_LAUNCH(true, 0); else ELM_LAUNCH(false, 0); }
        else { if (prune) ELM_LAUNCH(true, 1); else ELM_LAUNCH(false, 1); }
#undef ELM_LAUNCH
    } else if (prm.method == 2) {
        e = launch_pdl(icp_search_means_kernel, grid, kIcpThreads, 0, s, map, scan, orig, prm, static_cast<const IcpState*>(st), wk.match);
    }
    return e != cudaSuccess ? e : cudaGetLastError();
}

// One warm iteration of P2P / GICP (search + linearisation + reduction + solve) = the reuse kernel, then the refresh kernel.
cudaError_t launch_icp_warm_reuse(const MapView& map, const float* scan, const IcpParams& prm, const IcpState* st, const IcpWork& wk, int reuse_grid,
                                  cudaStream_t s) {
    const cudaError_t e = prm.method == 0 ? launch_pdl(icp_warm_reuse_kernel<0>, reuse_grid, kIcpThreads, 0, s, map, scan, prm, st, wk)
                                          : launch_pdl(icp_warm_reuse_kernel<1>, reuse_grid, kIcpThreads, 0, s, map, scan, prm, st, wk);
    return e != cudaSuccess ? e : cudaGetLastError();
}
cudaError_t launch_icp_warm_refresh(const MapView& map, const float* scan, const IcpParams& prm, IcpState* st, const IcpWork& wk, int reuse_grid,
                                    int refresh_grid, int solve_here, cudaStream_t s) {
    const cudaError_t e = prm.method == 0
                              ? launch_pdl(icp_warm_refresh_kernel<0>, refresh_grid, kIcpThreads, 0, s, map, scan, prm, st, wk, reuse_grid, solve_here)
                              : launch_pdl(icp_warm_refresh_kernel<1>, refresh_grid, kIcpThreads, 0, s, map, scan, prm, st, wk, reuse_grid, solve_here);
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_icp_warm(const MapView& map, const float* scan, const IcpParams& prm, IcpState* st, const IcpWork& wk, int grid, int solve_here,
                            cudaStream_t s) {
    const cudaError_t e = prm.method == 0 ? launch_pdl(icp_warm_kernel<0>, grid, kIcpThreads, 0, s, map, scan, prm, st, wk, solve_here)
                                          : launch_pdl(icp_warm_kernel<1>, grid, kIcpThreads, 0, s, map, scan, prm, st, wk, solve_here);
    return e != cudaSuccess ? e : cudaGetLastError();
}

// Blocks of the concurrent refresh kernel.  A refresh block (128 threads x 128 registers) becomes resident beside THREE reuse blocks
// of an SM, not beside four, and the fold needs every refresh block: the last one to start gates the iteration.  With the 512 reuse
// blocks of a 131072-point scan 80 SMs hold three of them: 80 refresh blocks are all resident at once, and they leave room for the
// NEXT iteration's 512 reuse blocks to become resident before this iteration has finished (80 x 3 + 68 x 4 = 512).  Measured on B200
// (profiles/r02_ab_chain.txt): 80 blocks +1.4 % at 131072 points and +4.6 % at 16384 against 128; GICP is indifferent.
int icp_warm_refresh_async_grid(int num_sms, int want) {
    const int g = want > 0 ? want : 80;
    return num_sms < g ? num_sms : g;
}
cudaError_t launch_icp_warm_refresh_async(const MapView& map, const float* scan, const IcpParams& prm, IcpState* st, const IcpWork& wk, int reuse_grid,
                                          int grid, int solve_here, cudaStream_t s) {
    const cudaError_t e = prm.method == 0
                              ? launch_pdl(icp_warm_refresh_async_kernel<0>, grid, kAsyncThreads, 0, s, map, scan, prm, st, wk, reuse_grid, solve_here)
                              : launch_pdl(icp_warm_refresh_async_kernel<1>, grid, kAsyncThreads, 0, s, map, scan, prm, st, wk, reuse_grid, solve_here);
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_icp_accumulate(const MapView& map, const float* scan, const IcpParams& prm, IcpState* st, const IcpWork& wk, int solve_here,
                                  int grid, cudaStream_t s) {
    cudaError_t e = cudaSuccess;
    switch (prm.method) {
        case 0: e = launch_pdl(icp_accumulate_kernel<0>, grid, kIcpThreads, 0, s, map, scan, prm, st, wk, solve_here); break;
        case 1: e = launch_pdl(icp_accumulate_kernel<1>, grid, kIcpThreads, 0, s, map, scan, prm, st, wk, solve_here); break;
        case 2: e = launch_pdl(icp_accumulate_kernel<2>, grid, kIcpThreads, 0, s, map, scan, prm, st, wk, solve_here); break;
        default: e = launch_pdl(icp_avgicp_kernel, grid, kIcpThreads, 0, s, map, scan, prm, st, wk, solve_here); break;
    }
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_icp_solve(IcpState* st, const IcpParams& prm, cudaStream_t s) {
    const cudaError_t e = launch_pdl(icp_solve_kernel, 1, 32, 0, s, st, prm);
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_icp_export(const MapView& map, const float* scan, const int* match, int n, const IcpState* st, int method, double max_dist2,
                              int* count, double* target, cudaStream_t s) {
    const int blocks = (n + kIcpThreads - 1) / kIcpThreads;
    icp_export_kernel<<<blocks < 1 ? 1 : blocks, kIcpThreads, 0, s>>>(map, scan, match, n, st, method, max_dist2, count, target);
    return cudaGetLastError();
}

}  // namespace elm
