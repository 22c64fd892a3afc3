// Micro-benchmark: how fast can one SM gather many SMALL contiguous runs (a voxel column's points, 100-1500 B) from random
// places of a large array into shared memory?  (a) one cp.async.bulk (1-D TMA, SASS UBLKCP) per thread and run,
// (b) warp-cooperative 16-byte cp.async (LDGSTS), one coalesced instruction per run, (c) plain per-thread LDG.128 loop
// that consumes the points directly (what the search kernel did).  Prints runs/us/SM and GB/s for each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_gather profiles/micro/tma_gather.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kThreads = 256;

// mode 0: TMA per thread; 1: warp-cooperative LDGSTS; 2: direct LDG streaming
template <int MODE>
__global__ void __launch_bounds__(kThreads) gather(const float4* __restrict__ pts, const uint32_t* __restrict__ starts, int run_pts, int rounds,
                                                   int runs_per_round, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    float4* pool = reinterpret_cast<float4*>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, kThreads); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    float acc = 0.f;
    uint32_t parity = 0;
    for (int r = 0; r < rounds; ++r) {
        const uint32_t* st = starts + (static_cast<size_t>(blockIdx.x) * rounds + r) * kThreads;
        const bool active = tid < runs_per_round;
        const uint32_t s = st[tid];
        if (MODE == 0) {
            if (active) { mbar_expect_tx(&bar, run_pts * 16); tma_load_1d(pool + tid * run_pts, pts + s, run_pts * 16, &bar); }
            else mbar_expect_tx(&bar, 0);
            mbar_wait(&bar, parity); parity ^= 1;
        } else if (MODE == 1) {
            for (int q = 0; q < 32; ++q) {
                const uint32_t sq = __shfl_sync(0xffffffffu, s, q);
                const int owner = warp * 32 + q;
                if (owner < runs_per_round)
                    for (int o = lane; o < run_pts; o += 32)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(pool + owner * run_pts + o)), "l"(pts + sq + o) : "memory");
            }
            asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        } else if (MODE == 3) {  // warp-cooperative coalesced LDG.128 -> STS.128, 8 runs in flight per lane
            for (int q0 = 0; q0 < 32; q0 += 8) {
                uint32_t sq[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) sq[u] = __shfl_sync(0xffffffffu, s, q0 + u);
                for (int o = lane; o < run_pts; o += 32) {
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        v[u] = (warp * 32 + q0 + u < runs_per_round) ? __ldg(pts + sq[u] + o) : make_float4(0, 0, 0, 0);
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (warp * 32 + q0 + u < runs_per_round) pool[(warp * 32 + q0 + u) * run_pts + o] = v[u];
                }
            }
            __syncwarp();
        }
        if (active) {
            if (MODE == 2) {
#pragma unroll 4
                for (int o = 0; o < run_pts; ++o) { const float4 q = __ldg(pts + s + o); acc += q.x * q.y + q.z; }
            } else {
#pragma unroll 4
                for (int o = 0; o < run_pts; ++o) { const float4 q = pool[tid * run_pts + o]; acc += q.x * q.y + q.z; }
            }
        }
        if (MODE == 0 || MODE == 1) __syncthreads();
        if (MODE == 3) __syncwarp();
    }
    out[blockIdx.x * kThreads + tid] = acc;
}

int main(int argc, char** argv) {
    const size_t npts = 9'000'000;
    std::vector<float4> h(npts);
    for (size_t i = 0; i < npts; ++i) h[i] = make_float4(float(i & 255), 1.f, 2.f, 0.f);
    float4* d_pts; cudaMalloc(&d_pts, npts * sizeof(float4)); cudaMemcpy(d_pts, h.data(), npts * sizeof(float4), cudaMemcpyHostToDevice);
    const int blocks_per_sm_list[] = {3, 4};
    const int rounds = 16;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float* d_out; cudaMalloc(&d_out, sizeof(float) * kThreads * sms * 8);
    for (int run_pts : {9, 12, 27}) {
        for (int bps : blocks_per_sm_list) {
            const int grid = sms * bps;
            int pool_bytes = 200 * 1024 / bps; if (pool_bytes > 200 * 1024) pool_bytes = 200 * 1024;
            int runs_per_round = pool_bytes / (run_pts * 16); if (runs_per_round > kThreads) runs_per_round = kThreads;
            pool_bytes = runs_per_round * run_pts * 16;
            std::vector<uint32_t> st(static_cast<size_t>(grid) * rounds * kThreads);
            uint64_t rng = 88172645463325252ull;
            for (auto& v : st) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; v = static_cast<uint32_t>(rng % (npts - 128)); }
            uint32_t* d_st; cudaMalloc(&d_st, st.size() * 4); cudaMemcpy(d_st, st.data(), st.size() * 4, cudaMemcpyHostToDevice);
            for (int mode = 0; mode < 4; ++mode) {
                auto launch = [&]() {
                    if (mode == 0) { cudaFuncSetAttribute(gather<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, pool_bytes); gather<0><<<grid, kThreads, pool_bytes>>>(d_pts, d_st, run_pts, rounds, runs_per_round, d_out); }
                    if (mode == 1) { cudaFuncSetAttribute(gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, pool_bytes); gather<1><<<grid, kThreads, pool_bytes>>>(d_pts, d_st, run_pts, rounds, runs_per_round, d_out); }
                    if (mode == 2) gather<2><<<grid, kThreads, 0>>>(d_pts, d_st, run_pts, rounds, runs_per_round, d_out);
                    if (mode == 3) { cudaFuncSetAttribute(gather<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, pool_bytes); gather<3><<<grid, kThreads, pool_bytes>>>(d_pts, d_st, run_pts, rounds, runs_per_round, d_out); }
                };
                launch(); cudaDeviceSynchronize();
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0); for (int i = 0; i < 5; ++i) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
                const double runs = double(grid) * rounds * runs_per_round;
                cudaError_t err = cudaGetLastError();
                printf("run_pts %3d blocks/SM %d runs/round %3d mode %s: %8.1f us  %7.2f runs/us/SM  %7.1f GB/s %s\n", run_pts, bps, runs_per_round,
                       mode == 0 ? "TMA   " : (mode == 1 ? "LDGSTS" : (mode == 2 ? "LDG   " : "COOP  ")), ms * 1e3, runs / (ms * 1e3) / sms, runs * run_pts * 16 / (ms * 1e-3) / 1e9,
                       err == cudaSuccess ? "" : cudaGetErrorString(err));
            }
            cudaFree(d_st);
        }
    }
    return 0;
}
