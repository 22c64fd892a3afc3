// Micro-benchmark: latency of DEPENDENT random 32-byte loads under the occupancy of the search kernels (one wave of 512 blocks
// x 256 threads on 148 SMs), for an L2-resident and a DRAM-resident array, 1 or 3 independent loads per step, all or a
// quarter of the lanes active.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dep_gather dep_gather.cu && ./dep_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ldg256(const uint4* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
template <int W, int ACTIVE>
__global__ void __launch_bounds__(256, 4) chase(const uint4* __restrict__ arr, uint32_t mask, int steps, uint32_t* out, long long* cyc) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if ((threadIdx.x & 31) >= ACTIVE) return;
    uint32_t idx = tid * 2654435761u;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        uint4 a[W], b[W];
#pragma unroll
        for (int w = 0; w < W; ++w) ldg256(arr + 2 * (((idx + w * 0x9E3779B9u) >> 3) & mask), a[w], b[w]);
#pragma unroll
        for (int w = 0; w < W; ++w) acc += a[w].x ^ b[w].w;
        idx = idx * 1664525u + 1013904223u + (acc & 1u);   // depends on the loaded data
    }
    const long long t1 = clock64();
    out[tid] = acc;
    if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)cyc, (unsigned long long)(t1 - t0));
}
template <int W, int ACTIVE>
void run(const char* name, const uint4* arr, uint32_t mask, uint32_t* out, long long* cyc) {
    const int steps = 16, blocks = 512;
    cudaMemset(cyc, 0, 8);
    chase<W, ACTIVE><<<blocks, 256>>>(arr, mask, steps, out, cyc);  // warm-up (fills L2 when it fits)
    cudaMemset(cyc, 0, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    chase<W, ACTIVE><<<blocks, 256>>>(arr, mask, steps, out, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_step = double(h) / (blocks * 8.0) / steps;
    const double sectors = double(blocks) * 256 * ACTIVE / 32 * steps * W;
    printf("%-40s W=%d active=%2d : %7.0f cycles per dependent step, kernel %.1f us, %.2f TB/s of sectors\n", name, W, ACTIVE, per_step, ms * 1e3,
           sectors * 32 / (ms * 1e-3) / 1e12);
}
int main() {
    uint4* small; uint4* big; uint32_t* out; long long* cyc;
    const size_t nsmall = 32u << 20, nbig = 1024u << 20;  // bytes
    cudaMalloc(&small, nsmall); cudaMalloc(&big, nbig); cudaMalloc(&out, 512 * 256 * 4); cudaMalloc(&cyc, 8);
    cudaMemset(small, 1, nsmall); cudaMemset(big, 1, nbig);
    const uint32_t msmall = nsmall / 32 - 1, mbig = nbig / 32 - 1;
    run<1, 32>("L2-resident 32 MB", small, msmall, out, cyc);
    run<3, 32>("L2-resident 32 MB", small, msmall, out, cyc);
    run<1, 8>("L2-resident 32 MB", small, msmall, out, cyc);
    run<3, 8>("L2-resident 32 MB", small, msmall, out, cyc);
    run<1, 32>("DRAM 1 GB", big, mbig, out, cyc);
    run<3, 32>("DRAM 1 GB", big, mbig, out, cyc);
    run<1, 8>("DRAM 1 GB", big, mbig, out, cyc);
    run<3, 8>("DRAM 1 GB", big, mbig, out, cyc);
    return 0;
}
