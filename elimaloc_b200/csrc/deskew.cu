// Point deskew on the GPU: the per-point body of PcmMatching::DeskewPointCloud
//   pcm_matching/src/pcm_matching.cpp  DeskewPoint :780-824, FindRotation :731-762, FindPosition :764-778
// (the tbb::parallel_for of :499-511).  float32 arithmetic exactly as the reference: the integrated-gyro table is double,
// every value is narrowed to float where the reference narrows it, products and sums are explicit round-to-nearest
// (no FMA contraction), sin/cos are evaluated in double and rounded to float (== the correctly rounded float result the
// host libm returns, except in vanishingly rare double-rounding cases — the tests allow 2 float ulps at |x| ~ 100 m).
// The table builders ImuDeskewInfo / OdomDeskewInfo (:533-729) walk ROS message queues and stay on the host.
#include "deskew.cuh"

namespace elm {

namespace {

constexpr int kSmemEntries = 1024;

__device__ __forceinline__ float f_sin(float x) { return static_cast<float>(sin(static_cast<double>(x))); }
__device__ __forceinline__ float f_cos(float x) { return static_cast<float>(cos(static_cast<double>(x))); }
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

__global__ void __launch_bounds__(256) deskew_points_kernel(const float* __restrict__ xyz, const float* __restrict__ rel_time, int n,
                                                            DeskewParams p, const double* __restrict__ table, float* __restrict__ out) {
    __shared__ double s_time[kSmemEntries], s_rx[kSmemEntries], s_ry[kSmemEntries], s_rz[kSmemEntries];
    const int entries = p.imu_pointer_cur + 1;
    const bool in_smem = entries <= kSmemEntries;
    const double* t_time = table;
    const double* t_rx = table + p.table_stride;
    const double* t_ry = table + 2 * p.table_stride;
    const double* t_rz = table + 3 * p.table_stride;
    if (in_smem && p.imu_available) {
        for (int i = threadIdx.x; i < entries; i += blockDim.x) { s_time[i] = t_time[i]; s_rx[i] = t_rx[i]; s_ry[i] = t_ry[i]; s_rz[i] = t_rz[i]; }
        __syncthreads();
        t_time = s_time; t_rx = s_rx; t_ry = s_ry; t_rz = s_rz;
    }
    const int cur = p.imu_pointer_cur;
    if (p.n_dev) n = min(n, *p.n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = xyz[3 * static_cast<size_t>(i)], y = xyz[3 * static_cast<size_t>(i) + 1], z = xyz[3 * static_cast<size_t>(i) + 2];
        if (!p.imu_available || !p.run_deskew) {  // :781, :512-525
            out[3 * static_cast<size_t>(i)] = x; out[3 * static_cast<size_t>(i) + 1] = y; out[3 * static_cast<size_t>(i) + 2] = z;
            continue;
        }
        const double d_rel_time = static_cast<double>(__fsub_rn(rel_time[i], p.rel_time_offset));  // (time -= front_time, :485)
        const double d_point_time = p.time_scan_cur + d_rel_time;  // :783
        const float rx_end = static_cast<float>(t_rx[cur]), ry_end = static_cast<float>(t_ry[cur]), rz_end = static_cast<float>(t_rz[cur]);
        // FindRotation: first table entry later than the point (linear scan like the reference; the table is short)
        int front = 0;
        while (front < cur) {
            if (d_point_time < t_time[front]) break;
            ++front;
        }
        float rx, ry, rz;
        if (d_point_time > t_time[front] || front == 0) {
            rx = static_cast<float>(t_rx[front]); ry = static_cast<float>(t_ry[front]); rz = static_cast<float>(t_rz[front]);
        } else {
            const int back = front - 1;
            const double span = __dsub_rn(t_time[front], t_time[back]);
            const double rf = __ddiv_rn(__dsub_rn(d_point_time, t_time[back]), span);
            const double rb = __ddiv_rn(__dsub_rn(t_time[front], d_point_time), span);
            rx = static_cast<float>(__dadd_rn(__dmul_rn(t_rx[front], rf), __dmul_rn(t_rx[back], rb)));
            ry = static_cast<float>(__dadd_rn(__dmul_rn(t_ry[front], rf), __dmul_rn(t_ry[back], rb)));
            rz = static_cast<float>(__dadd_rn(__dmul_rn(t_rz[front], rf), __dmul_rn(t_rz[back], rb)));
        }
        // FindPosition
        float px = 0.f, py = 0.f;
        if (p.odom_available) {
            const float ratio = static_cast<float>(__ddiv_rn(d_rel_time, __dsub_rn(p.time_scan_end, p.time_scan_cur)));
            px = mul(ratio, p.odom_incre_x); py = mul(ratio, p.odom_incre_y);
        }
        const float roll = sub(rx, rx_end), pitch = sub(ry, ry_end), yaw = sub(rz, rz_end);  // :796-799
        const float tx = sub(px, p.odom_incre_x), ty = sub(py, p.odom_incre_y);              // :801-803
        const float tz = sub(rz, p.odom_incre_z);                                            // :804 (Q3: rotation z, not position z)
        // pcl::getTransformation(tx, ty, tz, roll, pitch, yaw)
        const float A = f_cos(yaw), B = f_sin(yaw), C = f_cos(pitch), D = f_sin(pitch), E = f_cos(roll), F = f_sin(roll);
        const float DE = mul(D, E), DF = mul(D, F);
        const float m00 = mul(A, C), m01 = sub(mul(A, DF), mul(B, E)), m02 = add(mul(B, F), mul(A, DE));
        const float m10 = mul(B, C), m11 = add(mul(A, E), mul(B, DF)), m12 = sub(mul(B, DE), mul(A, F));
        const float m20 = -D, m21 = mul(C, F), m22 = mul(C, E);
        out[3 * static_cast<size_t>(i)] = add(add(add(mul(m00, x), mul(m01, y)), mul(m02, z)), tx);  // :815-820
        out[3 * static_cast<size_t>(i) + 1] = add(add(add(mul(m10, x), mul(m11, y)), mul(m12, z)), ty);
        out[3 * static_cast<size_t>(i) + 2] = add(add(add(mul(m20, x), mul(m21, y)), mul(m22, z)), tz);
    }
}

}  // namespace

cudaError_t launch_deskew_points(const float* xyz, const float* rel_time, int n, const DeskewParams& p, const double* table, float* out,
                                 int num_sms, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int blocks = (n + 255) / 256;
    if (blocks > 8 * num_sms) blocks = 8 * num_sms;
    deskew_points_kernel<<<blocks, 256, 0, s>>>(xyz, rel_time, n, p, table, out);
    return cudaGetLastError();
}

}  // namespace elm
