// C ABI of elimaloc_b200 (include/elimaloc_b200.h): the device-resident VoxelHashMap, the Registration handle that owns
// the ICP loop, and the NCCL plumbing for the scan-sharded multi-GPU mode.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/elimaloc_b200.h"
#include "host_map.hpp"
#include "deskew.cuh"
#include "deskew_tables.hpp"
#include "ekf.cuh"
#include "icp_kernels.cuh"
#include "map_build.cuh"
#include "pcd_reader.hpp"
#include "scan_prep.cuh"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

// Exception barrier of the C ABI: every extern "C" entry point is a function-try-block ending in ELM_API_CATCH, so that
// an allocation failure (std::bad_alloc, std::length_error from a huge or corrupt input) becomes a status code instead of
// terminating the caller's process — the header promises "never throws".
#define ELM_API_CATCH                                                                                            \
    catch (const std::bad_alloc&) { return fail(ELM_ERR_INVALID, "out of host memory"); }                        \
    catch (const std::exception& e__) { return fail(ELM_ERR_INVALID, std::string("internal error: ") + e__.what()); } \
    catch (...) { return fail(ELM_ERR_INVALID, "internal error"); }

#define ELM_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            return fail(ELM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));              \
        }                                                                                                \
    } while (0)

template <class T>
cudaError_t upload(T** dptr, const void* src, size_t count) {
    if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dptr), count * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dptr, src, count * sizeof(T), cudaMemcpyHostToDevice);
}

// ---- NCCL through dlopen: the single-GPU path has no link-time dependency on it ---------------------------------
struct NcclApi {
    struct Uid { char b[128]; };  // ncclUniqueId is passed by value: 128 bytes
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Uid, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (h) return true;
        // single-node path: keep NCCL's bootstrap on the loopback interface unless the user chose otherwise (interface
        // probing / reverse DNS in containers can take minutes)
        setenv("NCCL_SOCKET_IFNAME", "lo", 0);
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return false;
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(h, "ncclAllReduce"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        return GetUniqueId && CommInitRank && AllReduce && CommDestroy && GetErrorString;
    }
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;      // ncclSum

}  // namespace

// ======================================================================================================================
struct elm_map {
    elm::HostMap host;
    int device = -1;  // -1: host-only map (builder tests without a GPU)
    // AddPoints / CalVoxelCovAll / CalPointCovAll on the GPU (map_build.cu; bit-identical to the host builder) when the map lives
    // on a device; ELM_HOST_BUILD=1 or elm_map_set_gpu_build(map, 0) keeps the host builder
    int gpu_build = [] { const char* e = getenv("ELM_HOST_BUILD"); return (e && e[0] == '1') ? 0 : 1; }();
    double build_ms[3] = {0.0, 0.0, 0.0};  // last AddPoints / CalVoxelCovAll / CalPointCovAll: milliseconds in the builder proper
    uint4* d_dslots = nullptr;
    uint32_t* d_drows = nullptr;
    float4* d_pts = nullptr;
    double* d_prec = nullptr;
    double* d_vrec = nullptr;                 // 128-byte record per voxel: mean[3] cov[9] pad[4]
    unsigned long long* d_vcand8 = nullptr;   // VGICP candidate records (voxel_key.hpp: pack_vcand)
    int* d_dir7 = nullptr;

    elm::MapView view() const {
        elm::MapView v;
        v.dslots = d_dslots; v.drows = d_drows; v.bmask = host.dir_bmask; v.pts = d_pts; v.prec = d_prec; v.vrec = d_vrec; v.vcand8 = d_vcand8; v.dir7 = d_dir7;
        v.voxel_size = host.voxel_size;
        int e2 = 0;
        v.inv_voxel_size = (std::frexp(host.voxel_size, &e2) == 0.5) ? 1.0 / host.voxel_size : 0.0;
        return v;
    }
    int publish_points() {
        if (device < 0) return ELM_OK;
        ELM_CUDA(cudaSetDevice(device));
        const size_t P = host.P();
        std::vector<float4> p4(P + 1);  // one element of padding: the search reads aligned 32-byte pairs
        p4[P] = float4{0.f, 0.f, 0.f, 0.f};
        // device order = octant order inside every voxel (host_map.hpp); w = bits of the CANONICAL index, the rank that
        // breaks exact distance ties the way the reference's visit order does (vhm.cpp:45)
        for (size_t d = 0; d < P; ++d) {
            const size_t i = host.dev_order[d];
            p4[d].x = host.pxyz[3 * i]; p4[d].y = host.pxyz[3 * i + 1]; p4[d].z = host.pxyz[3 * i + 2];
            p4[d].w = __int_as_float_host(static_cast<uint32_t>(i));
        }
        ELM_CUDA(upload(&d_pts, p4.data(), P + 1));
        static_assert(sizeof(elm::DirSlot) == sizeof(uint4), "directory layout");
        ELM_CUDA(upload(&d_dslots, reinterpret_cast<const uint4*>(host.dir_slots.data()), host.dir_slots.size()));
        ELM_CUDA(upload(&d_drows, host.dir_rows.data(), host.dir_rows.size()));
        if (d_prec) { cudaFree(d_prec); d_prec = nullptr; }
        if (d_vrec) { cudaFree(d_vrec); d_vrec = nullptr; }
        if (d_vcand8) { cudaFree(d_vcand8); d_vcand8 = nullptr; }
        if (d_dir7) { cudaFree(d_dir7); d_dir7 = nullptr; }
        return ELM_OK;
    }
    static float __int_as_float_host(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
    int publish_voxel_cov() {
        if (device < 0) return ELM_OK;
        ELM_CUDA(cudaSetDevice(device));
        const size_t V = host.V();
        std::vector<double> rec(16 * V, 0.0);  // one 128-byte line per voxel: the accumulation reads mean + covariance in ONE DRAM burst
        for (size_t v = 0; v < V; ++v) {
            for (int k = 0; k < 3; ++k) rec[16 * v + k] = host.vmean[3 * v + k];
            for (int k = 0; k < 9; ++k) rec[16 * v + 3 + k] = host.vcov[9 * v + k];
        }
        ELM_CUDA(upload(&d_vrec, rec.data(), 16 * V));
        // candidate lists + the row headers that point into them
        ELM_CUDA(upload(&d_vcand8, reinterpret_cast<const unsigned long long*>(host.vcand8.data()), host.vcand8.size()));
        ELM_CUDA(upload(&d_dir7, host.dir7.data(), host.dir7.size()));
        ELM_CUDA(upload(&d_drows, host.dir_rows.data(), host.dir_rows.size()));
        return ELM_OK;
    }
    int publish_point_cov() {
        if (device < 0) return ELM_OK;
        ELM_CUDA(cudaSetDevice(device));
        const size_t P = host.P();
        std::vector<double> rec(16 * P, 0.0);
        for (size_t d = 0; d < P; ++d) {  // same device order as the points
            const size_t p = host.dev_order[d];
            for (int k = 0; k < 3; ++k) rec[16 * d + k] = host.pmean[3 * p + k];
            for (int k = 0; k < 9; ++k) rec[16 * d + 3 + k] = host.pcov[9 * p + k];
            for (int k = 0; k < 3; ++k) rec[16 * d + 12 + k] = host.pnormal[3 * p + k];
        }
        ELM_CUDA(upload(&d_prec, rec.data(), 16 * P));
        return ELM_OK;
    }
    ~elm_map() {
        if (device >= 0) {
            cudaSetDevice(device);
            cudaFree(d_dslots); cudaFree(d_drows); cudaFree(d_pts); cudaFree(d_prec); cudaFree(d_vrec); cudaFree(d_vcand8); cudaFree(d_dir7);
        }
    }
};

struct elm_registration {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    elm::IcpState* d_state = nullptr;
    elm::IcpState* h_state = nullptr;  // pinned
    double* d_partials = nullptr;
    int partial_rows = 0;
    int* d_match = nullptr;
    float4* d_win = nullptr;   // matched map point per scan point (streamed by the accumulation)
    uint4* d_memo = nullptr;   // warm start of the next iteration's search
    uint32_t* d_ncand = nullptr;
    float4* d_cand = nullptr;  // per-query candidate lists of the warm search (icp_device.cuh)
    uint32_t* d_refresh = nullptr;  // work list of the warm refresh kernel: match_cap entries (a segment per tile) + match_cap / 256 counts
    unsigned long long* d_tile_flag = nullptr;    // concurrent refresh: per-tile {epoch | stragglers}, match_cap / 256 + 1 entries
    double* d_tile_rows = nullptr;                // ... per-tile sums, (match_cap / 256 + 1) x 32
    unsigned long long* d_tile_ticket = nullptr;  // ... tiles handed out since the call began
    unsigned int warm_epoch = 0;                  // epoch of the last warm iteration enqueued on this handle
    int async_iterations = 0;                     // concurrent-refresh iterations enqueued since the call began (each owns a pair of tile counters)
    int cand_cap = 32;
    size_t match_cap = 0;
    int warm = 1;              // P2P / GICP: iterations after the first start their search from the previous match (same result)
    // How the warm iterations after the first of a call run (same results):
    //   0 "pair"    icp_warm_reuse_kernel, then icp_warm_refresh_kernel
    //   1 "single"  ONE kernel, icp_warm_kernel (stragglers refreshed in place by their own warp)
    //   2 "async"   icp_warm_reuse_kernel with icp_warm_refresh_async_kernel running BESIDE it (the stragglers leave the critical path)
    // Measured on B200 (profiles/r02_ab_warm_modes.txt): async 35.2k / 21.5k iterations/s (P2P / GICP), single 32.4k / 20.6k, pair 31.1k / 18.5k.
    // Default (-1): async.  ELM_WARM_MODE=pair|single|async forces one (A/B switch).
    int warm_mode = [] {
        const char* e = getenv("ELM_WARM_MODE");
        if (!e) return -1;
        return e[0] == 'p' ? 0 : (e[0] == 's' ? 1 : (e[0] == 'a' ? 2 : -1));
    }();
    int async_grid = [] {  // blocks of the concurrent refresh kernel (A/B switch ELM_ASYNC_GRID; 0 = the built-in default)
        const char* e = getenv("ELM_ASYNC_GRID");
        const int v = e ? atoi(e) : 0;
        return v > 0 ? v : 0;
    }();
    // spatially binned copy of the scan for the search kernels (scan_sort.cu)
    float* d_sorted = nullptr;
    int* d_orig = nullptr;
    uint32_t* d_bin = nullptr;
    uint32_t* d_hist = nullptr;
    size_t sort_cap = 0;
    int fuse = 0;             // 1: P2P / GICP linearise + reduce + solve inside the search kernel (measured slower: spills)
    bool keep_match = false;  // also write match[] in the fused mode (off: nobody reads it)
    int binning = 0;          // 1 = search the scan in spatially binned order (measured: no gain on B200)
    bool use_sorted = false;  // the current enqueue searches d_sorted / d_orig
    bool sorted_all = false;  // ... and accumulates in that order too (RunRegister: only sums leave the loop, no per-point output)
    unsigned int* d_ticket = nullptr;
    unsigned long long* d_stats = nullptr;  // [visited map points, queries] when stats are on
    bool stats_on = false;
    int prune = 1;  // exact pruning of voxels that cannot hold the nearest neighbour (0 = visit all 27 like the reference)
    float* d_scan = nullptr;
    size_t scan_cap = 0;
    int* d_count = nullptr;
    double* d_target = nullptr;
    size_t hook_cap = 0;
    int64_t launches = 0;
    // optional per-kernel timing of the linearisation kernel (cudaEvents on the launch stream)
    bool profiling = false;
    std::vector<cudaEvent_t> ev;
    int ev_used = 0;
    double prof_search_ms = 0.0, prof_accum_ms = 0.0;
    int64_t prof_launches = 0;
    // the same split by kind of iteration: 0 cold (search + accumulate), 1 first warm (reuse + bulk refresh), 2 later warm
    std::vector<int> ev_kind;
    double prof_kind_ms[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    int64_t prof_kind_n[3] = {0, 0, 0};
    int warm_iterations_enqueued = 0;  // warm iterations since the last cold one (of the current call)
    // last enqueue
    bool pending = false, trivial = false;  // trivial: empty map or n == 0 -> no kernels ran
    elm_reg_config cfg{};
    double T_init[16];
    // deskew scratch
    double* d_dtable = nullptr;
    double* h_dtable = nullptr;  // pinned staging of the four table rows
    int dtable_cap = 0;
    float* d_dsk_in = nullptr;   // xyz | rel_time
    float* d_dsk_out = nullptr;
    size_t dsk_cap = 0;
    // scan pre-processing scratch (scan_prep.cu)
    elm::ScanPrepScratch prep{};
    size_t prep_cap = 0;
    int* d_prep_n = nullptr;
    float* d_prep_in = nullptr;    // host-buffer entry point: xyz | aux in, xyz | aux | index out
    float* d_prep_out = nullptr;
    int* d_prep_idx = nullptr;
    size_t prep_io_cap = 0;
    // NCCL
    void* comm = nullptr;
    int rank = 0, world = 1;
    // peer-memory exchange (CUDA IPC mailboxes, see icp_device.cuh)
    elm::PeerMailbox* d_mailbox = nullptr;
    elm::PeerComm peer{};          // peer.world > 0: the accumulate kernel's last block all-reduces over the ranks itself
    void* peer_opened[elm::kMaxPeers] = {};
    bool sharded() const { return comm != nullptr || peer.world > 0; }
    elm::IcpWork work() const { return elm::IcpWork{d_match, d_win, d_memo, match_cap, d_ncand, d_cand, cand_cap, d_refresh, d_refresh ? d_refresh + match_cap : nullptr, d_partials, d_ticket,
                                                   d_tile_flag, d_tile_rows, d_tile_ticket, 0u}; }
    double warm_margin_vox = 0.08;  // refresh margin of the warm search in voxel sizes

    ~elm_registration() {
        cudaSetDevice(device);
        if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
        cudaFree(prep.tkeys); cudaFree(prep.tmin); cudaFree(prep.keep); cudaFree(prep.block_count); cudaFree(prep.block_offset);
        cudaFree(prep.error); cudaFree(d_prep_n); cudaFree(d_prep_in); cudaFree(d_prep_out); cudaFree(d_prep_idx);
        for (void* p : peer_opened) if (p) cudaIpcCloseMemHandle(p);
        cudaFree(d_mailbox);
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        cudaFree(d_state); cudaFreeHost(h_state); cudaFree(d_partials); cudaFree(d_match); cudaFree(d_win); cudaFree(d_memo); cudaFree(d_ncand); cudaFree(d_cand); cudaFree(d_refresh); cudaFree(d_ticket);
        cudaFree(d_tile_flag); cudaFree(d_tile_rows); cudaFree(d_tile_ticket);
        cudaFree(d_stats);
        cudaFree(d_dtable); cudaFreeHost(h_dtable); cudaFree(d_dsk_in); cudaFree(d_dsk_out);
        cudaFree(d_sorted); cudaFree(d_orig); cudaFree(d_bin); cudaFree(d_hist); cudaFree(d_scan); cudaFree(d_count); cudaFree(d_target);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

struct elm_ekf {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    elm_ekf_config cfg{};
    elm_ekf_state* d_state = nullptr;
    elm_ekf_state* h_state = nullptr;  // pinned
    elm::EkfRing ring{nullptr, nullptr, elm::kEkfRingCap};  // PublishInThread's deque of EgoStates in HBM (optional)
    bool ring_on = false;
    ~elm_ekf() {
        cudaSetDevice(device);
        cudaFree(ring.e); cudaFree(ring.meta);
        cudaFree(d_state); cudaFreeHost(h_state);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

namespace {

int ensure_scan(elm_registration* r, size_t n) {
    if (n > r->scan_cap) {
        if (r->d_scan) cudaFree(r->d_scan);
        r->d_scan = nullptr;
        const size_t cap = (n + 1023) / 1024 * 1024;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_scan), cap * 3 * sizeof(float)));
        r->scan_cap = cap;
    }
    return ELM_OK;
}

int ensure_partials(elm_registration* r, int rows) {
    if (rows > r->partial_rows) {
        if (r->d_partials) cudaFree(r->d_partials);
        r->d_partials = nullptr;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_partials), static_cast<size_t>(rows) * elm::kAcc * sizeof(double)));
        r->partial_rows = rows;
    }
    return ELM_OK;
}

int ensure_match(elm_registration* r, size_t n) {
    if (n > r->match_cap) {
        cudaFree(r->d_match); cudaFree(r->d_win); cudaFree(r->d_memo); cudaFree(r->d_ncand); cudaFree(r->d_cand); cudaFree(r->d_refresh);
        cudaFree(r->d_tile_flag); cudaFree(r->d_tile_rows);
        r->d_match = nullptr; r->d_win = nullptr; r->d_memo = nullptr; r->d_ncand = nullptr; r->d_cand = nullptr; r->d_refresh = nullptr;
        r->d_tile_flag = nullptr; r->d_tile_rows = nullptr;
        r->match_cap = 0;
        const size_t cap = (n + 1023) / 1024 * 1024;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_match), cap * sizeof(int)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_win), cap * sizeof(float4)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_memo), 2 * cap * sizeof(uint4)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_ncand), cap * sizeof(uint32_t)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_cand), static_cast<size_t>(r->cand_cap) * cap * sizeof(float4)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_refresh), (cap + cap / 256 + 1) * sizeof(uint32_t)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_tile_flag), (cap / 256 + 1) * sizeof(unsigned long long)));
        ELM_CUDA(cudaMemsetAsync(r->d_tile_flag, 0, (cap / 256 + 1) * sizeof(unsigned long long), r->stream));  // epoch 0 = never published
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_tile_rows), (cap / 256 + 1) * elm::kAcc * sizeof(double)));
        r->match_cap = cap;
    }
    return ELM_OK;
}

int ensure_sort(elm_registration* r, size_t n) {
    if (n > r->sort_cap) {
        cudaFree(r->d_sorted); cudaFree(r->d_orig); cudaFree(r->d_bin);
        r->d_sorted = nullptr; r->d_orig = nullptr; r->d_bin = nullptr;
        const size_t cap = (n + 1023) / 1024 * 1024;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_sorted), cap * 3 * sizeof(float)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_orig), cap * sizeof(int)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_bin), cap * sizeof(uint32_t)));
        r->sort_cap = cap;
    }
    if (!r->d_hist) ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_hist), (1u << 18) * sizeof(uint32_t)));
    return ELM_OK;
}

// Bin the scan under pose T (once per call); afterwards the search kernels read d_sorted / d_orig.
int enqueue_binning(elm_registration* r, const elm_map* map, const float* d_scan, size_t n, const double T[16], int method) {
    r->use_sorted = false;
    r->sorted_all = false;
    if (!r->binning || method == ELM_AVGICP || n < 2048) return ELM_OK;
    int rc = ensure_sort(r, n);
    if (rc) return rc;
    ELM_CUDA(elm::launch_scan_binning(d_scan, static_cast<int>(n), T, map->host.voxel_size, r->d_bin, r->d_hist, r->d_sorted, r->d_orig, r->stream));
    r->launches += 3;
    r->use_sorted = true;
    return ELM_OK;
}

int check_method(const elm_map* map, const elm_reg_config* cfg) {
    if (cfg->icp_method < ELM_P2P || cfg->icp_method > ELM_AVGICP) return fail(ELM_ERR_INVALID, "unknown icp_method");
    if (cfg->use_radar_cov) return fail(ELM_ERR_UNSUPPORTED, "use_radar_cov = 1 is out of scope (reference quirk Q14)");
    if (map->device < 0) return fail(ELM_ERR_CUDA, "map is host-only (device = -1): no CUDA device to run on");
    if (!map->host.vkey.empty()) {
        if (cfg->icp_method == ELM_GICP && !map->d_prec) return fail(ELM_ERR_STATE, "GICP needs elm_map_cal_point_cov first");
        if ((cfg->icp_method == ELM_VGICP || cfg->icp_method == ELM_AVGICP) && !map->d_vrec)
            return fail(ELM_ERR_STATE, "VGICP/AVGICP need elm_map_cal_voxel_cov first");
    }
    return ELM_OK;
}

elm::IcpParams make_params(const elm_registration* r, const elm_reg_config* cfg, size_t n) {
    elm::IcpParams p;
    p.method = cfg->icp_method;
    p.n = static_cast<int>(n);
    p.queries_per_warp = 1;
    p.max_dist2 = cfg->max_search_dist * cfg->max_search_dist;
    p.th = cfg->max_search_dist;
    p.lm_lambda = cfg->lm_lambda;
    p.term_thr = cfg->icp_termination_threshold_m;
    p.min_overlap = cfg->min_overlap_ratio;
    p.warm_margin = 0.0;  // (set per map in enqueue_linearize)
    p.stats = r->stats_on ? r->d_stats : nullptr;
    p.peer = r->peer;
    return p;
}

// one linearisation: search kernel -> accumulate kernel (its last block reduces in a fixed order and, on a single
// GPU, also solves) -> multi-GPU: allreduce of the 30 sums over ranks, then the solve kernel
// `warm`: the work buffers hold the previous iteration's matches of the SAME scan (P2P / GICP: warm-started search)
int enqueue_linearize(elm_registration* r, const elm_map* map, const float* d_scan, const elm::IcpParams& prm_in, bool solve, bool warm = false) {
    const elm::IcpParams& prm0 = prm_in;
    const int sgrid = elm::icp_search_grid(prm0, r->num_sms);
    const int agrid = elm::icp_accumulate_grid(prm0, r->num_sms);
    // P2P / GICP: search, linearisation, reduction and solve are ONE kernel (unless the search runs on the binned copy,
    // whose order differs from the caller's: then the accumulation stays a separate launch in the caller's order)
    const bool fuse = r->fuse && prm0.method <= ELM_GICP && !(r->use_sorted && !r->sorted_all);
    const int wgrid = elm::icp_warm_grid(prm0, r->num_sms), rgrid = elm::icp_warm_refresh_grid(prm0, r->num_sms), xgrid = elm::icp_warm_refresh_async_grid(r->num_sms, r->async_grid);
    int rc = ensure_partials(r, std::max(std::max(sgrid, agrid), wgrid + std::max(rgrid, xgrid)));
    if (rc) return rc;
    rc = ensure_match(r, prm0.n);
    if (rc) return rc;
    if (r->profiling) {
        while (static_cast<int>(r->ev.size()) < r->ev_used + 3) {
            cudaEvent_t e;
            ELM_CUDA(cudaEventCreate(&e));
            r->ev.push_back(e);
        }
        ELM_CUDA(cudaEventRecord(r->ev[r->ev_used], r->stream));
    }
    // peer mode: the exchange happens inside the kernel, then the solve; NCCL mode: allreduce + separate solve launch below
    const bool use_nccl = r->comm != nullptr && r->peer.world == 0;
    const int solve_here = (solve && !use_nccl) ? 1 : 0;
    elm::IcpParams prm = prm_in;
    prm.warm_margin = r->warm_margin_vox * map->host.voxel_size;
    elm::IcpWork wk = r->work();
    if (fuse && !r->keep_match && prm.method == ELM_P2P) wk.match = nullptr;  // (GICP's accumulation reads the covariance record by index)
    if (r->use_sorted && r->sorted_all) d_scan = r->d_sorted;  // the whole iteration runs on the binned copy
    const bool mapped = r->use_sorted && !r->sorted_all;       // search in binned order, outputs under the caller's index
    // P2P / GICP from the second iteration of a call on: the warm pair of kernels (reuse: search + linearisation of the
    // queries whose candidate lists still hold; refresh: the rest, then reduction and solve)
    const bool use_warm = warm && r->warm && r->prune && !mapped && !fuse && prm.method <= ELM_GICP;
    const int warm_mode = r->warm_mode >= 0 ? r->warm_mode : 2;
    if (use_warm && warm_mode == 1 && r->warm_iterations_enqueued > 0) {
        // every warm iteration after the first: ONE kernel (stragglers refreshed in place by their own warp)
        if (r->profiling) ELM_CUDA(cudaEventRecord(r->ev[r->ev_used + 1], r->stream));
        ELM_CUDA(elm::launch_icp_warm(map->view(), d_scan, prm, r->d_state, wk, wgrid, solve_here, r->stream));
        r->launches += 1;
    } else if (use_warm && warm_mode == 2 && r->warm_iterations_enqueued > 0 && r->async_iterations < elm::kMaxAsyncIterations) {
        // ... or the reuse kernel with the refresh kernel running beside it: the reuse blocks publish every tile's work list under
        // this iteration's epoch, the refresh blocks take the tiles as they arrive
        if (++r->warm_epoch == 0) {  // (wrapped: forget every flag)
            ELM_CUDA(cudaMemsetAsync(r->d_tile_flag, 0, (r->match_cap / 256 + 1) * sizeof(unsigned long long), r->stream));
            r->warm_epoch = 1;
        }
        wk.epoch = r->warm_epoch;
        wk.tile_ticket = r->d_tile_ticket + 2 * static_cast<size_t>(r->async_iterations++);
        ELM_CUDA(elm::launch_icp_warm_reuse(map->view(), d_scan, prm, r->d_state, wk, wgrid, r->stream));
        if (r->profiling) ELM_CUDA(cudaEventRecord(r->ev[r->ev_used + 1], r->stream));
        ELM_CUDA(elm::launch_icp_warm_refresh_async(map->view(), d_scan, prm, r->d_state, wk, wgrid, xgrid, solve_here, r->stream));
        r->launches += 2;
    } else if (use_warm && r->warm_iterations_enqueued == 0) {
        // the first warm iteration of a call: no candidate lists exist yet, EVERY query is refreshed — the refresh kernel alone, in
        // its all-queries mode (a reuse kernel in front of it would only fill the work list with every index: 13 us)
        if (r->profiling) ELM_CUDA(cudaEventRecord(r->ev[r->ev_used + 1], r->stream));
        ELM_CUDA(elm::launch_icp_warm_refresh(map->view(), d_scan, prm, r->d_state, wk, -1, rgrid, solve_here, r->stream));
        r->launches += 1;
    } else if (use_warm) {
        ELM_CUDA(elm::launch_icp_warm_reuse(map->view(), d_scan, prm, r->d_state, wk, wgrid, r->stream));
        if (r->profiling) ELM_CUDA(cudaEventRecord(r->ev[r->ev_used + 1], r->stream));
        ELM_CUDA(elm::launch_icp_warm_refresh(map->view(), d_scan, prm, r->d_state, wk, wgrid, rgrid, solve_here, r->stream));
        r->launches += 2;
    } else {
        if (prm.method != ELM_AVGICP) {
            ELM_CUDA(elm::launch_icp_search(map->view(), mapped ? r->d_sorted : d_scan, mapped ? r->d_orig : nullptr, prm, r->d_state,
                                            wk, sgrid, r->prune, fuse ? 1 : 0, solve_here, r->stream));
            r->launches += 1;
        }
        if (r->profiling) ELM_CUDA(cudaEventRecord(r->ev[r->ev_used + 1], r->stream));
        if (!fuse) {
            ELM_CUDA(elm::launch_icp_accumulate(map->view(), d_scan, prm, r->d_state, wk, solve_here, agrid, r->stream));
            r->launches += 1;
        }
    }
    if (r->profiling) {
        ELM_CUDA(cudaEventRecord(r->ev[r->ev_used + 2], r->stream));
        r->ev_kind.resize(r->ev_used / 3 + 1);
        r->ev_kind[r->ev_used / 3] = use_warm ? (r->warm_iterations_enqueued == 0 ? 1 : 2) : 0;
        r->ev_used += 3;
    }
    r->warm_iterations_enqueued = use_warm ? r->warm_iterations_enqueued + 1 : 0;
    if (use_nccl) {
        const int e = g_nccl.AllReduce(r->d_state->acc, r->d_state->acc, elm::kAcc, kNcclFloat64, kNcclSum, r->comm, r->stream);
        if (e != 0) return fail(ELM_ERR_NCCL, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(e));
        r->launches += 1;
        if (solve) {
            ELM_CUDA(elm::launch_icp_solve(r->d_state, prm, r->stream));
            r->launches += 1;
        }
    }
    return ELM_OK;
}

void collect_profile(elm_registration* reg) {
    for (int i = 0; i + 2 < reg->ev_used; i += 3) {
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, reg->ev[i], reg->ev[i + 1]) == cudaSuccess &&
            cudaEventElapsedTime(&b, reg->ev[i + 1], reg->ev[i + 2]) == cudaSuccess) {
            reg->prof_search_ms += a; reg->prof_accum_ms += b; reg->prof_launches += 1;
            const int kind = reg->ev_kind[i / 3];
            reg->prof_kind_ms[kind][0] += a; reg->prof_kind_ms[kind][1] += b; reg->prof_kind_n[kind] += 1;
        }
    }
    reg->ev_used = 0;
}

}  // namespace

// ======================================================================================================================
extern "C" {

const char* elm_last_error(void) { return g_err.c_str(); }

int elm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int elm_map_create(elm_map** out, double voxel_size, int max_points_per_voxel, int device) try {
    if (!out || !(voxel_size > 0.0) || max_points_per_voxel < 1) return fail(ELM_ERR_INVALID, "elm_map_create: bad argument");
    if (device >= 0) {
        if (device >= elm_device_count()) return fail(ELM_ERR_CUDA, "elm_map_create: no such CUDA device");
        ELM_CUDA(cudaSetDevice(device));
    }
    elm_map* m = new (std::nothrow) elm_map();
    if (!m) return fail(ELM_ERR_INVALID, "out of memory");
    m->host.voxel_size = voxel_size;
    m->host.cap = max_points_per_voxel;
    m->device = device;
    *out = m;
    return ELM_OK;
} ELM_API_CATCH

void elm_map_destroy(elm_map* map) { delete map; }

int elm_map_add_points(elm_map* map, const float* xyz, size_t n) try {
    if (!map || (!xyz && n)) return fail(ELM_ERR_INVALID, "elm_map_add_points: bad argument");
    const auto t0 = std::chrono::steady_clock::now();
    if (map->device >= 0 && map->gpu_build && map->host.vkey.empty() && n > 0) {
        ELM_CUDA(cudaSetDevice(map->device));
        elm::GpuCanonicalMap g;
        std::string e = elm::gpu_add_points(xyz, n, map->host.voxel_size, map->host.cap, g);
        if (!e.empty()) { cudaGetLastError(); return fail(e.find("outside") != std::string::npos ? ELM_ERR_RANGE : ELM_ERR_CUDA, "elm_map_add_points (GPU builder): " + e); }
        map->build_ms[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        e = map->host.adopt_canonical(g.vkey, g.vstart, g.pxyz, g.porig, n);
        if (!e.empty()) return fail(ELM_ERR_RANGE, e);
        return map->publish_points();
    }
    const std::string e = map->host.add_points(xyz, n);
    if (!e.empty()) return fail(ELM_ERR_RANGE, e);
    map->build_ms[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return map->publish_points();
} ELM_API_CATCH

int elm_map_cal_voxel_cov(elm_map* map) try {
    if (!map) return fail(ELM_ERR_INVALID, "null map");
    if (map->host.V() >= (1ull << elm::kVcandVoxelBits)) return fail(ELM_ERR_RANGE, "elm_map_cal_voxel_cov: more than 2^25 voxels");
    const auto t0 = std::chrono::steady_clock::now();
    if (map->device >= 0 && map->gpu_build && map->host.V() > 0) {
        ELM_CUDA(cudaSetDevice(map->device));
        std::vector<double> mean, cov;
        const std::string e = elm::gpu_cal_voxel_cov(map->host.pxyz, map->host.vstart, mean, cov);
        if (!e.empty()) { cudaGetLastError(); return fail(ELM_ERR_CUDA, "elm_map_cal_voxel_cov (GPU builder): " + e); }
        map->build_ms[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        map->host.adopt_voxel_cov(mean, cov);
        return map->publish_voxel_cov();
    }
    map->host.cal_voxel_cov();
    map->build_ms[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return map->publish_voxel_cov();
} ELM_API_CATCH

int elm_map_cal_point_cov(elm_map* map, double search_dist) try {
    if (!map) return fail(ELM_ERR_INVALID, "null map");
    const auto t0 = std::chrono::steady_clock::now();
    if (map->device >= 0 && map->gpu_build && map->host.P() > 0 && map->d_dslots) {
        ELM_CUDA(cudaSetDevice(map->device));
        std::vector<double> mean, cov, nrm;
        const std::string e = elm::gpu_cal_point_cov(map->host.pxyz, map->d_dslots, map->d_drows, map->host.dir_bmask, map->host.voxel_size, search_dist, mean, cov, nrm);
        if (!e.empty()) { cudaGetLastError(); return fail(ELM_ERR_CUDA, "elm_map_cal_point_cov (GPU builder): " + e); }
        map->build_ms[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        map->host.adopt_point_cov(mean, cov, nrm);
        return map->publish_point_cov();
    }
    map->host.cal_point_cov(search_dist);
    map->build_ms[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return map->publish_point_cov();
} ELM_API_CATCH

int elm_map_set_gpu_build(elm_map* map, int enable) try {
    if (!map) return fail(ELM_ERR_INVALID, "null map");
    map->gpu_build = enable ? 1 : 0;
    return ELM_OK;
} ELM_API_CATCH

int elm_map_build_times(const elm_map* map, double ms[3]) try {
    if (!map || !ms) return fail(ELM_ERR_INVALID, "bad argument");
    for (int i = 0; i < 3; ++i) ms[i] = map->build_ms[i];
    return ELM_OK;
} ELM_API_CATCH

int elm_map_empty(const elm_map* map) { return (!map || map->host.vkey.empty()) ? 1 : 0; }
size_t elm_map_num_voxels(const elm_map* map) { return map ? map->host.V() : 0; }
size_t elm_map_num_points(const elm_map* map) { return map ? map->host.P() : 0; }

int elm_map_export(const elm_map* map, int32_t* keys, int32_t* counts, double* vmean, double* vcov, float* pxyz,
                   double* pmean, double* pcov) try {
    if (!map) return fail(ELM_ERR_INVALID, "null map");
    const elm::HostMap& h = map->host;
    if ((vmean || vcov) && !h.has_vcov) return fail(ELM_ERR_STATE, "voxel covariances not computed");
    if ((pmean || pcov) && !h.has_pcov) return fail(ELM_ERR_STATE, "point covariances not computed");
    for (size_t v = 0; v < h.V(); ++v) {
        if (keys) elm::unpack_key(h.vkey[v], keys[3 * v], keys[3 * v + 1], keys[3 * v + 2]);
        if (counts) counts[v] = static_cast<int32_t>(h.vstart[v + 1] - h.vstart[v]);
    }
    if (vmean) std::memcpy(vmean, h.vmean.data(), h.vmean.size() * sizeof(double));
    if (vcov) std::memcpy(vcov, h.vcov.data(), h.vcov.size() * sizeof(double));
    if (pxyz) std::memcpy(pxyz, h.pxyz.data(), h.pxyz.size() * sizeof(float));
    if (pmean) std::memcpy(pmean, h.pmean.data(), h.pmean.size() * sizeof(double));
    if (pcov) std::memcpy(pcov, h.pcov.data(), h.pcov.size() * sizeof(double));
    return ELM_OK;
} ELM_API_CATCH

int elm_shape_pcm_covariance(const double R_ego[9], const double local_cov[36], double icp_pose_std_m, double pose_cov[36]) try {
    if (!R_ego || !local_cov || !pose_cov) return fail(ELM_ERR_INVALID, "elm_shape_pcm_covariance: bad argument");
    elm::shape_pcm_covariance_hd(R_ego, local_cov, icp_pose_std_m, pose_cov);
    return ELM_OK;
} ELM_API_CATCH

int elm_map_find_ground_height(const elm_map* map, double x, double y, double* ground_z, int32_t* found) try {
    if (!map || !ground_z || !found) return fail(ELM_ERR_INVALID, "elm_map_find_ground_height: bad argument");
    double z = 0.0;
    *found = map->host.find_ground_height(x, y, z) ? 1 : 0;
    if (*found) *ground_z = z;  // untouched otherwise, like the reference's out-parameter
    return ELM_OK;
} ELM_API_CATCH

int elm_pcd_read_xyz(const char* path, float* xyz, size_t capacity_points, size_t* n_points, size_t* n_dropped) try {
    if (!path || !n_points) return fail(ELM_ERR_INVALID, "elm_pcd_read_xyz: bad argument");
    std::vector<float> v;
    size_t dropped = 0;
    const std::string e = elm::read_pcd_xyz(path, v, &dropped);
    if (!e.empty()) return fail(ELM_ERR_IO, e);
    *n_points = v.size() / 3;
    if (n_dropped) *n_dropped = dropped;
    if (xyz) {
        if (capacity_points < v.size() / 3) return fail(ELM_ERR_INVALID, "elm_pcd_read_xyz: buffer too small");
        std::memcpy(xyz, v.data(), v.size() * sizeof(float));
    }
    return ELM_OK;
} ELM_API_CATCH

int elm_map_add_points_pcd(elm_map* map, const char* path, size_t* n_points) try {
    if (!map || !path) return fail(ELM_ERR_INVALID, "elm_map_add_points_pcd: bad argument");
    std::vector<float> v;
    const std::string e = elm::read_pcd_xyz(path, v, nullptr);
    if (!e.empty()) return fail(ELM_ERR_IO, e);
    if (n_points) *n_points = v.size() / 3;
    return elm_map_add_points(map, v.data(), v.size() / 3);
} ELM_API_CATCH

int elm_map_save(const elm_map* map, const char* path) try {
    if (!map || !path) return fail(ELM_ERR_INVALID, "elm_map_save: bad argument");
    const std::string e = map->host.save(path);
    if (!e.empty()) return fail(ELM_ERR_IO, e);
    return ELM_OK;
} ELM_API_CATCH

int elm_map_load(elm_map** out, const char* path, int device) try {
    if (!out || !path) return fail(ELM_ERR_INVALID, "elm_map_load: bad argument");
    if (device >= 0) {
        if (device >= elm_device_count()) return fail(ELM_ERR_CUDA, "elm_map_load: no such CUDA device");
        ELM_CUDA(cudaSetDevice(device));
    }
    std::unique_ptr<elm_map> m(new (std::nothrow) elm_map());
    if (!m) return fail(ELM_ERR_INVALID, "out of memory");
    const std::string e = m->host.load(path);
    if (!e.empty()) return fail(ELM_ERR_IO, e);
    m->device = device;
    int rc = m->publish_points();
    if (rc == ELM_OK && m->host.has_vcov) rc = m->publish_voxel_cov();
    if (rc == ELM_OK && m->host.has_pcov) rc = m->publish_point_cov();
    if (rc != ELM_OK) return rc;
    *out = m.release();
    return ELM_OK;
} ELM_API_CATCH

int elm_map_directory_check(const elm_map* map, uint64_t* entries, uint64_t* slots, uint64_t* mismatches) try {
    if (!map || !entries || !slots || !mismatches) return fail(ELM_ERR_INVALID, "elm_map_directory_check: bad argument");
    const elm::HostMap& h = map->host;
    *entries = h.dir_entries;
    *slots = h.dir_slots.size();
    uint64_t bad = 0, found = 0;
    // every stored slot: descriptors against the canonical arrays
    for (size_t s = 0; s < h.dir_slots.size(); ++s) {
        const elm::DirSlot& sl = h.dir_slots[s];
        if ((sl.key_lo & sl.key_hi) == 0xffffffffu) continue;
        ++found;
        const uint64_t key = (static_cast<uint64_t>(sl.key_hi) << 32) | sl.key_lo;
        if (h.dir_find(key) != static_cast<int64_t>(s)) ++bad;
        int32_t x, y, z;
        elm::unpack_key(key, x, y, z);
        bool any = false;
        for (int c = 0; c < 9; ++c) {
            const elm::DirDesc d = h.row_column(s, c);
            uint32_t first = 0, counts = 0;
            bool have = false;
            for (int dz = -1; dz <= 1; ++dz) {
                const int32_t vx = x + c / 3 - 1, vy = y + c % 3 - 1, vz = z + dz;
                if (!elm::key_in_range(vx) || !elm::key_in_range(vy) || !elm::key_in_range(vz)) continue;
                const int64_t v = h.find(elm::pack_key(vx, vy, vz));
                if (v < 0) continue;
                // octant word of the voxel: cumulative counts of its points by octant, recomputed from the device order
                if (h.octants) {
                    uint8_t want[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    int prev = 0;
                    for (uint32_t p = h.vstart[v]; p < h.vstart[v + 1]; ++p) {
                        const float* q = &h.pxyz[3 * static_cast<size_t>(h.dev_order[p])];
                        const int o = (elm::axis_half(static_cast<double>(q[2]) / h.voxel_size, vz) << 2) |
                                      (elm::axis_half(static_cast<double>(q[1]) / h.voxel_size, vy) << 1) |
                                      elm::axis_half(static_cast<double>(q[0]) / h.voxel_size, vx);
                        if (o < prev) ++bad;  // device order sorted by octant
                        if (p > h.vstart[v] && o == prev && h.dev_order[p] < h.dev_order[p - 1]) ++bad;  // stable inside an octant
                        prev = o;
                        for (int k = o + 1; k <= 7; ++k) ++want[k - 1];
                    }
                    want[7] = static_cast<uint8_t>(h.vstart[v + 1] - h.vstart[v]);
                    if (std::memcmp(want, &h.dir_rows[elm::row_col_word(s, c) + 2 + 2 * static_cast<size_t>(dz + 1)], 8) != 0) ++bad;
                }
                if (!have) { first = h.vstart[v]; have = true; }
                else if (h.vstart[v] != first + (counts & elm::kDirCountMask) + ((counts >> elm::kDirCountBits) & elm::kDirCountMask)) ++bad;  // contiguity
                counts |= (h.vstart[v + 1] - h.vstart[v]) << (elm::kDirCountBits * static_cast<uint32_t>(dz + 1));
            }
            any = any || have;
            if (d.counts != counts || (have && d.first != first)) ++bad;
            if (c == 4 && (sl.counts != counts || (have && sl.first != first))) ++bad;
        }
        if (!any) ++bad;  // an entry whose neighbourhood is empty should not exist
        if (h.has_vcov) {  // VGICP / AVGICP candidate list of the entry: occupancy mask, order, fp32 means, voxel slots
            const elm::DirDesc run{h.dir_rows[elm::row_word(s, elm::kRowCandFirst)], h.dir_rows[elm::row_word(s, elm::kRowCandCount)]};
            const elm::DirDesc occ{h.dir_rows[elm::row_word(s, elm::kRowOccMask)], 0};
            if (h.dir_rows[elm::row_word(s, elm::kRowKeyLo)] != sl.key_lo || h.dir_rows[elm::row_word(s, elm::kRowKeyHi)] != sl.key_hi) ++bad;
            uint32_t c = 0;
            for (int L = 0; L < 27; ++L) {
                const int32_t vx = x + L / 9 - 1, vy = y + (L / 3) % 3 - 1, vz = z + L % 3 - 1;
                const int64_t v = (elm::key_in_range(vx) && elm::key_in_range(vy) && elm::key_in_range(vz)) ? h.find(elm::pack_key(vx, vy, vz)) : -1;
                if ((v >= 0) != (((occ.first >> L) & 1u) != 0)) { ++bad; continue; }
                if (v < 0) continue;
                float ox, oy, oz;
                uint32_t vox;
                elm::unpack_vcand(h.vcand8[run.first + c], ox, oy, oz, vox);
                if (vox != static_cast<uint32_t>(v)) ++bad;
                const double want[3] = {h.vmean[3 * v] / h.voxel_size - x, h.vmean[3 * v + 1] / h.voxel_size - y, h.vmean[3 * v + 2] / h.voxel_size - z};
                const float got[3] = {ox, oy, oz};
                for (int k = 0; k < 3; ++k) if (!(std::fabs(static_cast<double>(got[k]) - want[k]) <= 2.45e-4)) ++bad;  // the band of the search relies on this
                ++c;
            }
            if (c != run.counts) ++bad;
            static const int kL7[7] = {13, 22, 4, 16, 10, 14, 12};
            for (int j = 0; j < 7; ++j) {
                const int L = kL7[j];
                const int32_t vx = x + L / 9 - 1, vy = y + (L / 3) % 3 - 1, vz = z + L % 3 - 1;
                const int64_t v = (elm::key_in_range(vx) && elm::key_in_range(vy) && elm::key_in_range(vz)) ? h.find(elm::pack_key(vx, vy, vz)) : -1;
                const int32_t sl7 = h.dir7[8 * s + j];
                if ((v < 0) != (sl7 < 0) || (v >= 0 && sl7 != v)) ++bad;
            }
        }
    }
    if (found != h.dir_entries) ++bad;
    // every occupied voxel makes its 27 surrounding centres findable; two voxels further out along x there is a miss
    // unless that centre has its own occupied neighbour
    for (size_t v = 0; v < h.V(); ++v) {
        int32_t x, y, z;
        elm::unpack_key(h.vkey[v], x, y, z);
        for (int dx = -1; dx <= 1; ++dx) for (int dy = -1; dy <= 1; ++dy) for (int dz = -1; dz <= 1; ++dz)
            if (h.dir_find(elm::pack_key(x + dx, y + dy, z + dz)) < 0) ++bad;
        for (int dx : {-2, 2}) {
            const int32_t cx = x + dx;
            if (!elm::key_in_range(cx)) continue;
            bool occupied_near = false;
            for (int ex = -1; ex <= 1 && !occupied_near; ++ex) for (int ey = -1; ey <= 1 && !occupied_near; ++ey) for (int ez = -1; ez <= 1; ++ez)
                if (elm::key_in_range(cx + ex) && h.find(elm::pack_key(cx + ex, y + ey, z + ez)) >= 0) { occupied_near = true; break; }
            if ((h.dir_find(elm::pack_key(cx, y, z)) >= 0) != occupied_near) ++bad;
        }
    }
    *mismatches = bad;
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_create(elm_registration** out, int device, void* stream) try {
    if (!out) return fail(ELM_ERR_INVALID, "null out");
    if (device < 0 || device >= elm_device_count()) return fail(ELM_ERR_CUDA, "elm_registration_create: no such CUDA device");
    ELM_CUDA(cudaSetDevice(device));
    elm_registration* r = new (std::nothrow) elm_registration();
    if (!r) return fail(ELM_ERR_INVALID, "out of memory");
    r->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) r->num_sms = prop.multiProcessorCount;
    if (stream) { r->stream = static_cast<cudaStream_t>(stream); }
    else {
        if (cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess) { delete r; return fail(ELM_ERR_CUDA, "cudaStreamCreate failed"); }
        r->own_stream = true;
    }
    if (cudaMalloc(reinterpret_cast<void**>(&r->d_ticket), sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(r->d_ticket, 0, sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&r->d_tile_ticket), 2 * elm::kMaxAsyncIterations * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(r->d_tile_ticket, 0, 2 * elm::kMaxAsyncIterations * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&r->d_state), sizeof(elm::IcpState)) != cudaSuccess ||
        cudaMemset(r->d_state, 0, sizeof(elm::IcpState)) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&r->h_state), sizeof(elm::IcpState)) != cudaSuccess) {
        delete r;
        return fail(ELM_ERR_CUDA, "state allocation failed");
    }
    std::memset(r->h_state, 0, sizeof(elm::IcpState));
    *out = r;
    return ELM_OK;
} ELM_API_CATCH

void elm_registration_destroy(elm_registration* reg) { delete reg; }

int elm_register_enqueue(elm_registration* reg, const elm_map* map, const float* d_src_xyz, size_t n, const double T_init[16],
                         const elm_reg_config* cfg) try {
    if (!reg || !map || !T_init || !cfg || (!d_src_xyz && n)) return fail(ELM_ERR_INVALID, "elm_register_enqueue: bad argument");
    if (n > 0x7fffffffull / 8) return fail(ELM_ERR_INVALID, "scan too large");
    int rc = check_method(map, cfg);
    if (rc) return rc;
    if (map->device != reg->device) return fail(ELM_ERR_INVALID, "map and registration live on different devices");
    ELM_CUDA(cudaSetDevice(reg->device));
    reg->cfg = *cfg;
    std::memcpy(reg->T_init, T_init, sizeof reg->T_init);
    reg->launches = 0;
    reg->pending = true;
    // reg.cpp:291-295 (empty map) and the n == 0 guard: nothing to run.  In the sharded mode a rank with an empty shard
    // still has to take part in the allreduces, so only the single-rank case short-circuits on n == 0.
    reg->trivial = map->host.vkey.empty() || (n == 0 && !reg->sharded());
    if (reg->trivial) return ELM_OK;
    const elm::IcpParams prm = make_params(reg, cfg, n);
    ELM_CUDA(elm::launch_icp_begin(reg->d_state, T_init, reg->d_ticket, reg->d_tile_ticket, reg->stream));
    reg->async_iterations = 0;
    reg->launches += 1;
    rc = enqueue_binning(reg, map, d_src_xyz, n, T_init, cfg->icp_method);
    if (rc) return rc;
    reg->sorted_all = true;
    for (int j = 0; j < cfg->max_iteration; ++j) {  // reg.cpp:310
        rc = enqueue_linearize(reg, map, d_src_xyz, prm, true, j > 0);
        if (rc) return rc;
    }
    ELM_CUDA(cudaMemcpyAsync(reg->h_state, reg->d_state, sizeof(elm::IcpState), cudaMemcpyDeviceToHost, reg->stream));
    return ELM_OK;
} ELM_API_CATCH

int elm_register_fetch(elm_registration* reg, double T_out[16], int32_t* is_success, double* fitness_score, double local_cov[36],
                       int32_t* iterations_run) try {
    if (!reg || !T_out || !is_success) return fail(ELM_ERR_INVALID, "elm_register_fetch: bad argument");
    if (!reg->pending) return fail(ELM_ERR_STATE, "elm_register_fetch without elm_register_enqueue");
    ELM_CUDA(cudaSetDevice(reg->device));
    reg->pending = false;
    if (reg->trivial) {
        std::memcpy(T_out, reg->T_init, 16 * sizeof(double));
        *is_success = 0;
        if (local_cov) for (int i = 0; i < 36; ++i) local_cov[i] = (i % 7 == 0) ? 1.0 : 0.0;
        if (iterations_run) *iterations_run = 0;
        return ELM_OK;
    }
    ELM_CUDA(cudaStreamSynchronize(reg->stream));
    if (reg->profiling) collect_profile(reg);
    const elm::IcpState& st = *reg->h_state;
    if (st.comm_error) return fail(ELM_ERR_NCCL, "peer-memory exchange timed out: a rank never delivered its accumulators");
    std::memcpy(T_out, st.T, 16 * sizeof(double));
    if (local_cov) std::memcpy(local_cov, st.local_cov, 36 * sizeof(double));
    if (iterations_run) *iterations_run = st.iterations;
    if (st.overlap_fail) {                                   // reg.cpp:352-356
        *is_success = 0;
    } else if (st.fitness > reg->cfg.max_fitness_score) {    // reg.cpp:405-409
        *is_success = 0;
    } else {                                                 // reg.cpp:415-417
        if (fitness_score) *fitness_score = st.fitness;
        *is_success = 1;
    }
    if (reg->cfg.debug_print)
        std::printf("[elimaloc_b200] RunRegister: %d iterations, fitness %.6f, success %d\n", st.iterations, st.fitness, *is_success);
    return ELM_OK;
} ELM_API_CATCH

int elm_run_register(elm_registration* reg, const elm_map* map, const float* src_xyz, size_t n, const double T_init[16],
                     const elm_reg_config* cfg, double T_out[16], int32_t* is_success, double* fitness_score, double local_cov[36]) try {
    if (!reg || !map || !T_init || !cfg || !T_out || !is_success || (!src_xyz && n)) return fail(ELM_ERR_INVALID, "elm_run_register: bad argument");
    int rc = check_method(map, cfg);
    if (rc) return rc;
    ELM_CUDA(cudaSetDevice(reg->device));
    if (n && !map->host.vkey.empty()) {
        rc = ensure_scan(reg, n);
        if (rc) return rc;
        ELM_CUDA(cudaMemcpyAsync(reg->d_scan, src_xyz, n * 3 * sizeof(float), cudaMemcpyHostToDevice, reg->stream));
    }
    rc = elm_register_enqueue(reg, map, reg->d_scan, n, T_init, cfg);
    if (rc) return rc;
    return elm_register_fetch(reg, T_out, is_success, fitness_score, local_cov, nullptr);
} ELM_API_CATCH

int elm_linearize(elm_registration* reg, const elm_map* map, const float* src_xyz, size_t n, const double T[16],
                  const elm_reg_config* cfg, double JTJ[36], double JTr[6], double* residual_sum, int64_t* n_corr) try {
    if (!reg || !map || !T || !cfg || !JTJ || !JTr || !residual_sum || !n_corr || (!src_xyz && n)) return fail(ELM_ERR_INVALID, "elm_linearize: bad argument");
    int rc = check_method(map, cfg);
    if (rc) return rc;
    if (map->device != reg->device) return fail(ELM_ERR_INVALID, "map and registration live on different devices");
    ELM_CUDA(cudaSetDevice(reg->device));
    std::memset(JTJ, 0, 36 * sizeof(double));
    std::memset(JTr, 0, 6 * sizeof(double));
    *residual_sum = 0.0;
    *n_corr = 0;
    if (map->host.vkey.empty() || (n == 0 && !reg->sharded())) return ELM_OK;
    rc = ensure_scan(reg, n);
    if (rc) return rc;
    if (n) ELM_CUDA(cudaMemcpyAsync(reg->d_scan, src_xyz, n * 3 * sizeof(float), cudaMemcpyHostToDevice, reg->stream));
    const elm::IcpParams prm = make_params(reg, cfg, n);
    ELM_CUDA(elm::launch_icp_begin(reg->d_state, T, reg->d_ticket, reg->d_tile_ticket, reg->stream));
    reg->async_iterations = 0;
    rc = enqueue_binning(reg, map, reg->d_scan, n, T, cfg->icp_method);
    if (rc) return rc;
    rc = enqueue_linearize(reg, map, reg->d_scan, prm, false);
    if (rc) return rc;
    ELM_CUDA(cudaMemcpyAsync(reg->h_state, reg->d_state, sizeof(elm::IcpState), cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaStreamSynchronize(reg->stream));
    const double* a = reg->h_state->acc;
    int k = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { JTJ[6 * i + j] = a[k]; JTJ[6 * j + i] = a[k]; ++k; }
    for (int i = 0; i < 6; ++i) JTr[i] = a[elm::kIdxJtr + i];
    *residual_sum = a[elm::kIdxRes];
    *n_corr = static_cast<int64_t>(std::llround(a[elm::kIdxNcorr]));
    return ELM_OK;
} ELM_API_CATCH

int elm_correspondences_sequence(elm_registration* reg, const elm_map* map, const float* src_xyz, size_t n, const double* T_seq, int n_poses,
                                 int method, double max_search_dist, int32_t* count, double* target) try {
    if (!reg || !map || !T_seq || n_poses < 1 || !count || !target || (!src_xyz && n)) return fail(ELM_ERR_INVALID, "elm_correspondences: bad argument");
    elm_reg_config c{};
    c.icp_method = method;
    int rc = check_method(map, &c);
    if (rc) return rc;
    const int K = (method == ELM_AVGICP) ? 7 : 1;
    std::memset(count, 0, n * sizeof(int32_t));
    std::memset(target, 0, n * K * 3 * sizeof(double));
    if (map->host.vkey.empty() || n == 0) return ELM_OK;
    ELM_CUDA(cudaSetDevice(reg->device));
    rc = ensure_scan(reg, n);
    if (rc) return rc;
    if (n * 7 > reg->hook_cap) {
        cudaFree(reg->d_count); cudaFree(reg->d_target);
        reg->d_count = nullptr; reg->d_target = nullptr;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_count), n * sizeof(int)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_target), n * 21 * sizeof(double)));
        reg->hook_cap = n * 7;
    }
    ELM_CUDA(cudaMemcpyAsync(reg->d_scan, src_xyz, n * 3 * sizeof(float), cudaMemcpyHostToDevice, reg->stream));
    // the hook runs the PRODUCTION search kernels — cold at the first pose, warm-started at the following ones exactly as
    // the ICP loop runs them — and only converts the last match[] into positions
    c.max_search_dist = max_search_dist;
    elm::IcpParams prm = make_params(reg, &c, n);
    prm.warm_margin = reg->warm_margin_vox * map->host.voxel_size;
    rc = ensure_match(reg, n);
    if (rc) return rc;
    for (int k = 0; k < n_poses; ++k) {
        const double* T = T_seq + 16 * static_cast<size_t>(k);
        ELM_CUDA(elm::launch_icp_begin(reg->d_state, T, reg->d_ticket, reg->d_tile_ticket, reg->stream));
    reg->async_iterations = 0;
        if (k == 0) {
            rc = enqueue_binning(reg, map, reg->d_scan, n, T, method);
            if (rc) return rc;
        }
        const bool warm = k > 0 && reg->warm && reg->prune && !reg->use_sorted && method <= ELM_GICP;
        if (warm) {
            const int wgrid = elm::icp_warm_grid(prm, reg->num_sms), rgrid = elm::icp_warm_refresh_grid(prm, reg->num_sms);
            rc = ensure_partials(reg, wgrid + rgrid);
            if (rc) return rc;
            ELM_CUDA(elm::launch_icp_warm_reuse(map->view(), reg->d_scan, prm, reg->d_state, reg->work(), wgrid, reg->stream));
            ELM_CUDA(elm::launch_icp_warm_refresh(map->view(), reg->d_scan, prm, reg->d_state, reg->work(), wgrid, rgrid, 0, reg->stream));
        } else if (method != ELM_AVGICP) {
            ELM_CUDA(elm::launch_icp_search(map->view(), reg->use_sorted ? reg->d_sorted : reg->d_scan, reg->use_sorted ? reg->d_orig : nullptr, prm,
                                            reg->d_state, reg->work(), elm::icp_search_grid(prm, reg->num_sms), reg->prune, 0, 0, reg->stream));
        }
    }
    ELM_CUDA(elm::launch_icp_export(map->view(), reg->d_scan, reg->d_match, static_cast<int>(n), reg->d_state, method,
                                    max_search_dist * max_search_dist, reg->d_count, reg->d_target, reg->stream));
    ELM_CUDA(cudaMemcpyAsync(count, reg->d_count, n * sizeof(int), cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaMemcpyAsync(target, reg->d_target, n * K * 3 * sizeof(double), cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaStreamSynchronize(reg->stream));
    return ELM_OK;
} ELM_API_CATCH

int elm_correspondences(elm_registration* reg, const elm_map* map, const float* src_xyz, size_t n, const double T[16],
                        int method, double max_search_dist, int32_t* count, double* target) try {
    return elm_correspondences_sequence(reg, map, src_xyz, n, T, 1, method, max_search_dist, count, target);
} ELM_API_CATCH

int elm_registration_set_warm_start(elm_registration* reg, int enable) try {
    if (!reg) return fail(ELM_ERR_INVALID, "null registration");
    reg->warm = enable ? 1 : 0;
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_launch_count(const elm_registration* reg, int64_t* launches) try {
    if (!reg || !launches) return fail(ELM_ERR_INVALID, "bad argument");
    *launches = reg->launches;
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_set_profiling(elm_registration* reg, int enable) try {
    if (!reg) return fail(ELM_ERR_INVALID, "null registration");
    reg->profiling = enable != 0;
    reg->ev_used = 0;
    reg->prof_search_ms = reg->prof_accum_ms = 0.0;
    reg->prof_launches = 0;
    for (int k = 0; k < 3; ++k) { reg->prof_kind_ms[k][0] = reg->prof_kind_ms[k][1] = 0.0; reg->prof_kind_n[k] = 0; }
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_profile(const elm_registration* reg, double* search_ms, double* accumulate_ms, int64_t* iterations) try {
    if (!reg || !search_ms || !accumulate_ms || !iterations) return fail(ELM_ERR_INVALID, "bad argument");
    *search_ms = reg->prof_search_ms;
    *accumulate_ms = reg->prof_accum_ms;
    *iterations = reg->prof_launches;
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_profile_by_kind(const elm_registration* reg, double ms[6], int64_t iterations[3]) try {
    if (!reg || !ms || !iterations) return fail(ELM_ERR_INVALID, "bad argument");
    for (int k = 0; k < 3; ++k) { ms[2 * k] = reg->prof_kind_ms[k][0]; ms[2 * k + 1] = reg->prof_kind_ms[k][1]; iterations[k] = reg->prof_kind_n[k]; }
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_set_stats(elm_registration* reg, int enable) try {
    if (!reg) return fail(ELM_ERR_INVALID, "null registration");
    ELM_CUDA(cudaSetDevice(reg->device));
    if (!reg->d_stats) ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_stats), 32 * sizeof(unsigned long long)));
    ELM_CUDA(cudaMemsetAsync(reg->d_stats, 0, 32 * sizeof(unsigned long long), reg->stream));
    reg->stats_on = enable != 0;
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_stats(elm_registration* reg, uint64_t* map_points_visited, uint64_t* queries) try {
    if (!reg || !map_points_visited || !queries) return fail(ELM_ERR_INVALID, "bad argument");
    if (!reg->d_stats) return fail(ELM_ERR_STATE, "stats were never enabled");
    ELM_CUDA(cudaSetDevice(reg->device));
    unsigned long long h[32] = {0};
    ELM_CUDA(cudaMemcpyAsync(h, reg->d_stats, sizeof h, cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaStreamSynchronize(reg->stream));
    if (getenv("ELM_PHASE_TIMING"))  // developer aid: per-warp cycle sums of the search phases (-DELM_PHASE_TIMING builds)
        std::fprintf(stderr, "[phase cycles] warps %llu | tile %llu lookup %llu home %llu prune+items %llu B %llu C %llu\n", h[8], h[2], h[3], h[4], h[5], h[6], h[7]);
    if (getenv("ELM_PHASE_TIMING"))
        std::fprintf(stderr, "[fine cycles] warps %llu | transform+keys %llu lookup-loads %llu home-desc %llu\n", h[8], h[16], h[17], h[18]);
    if (getenv("ELM_PHASE_TIMING"))
        std::fprintf(stderr, "[accumulate cycles] blocks %llu | state %llu gather+linearise %llu block-reduce %llu publish+ticket %llu | last block: final-reduce %llu solve %llu\n",
                     h[15], h[9], h[10], h[11], h[12], h[13], h[14]);
    if (getenv("ELM_PHASE_TIMING"))
        std::fprintf(stderr, "[warm cycles] warp-tiles %llu | inputs %llu decide %llu reuse-scan %llu refresh %llu winner-reload %llu | refresh reasons: key changed %llu, no match %llu, no list %llu, margin used up %llu\n", h[22], h[23], h[24], h[25], h[26], h[27], h[30], h[31], h[28], h[29]);
    if (getenv("ELM_WARM_STATS")) std::fprintf(stderr, "[warm search] %llu of %llu warm searches refreshed their runs\n", h[20], h[21]);
    *map_points_visited = h[0];
    *queries = h[1];
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_stats_raw(elm_registration* reg, uint64_t counters[32]) try {
    if (!reg || !counters) return fail(ELM_ERR_INVALID, "bad argument");
    if (!reg->d_stats) return fail(ELM_ERR_STATE, "stats were never enabled");
    ELM_CUDA(cudaSetDevice(reg->device));
    unsigned long long h[32] = {0};
    ELM_CUDA(cudaMemcpyAsync(h, reg->d_stats, sizeof h, cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaStreamSynchronize(reg->stream));
    for (int i = 0; i < 32; ++i) counters[i] = h[i];
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_set_fused(elm_registration* reg, int enable) try {
    if (!reg) return fail(ELM_ERR_INVALID, "null registration");
    reg->fuse = enable ? 1 : 0;
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_set_binning(elm_registration* reg, int enable) try {
    if (!reg) return fail(ELM_ERR_INVALID, "null registration");
    reg->binning = enable ? 1 : 0;
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_set_exhaustive(elm_registration* reg, int exhaustive) try {
    if (!reg) return fail(ELM_ERR_INVALID, "null registration");
    reg->prune = exhaustive ? 0 : 1;
    return ELM_OK;
} ELM_API_CATCH

int elm_deskew_points_device(elm_registration* reg, const float* d_xyz, const float* d_rel_time, size_t n, const elm_deskew_tables* t,
                             float* d_xyz_out) try {
    if (!reg || !t || (n && (!d_xyz || !d_rel_time || !d_xyz_out))) return fail(ELM_ERR_INVALID, "elm_deskew_points: bad argument");
    if (n > 0x7fffffffull / 4) return fail(ELM_ERR_INVALID, "scan too large");
    if (t->imu_available && (t->imu_pointer_cur < 0 || !t->imu_time || !t->imu_rot_x || !t->imu_rot_y || !t->imu_rot_z))
        return fail(ELM_ERR_INVALID, "elm_deskew_points: IMU table missing");
    if (n == 0) return ELM_OK;
    ELM_CUDA(cudaSetDevice(reg->device));
    const int entries = t->imu_available ? t->imu_pointer_cur + 1 : 1;
    if (entries > reg->dtable_cap) {
        cudaFree(reg->d_dtable); cudaFreeHost(reg->h_dtable);
        reg->d_dtable = nullptr; reg->h_dtable = nullptr;
        const int cap = (entries + 255) / 256 * 256;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_dtable), static_cast<size_t>(cap) * 4 * sizeof(double)));
        ELM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&reg->h_dtable), static_cast<size_t>(cap) * 4 * sizeof(double)));
        reg->dtable_cap = cap;
    }
    elm::DeskewParams p{};
    p.imu_pointer_cur = t->imu_available ? t->imu_pointer_cur : 0;
    p.imu_available = t->imu_available;
    p.odom_available = t->odom_available;
    p.table_stride = reg->dtable_cap;
    p.odom_incre_x = t->odom_incre_x; p.odom_incre_y = t->odom_incre_y; p.odom_incre_z = t->odom_incre_z;
    p.time_scan_cur = t->time_scan_cur; p.time_scan_end = t->time_scan_end;
    p.n_dev = nullptr; p.rel_time_offset = 0.f; p.run_deskew = 1;
    if (t->imu_available) {
        ELM_CUDA(cudaStreamSynchronize(reg->stream));  // the pinned staging buffer may still feed the previous call
        const double* rows[4] = {t->imu_time, t->imu_rot_x, t->imu_rot_y, t->imu_rot_z};
        for (int r = 0; r < 4; ++r) std::memcpy(reg->h_dtable + static_cast<size_t>(r) * reg->dtable_cap, rows[r], entries * sizeof(double));
        ELM_CUDA(cudaMemcpyAsync(reg->d_dtable, reg->h_dtable, static_cast<size_t>(reg->dtable_cap) * 4 * sizeof(double), cudaMemcpyHostToDevice, reg->stream));
    }
    ELM_CUDA(elm::launch_deskew_points(d_xyz, d_rel_time, static_cast<int>(n), p, reg->d_dtable, d_xyz_out, reg->num_sms, reg->stream));
    return ELM_OK;
} ELM_API_CATCH

int elm_deskew_points(elm_registration* reg, const float* xyz, const float* rel_time, size_t n, const elm_deskew_tables* t, float* xyz_out) try {
    if (!reg || !t || (n && (!xyz || !rel_time || !xyz_out))) return fail(ELM_ERR_INVALID, "elm_deskew_points: bad argument");
    if (n == 0) return ELM_OK;
    ELM_CUDA(cudaSetDevice(reg->device));
    if (n > reg->dsk_cap) {
        cudaFree(reg->d_dsk_in); cudaFree(reg->d_dsk_out);
        reg->d_dsk_in = nullptr; reg->d_dsk_out = nullptr;
        const size_t cap = (n + 1023) / 1024 * 1024;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_dsk_in), cap * 4 * sizeof(float)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_dsk_out), cap * 3 * sizeof(float)));
        reg->dsk_cap = cap;
    }
    float* d_xyz = reg->d_dsk_in;
    float* d_t = reg->d_dsk_in + 3 * reg->dsk_cap;
    ELM_CUDA(cudaMemcpyAsync(d_xyz, xyz, n * 3 * sizeof(float), cudaMemcpyHostToDevice, reg->stream));
    ELM_CUDA(cudaMemcpyAsync(d_t, rel_time, n * sizeof(float), cudaMemcpyHostToDevice, reg->stream));
    const int rc = elm_deskew_points_device(reg, d_xyz, d_t, n, t, reg->d_dsk_out);
    if (rc) return rc;
    ELM_CUDA(cudaMemcpyAsync(xyz_out, reg->d_dsk_out, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaStreamSynchronize(reg->stream));
    return ELM_OK;
} ELM_API_CATCH

int elm_scan_preprocess_device(elm_registration* reg, const float* d_xyz, const float* d_aux, size_t n, double max_dist, double voxel_size,
                               float* d_xyz_out, float* d_aux_out, int32_t* d_index_out, size_t* n_out) try {
    if (!reg || !n_out || (n && (!d_xyz || !d_xyz_out)) || (d_aux_out && !d_aux)) return fail(ELM_ERR_INVALID, "elm_scan_preprocess: bad argument");
    if (n > 0x7fffffffull / 4) return fail(ELM_ERR_INVALID, "scan too large");
    *n_out = 0;
    if (n == 0) return ELM_OK;
    ELM_CUDA(cudaSetDevice(reg->device));
    if (n > reg->prep_cap) {
        cudaFree(reg->prep.tkeys); cudaFree(reg->prep.tmin); cudaFree(reg->prep.keep); cudaFree(reg->prep.block_count); cudaFree(reg->prep.block_offset);
        reg->prep = elm::ScanPrepScratch{};
        reg->prep_cap = 0;
        const size_t cap = (n + 4095) / 4096 * 4096, slots = elm::scan_prep_table_slots(cap), blocks = (cap + 255) / 256;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.tkeys), slots * sizeof(unsigned long long)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.tmin), slots * sizeof(uint32_t)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.keep), cap));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.block_count), blocks * sizeof(uint32_t)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.block_offset), blocks * sizeof(uint32_t)));
        reg->prep.table_slots = slots;
        reg->prep_cap = cap;
    }
    if (!reg->prep.error) ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.error), sizeof(int)));
    if (!reg->d_prep_n) ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_prep_n), sizeof(int)));
    ELM_CUDA(elm::launch_scan_prep(d_xyz, d_aux, static_cast<int>(n), max_dist, voxel_size, reg->prep, d_xyz_out, d_aux_out, d_index_out,
                                   reg->d_prep_n, reg->stream));
    int h[2] = {0, 0};
    ELM_CUDA(cudaMemcpyAsync(&h[0], reg->d_prep_n, sizeof(int), cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaMemcpyAsync(&h[1], reg->prep.error, sizeof(int), cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaStreamSynchronize(reg->stream));
    if (h[1]) return fail(ELM_ERR_RANGE, "scan point not finite or beyond +-2^20 voxels of the down-sampling grid");
    *n_out = static_cast<size_t>(h[0]);
    return ELM_OK;
} ELM_API_CATCH

int elm_scan_preprocess(elm_registration* reg, const float* xyz, const float* aux, size_t n, double max_dist, double voxel_size, float* xyz_out,
                        float* aux_out, int32_t* index_out, size_t* n_out) try {
    if (!reg || !n_out || (n && (!xyz || !xyz_out)) || (aux_out && !aux)) return fail(ELM_ERR_INVALID, "elm_scan_preprocess: bad argument");
    *n_out = 0;
    if (n == 0) return ELM_OK;
    ELM_CUDA(cudaSetDevice(reg->device));
    if (n > reg->prep_io_cap) {
        cudaFree(reg->d_prep_in); cudaFree(reg->d_prep_out); cudaFree(reg->d_prep_idx);
        reg->d_prep_in = reg->d_prep_out = nullptr; reg->d_prep_idx = nullptr;
        reg->prep_io_cap = 0;
        const size_t cap = (n + 4095) / 4096 * 4096;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_prep_in), cap * 4 * sizeof(float)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_prep_out), cap * 4 * sizeof(float)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_prep_idx), cap * sizeof(int)));
        reg->prep_io_cap = cap;
    }
    float* d_aux = aux ? reg->d_prep_in + 3 * reg->prep_io_cap : nullptr;
    float* d_aux_out = aux_out ? reg->d_prep_out + 3 * reg->prep_io_cap : nullptr;
    ELM_CUDA(cudaMemcpyAsync(reg->d_prep_in, xyz, n * 3 * sizeof(float), cudaMemcpyHostToDevice, reg->stream));
    if (aux) ELM_CUDA(cudaMemcpyAsync(d_aux, aux, n * sizeof(float), cudaMemcpyHostToDevice, reg->stream));
    const int rc = elm_scan_preprocess_device(reg, reg->d_prep_in, d_aux, n, max_dist, voxel_size, reg->d_prep_out, d_aux_out,
                                              index_out ? reg->d_prep_idx : nullptr, n_out);
    if (rc) return rc;
    if (*n_out) {
        ELM_CUDA(cudaMemcpyAsync(xyz_out, reg->d_prep_out, *n_out * 3 * sizeof(float), cudaMemcpyDeviceToHost, reg->stream));
        if (aux_out) ELM_CUDA(cudaMemcpyAsync(aux_out, d_aux_out, *n_out * sizeof(float), cudaMemcpyDeviceToHost, reg->stream));
        if (index_out) ELM_CUDA(cudaMemcpyAsync(index_out, reg->d_prep_idx, *n_out * sizeof(int), cudaMemcpyDeviceToHost, reg->stream));
        ELM_CUDA(cudaStreamSynchronize(reg->stream));
    }
    return ELM_OK;
} ELM_API_CATCH

int elm_ekf_create(elm_ekf** out, const elm_ekf_config* cfg, int device, void* stream) try {
    if (!out || !cfg) return fail(ELM_ERR_INVALID, "elm_ekf_create: bad argument");
    if (device < 0 || device >= elm_device_count()) return fail(ELM_ERR_CUDA, "elm_ekf_create: no such CUDA device");
    ELM_CUDA(cudaSetDevice(device));
    elm_ekf* e = new (std::nothrow) elm_ekf();
    if (!e) return fail(ELM_ERR_INVALID, "out of memory");
    e->device = device;
    e->cfg = *cfg;
    if (stream) e->stream = static_cast<cudaStream_t>(stream);
    else {
        if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { delete e; return fail(ELM_ERR_CUDA, "cudaStreamCreate failed"); }
        e->own_stream = true;
    }
    if (cudaMalloc(reinterpret_cast<void**>(&e->d_state), sizeof(elm_ekf_state)) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&e->h_state), sizeof(elm_ekf_state)) != cudaSuccess) {
        delete e;
        return fail(ELM_ERR_CUDA, "EKF state allocation failed");
    }
    elm::ekf_init_state(*cfg, *e->h_state);
    if (cudaMemcpyAsync(e->d_state, e->h_state, sizeof(elm_ekf_state), cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
        cudaStreamSynchronize(e->stream) != cudaSuccess) {
        delete e;
        return fail(ELM_ERR_CUDA, "EKF state upload failed");
    }
    *out = e;
    return ELM_OK;
} ELM_API_CATCH

void elm_ekf_destroy(elm_ekf* ekf) { delete ekf; }

int elm_ekf_predict_imu(elm_ekf* ekf, double timestamp, const double gyro[3], const double acc[3]) try {
    if (!ekf || !gyro || !acc) return fail(ELM_ERR_INVALID, "elm_ekf_predict_imu: bad argument");
    ELM_CUDA(cudaSetDevice(ekf->device));
    ELM_CUDA(elm::launch_ekf_predict_imu(ekf->d_state, ekf->cfg, timestamp, gyro, acc, ekf->stream));
    if (ekf->ring_on) ELM_CUDA(elm::launch_ekf_ring_push(ekf->d_state, ekf->ring, ekf->stream));
    return ELM_OK;
} ELM_API_CATCH

int elm_ekf_update_pose(elm_ekf* ekf, const elm_ekf_measurement* meas) try {
    if (!ekf || !meas) return fail(ELM_ERR_INVALID, "elm_ekf_update_pose: bad argument");
    if (meas->source != 3 && meas->source != 4) return fail(ELM_ERR_UNSUPPORTED, "only the PCM (3) and PCM_INIT (4) sources are in scope");
    ELM_CUDA(cudaSetDevice(ekf->device));
    ELM_CUDA(elm::launch_ekf_update_pose(ekf->d_state, ekf->cfg, *meas, ekf->stream));
    return ELM_OK;
} ELM_API_CATCH

int elm_ekf_get_state(elm_ekf* ekf, elm_ekf_state* out) try {
    if (!ekf || !out) return fail(ELM_ERR_INVALID, "elm_ekf_get_state: bad argument");
    ELM_CUDA(cudaSetDevice(ekf->device));
    ELM_CUDA(cudaMemcpyAsync(ekf->h_state, ekf->d_state, sizeof(elm_ekf_state), cudaMemcpyDeviceToHost, ekf->stream));
    ELM_CUDA(cudaStreamSynchronize(ekf->stream));
    std::memcpy(out, ekf->h_state, sizeof(elm_ekf_state));
    return ELM_OK;
} ELM_API_CATCH

int elm_ekf_set_state(elm_ekf* ekf, const elm_ekf_state* in) try {
    if (!ekf || !in) return fail(ELM_ERR_INVALID, "elm_ekf_set_state: bad argument");
    ELM_CUDA(cudaSetDevice(ekf->device));
    ELM_CUDA(cudaStreamSynchronize(ekf->stream));
    std::memcpy(ekf->h_state, in, sizeof(elm_ekf_state));
    ELM_CUDA(cudaMemcpyAsync(ekf->d_state, ekf->h_state, sizeof(elm_ekf_state), cudaMemcpyHostToDevice, ekf->stream));
    ELM_CUDA(cudaStreamSynchronize(ekf->stream));
    return ELM_OK;
} ELM_API_CATCH

int elm_ekf_get_current_state(elm_ekf* ekf, double ego[26]) try {
    if (!ekf || !ego) return fail(ELM_ERR_INVALID, "elm_ekf_get_current_state: bad argument");
    ELM_CUDA(cudaSetDevice(ekf->device));
    ELM_CUDA(cudaMemcpyAsync(ekf->h_state, ekf->d_state, sizeof(elm_ekf_state), cudaMemcpyDeviceToHost, ekf->stream));
    ELM_CUDA(cudaStreamSynchronize(ekf->stream));
    if (!elm::ekf_current_state(*ekf->h_state, ego)) {
        // prev_ego_state_ is part of the filter object: keep the device copy of the cache in step
        const size_t off = offsetof(elm_ekf_state, ego);
        ELM_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(ekf->d_state) + off, reinterpret_cast<char*>(ekf->h_state) + off,
                                 27 * sizeof(double), cudaMemcpyHostToDevice, ekf->stream));
        ELM_CUDA(cudaStreamSynchronize(ekf->stream));
    }
    return ELM_OK;
} ELM_API_CATCH

int elm_ekf_enable_state_ring(elm_ekf* ekf, int enable) try {
    if (!ekf) return fail(ELM_ERR_INVALID, "null ekf");
    ELM_CUDA(cudaSetDevice(ekf->device));
    if (enable && !ekf->ring.e) {
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&ekf->ring.e), static_cast<size_t>(ekf->ring.cap) * 8 * sizeof(double)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&ekf->ring.meta), 2 * sizeof(int)));
    }
    if (enable) ELM_CUDA(cudaMemsetAsync(ekf->ring.meta, 0, 2 * sizeof(int), ekf->stream));
    ekf->ring_on = enable != 0;
    return ELM_OK;
} ELM_API_CATCH

// ---- deskew tables (a18) + the device-resident scan chain ------------------------------------------------------------------
int elm_deskew_build_tables(const elm_imu_queue* imu, const elm_odom_queue* odom, elm_deskew_tables* tables, double* storage,
                            size_t* imu_drop, size_t* odom_drop) try {
    if (!imu || !odom || !tables || !storage || (imu->n && (!imu->stamp || !imu->gyro)) ||
        (odom->n && (!odom->stamp || !odom->pos || !odom->quat_xyzw || !odom->lin_vel || !odom->ang_vel)))
        return fail(ELM_ERR_INVALID, "elm_deskew_build_tables: bad argument");
    elm::DeskewTableSet t;
    t.time_scan_cur = tables->time_scan_cur; t.time_scan_end = tables->time_scan_end;
    elm::imu_deskew_info(elm::ImuQueueView{imu->stamp, imu->gyro, imu->n}, t, imu_drop);
    elm::odom_deskew_info(elm::OdomQueueView{odom->stamp, odom->pos, odom->quat_xyzw, odom->lin_vel, odom->ang_vel, odom->n}, t, odom_drop);
    const size_t L = ELM_IMU_QUEUE_LENGTH;
    std::memcpy(storage, t.imu_time.data(), L * sizeof(double)); std::memcpy(storage + L, t.imu_rot_x.data(), L * sizeof(double));
    std::memcpy(storage + 2 * L, t.imu_rot_y.data(), L * sizeof(double)); std::memcpy(storage + 3 * L, t.imu_rot_z.data(), L * sizeof(double));
    tables->imu_time = storage; tables->imu_rot_x = storage + L; tables->imu_rot_y = storage + 2 * L; tables->imu_rot_z = storage + 3 * L;
    tables->imu_pointer_cur = t.imu_pointer_cur; tables->imu_available = t.imu_available ? 1 : 0; tables->odom_available = t.odom_available ? 1 : 0;
    tables->odom_incre_x = t.odom_incre[0]; tables->odom_incre_y = t.odom_incre[1]; tables->odom_incre_z = t.odom_incre[2];
    return ELM_OK;
} ELM_API_CATCH

struct elm_scan_pipeline {
    elm_registration* reg = nullptr;
    elm_scan_pipeline_config cfg{};
    double T_lidar_to_ego[16];
    // device buffers, all sized for the largest raw scan seen
    float* d_raw = nullptr;     // xyz[3 cap] | time[cap]
    float* d_flt = nullptr;     // filtered: xyz | time
    float* d_dsk = nullptr;     // deskewed xyz
    float* d_ds = nullptr;      // down-sampled xyz (what RunRegister sees)
    int* d_counts = nullptr;    // [0] after the filter, [1] after the down-sampling
    int* h_counts = nullptr;    // pinned: counts + the pre-processing error flag
    size_t cap = 0;
    elm::DeskewTableSet tables;
    cudaEvent_t ev_icp = nullptr;
    // last scan
    size_t n_raw = 0;
    int n_filter = 0, n_reg = 0;
    bool deskewed = false, registered = false;
    double max_fitness = 0.0;
    ~elm_scan_pipeline() {
        if (reg) cudaSetDevice(reg->device);
        cudaFree(d_raw); cudaFree(d_flt); cudaFree(d_dsk); cudaFree(d_ds); cudaFree(d_counts); cudaFreeHost(h_counts);
        if (ev_icp) cudaEventDestroy(ev_icp);
    }
};

int elm_scan_pipeline_create(elm_scan_pipeline** out, elm_registration* reg, const elm_scan_pipeline_config* cfg) try {
    if (!out || !reg || !cfg) return fail(ELM_ERR_INVALID, "elm_scan_pipeline_create: bad argument");
    ELM_CUDA(cudaSetDevice(reg->device));
    std::unique_ptr<elm_scan_pipeline> p(new (std::nothrow) elm_scan_pipeline());
    if (!p) return fail(ELM_ERR_INVALID, "out of memory");
    p->reg = reg;
    p->cfg = *cfg;
    {   // tf_ego_to_lidar^-1 (pcm_matching.cpp:298: Matrix4d::inverse(); a rigid transform: cofactor-free closed form would
        // differ in the last bits, so the general inverse is used here too)
        const double* m = cfg->tf_ego_to_lidar;
        double inv[16];
        const double s0 = m[0] * m[5] - m[4] * m[1], s1 = m[0] * m[6] - m[4] * m[2], s2 = m[0] * m[7] - m[4] * m[3];
        const double s3 = m[1] * m[6] - m[5] * m[2], s4 = m[1] * m[7] - m[5] * m[3], s5 = m[2] * m[7] - m[6] * m[3];
        const double c5 = m[10] * m[15] - m[14] * m[11], c4 = m[9] * m[15] - m[13] * m[11], c3 = m[9] * m[14] - m[13] * m[10];
        const double c2 = m[8] * m[15] - m[12] * m[11], c1 = m[8] * m[14] - m[12] * m[10], c0 = m[8] * m[13] - m[12] * m[9];
        const double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
        if (!(std::fabs(det) > 0.0)) return fail(ELM_ERR_INVALID, "elm_scan_pipeline_create: tf_ego_to_lidar is singular");
        const double id = 1.0 / det;
        inv[0] = (m[5] * c5 - m[6] * c4 + m[7] * c3) * id;   inv[1] = (-m[1] * c5 + m[2] * c4 - m[3] * c3) * id;
        inv[2] = (m[13] * s5 - m[14] * s4 + m[15] * s3) * id; inv[3] = (-m[9] * s5 + m[10] * s4 - m[11] * s3) * id;
        inv[4] = (-m[4] * c5 + m[6] * c2 - m[7] * c1) * id;  inv[5] = (m[0] * c5 - m[2] * c2 + m[3] * c1) * id;
        inv[6] = (-m[12] * s5 + m[14] * s2 - m[15] * s1) * id; inv[7] = (m[8] * s5 - m[10] * s2 + m[11] * s1) * id;
        inv[8] = (m[4] * c4 - m[5] * c2 + m[7] * c0) * id;   inv[9] = (-m[0] * c4 + m[1] * c2 - m[3] * c0) * id;
        inv[10] = (m[12] * s4 - m[13] * s2 + m[15] * s0) * id; inv[11] = (-m[8] * s4 + m[9] * s2 - m[11] * s0) * id;
        inv[12] = (-m[4] * c3 + m[5] * c1 - m[6] * c0) * id; inv[13] = (m[0] * c3 - m[1] * c1 + m[2] * c0) * id;
        inv[14] = (-m[12] * s3 + m[13] * s1 - m[14] * s0) * id; inv[15] = (m[8] * s3 - m[9] * s1 + m[10] * s0) * id;
        std::memcpy(p->T_lidar_to_ego, inv, sizeof inv);
    }
    ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_counts), 2 * sizeof(int)));
    ELM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&p->h_counts), 4 * sizeof(int)));
    ELM_CUDA(cudaEventCreateWithFlags(&p->ev_icp, cudaEventDisableTiming));
    *out = p.release();
    return ELM_OK;
} ELM_API_CATCH

void elm_scan_pipeline_destroy(elm_scan_pipeline* p) { delete p; }

namespace {
int ensure_prep_scratch(elm_registration* reg, size_t n) {
    if (n > reg->prep_cap) {
        cudaFree(reg->prep.tkeys); cudaFree(reg->prep.tmin); cudaFree(reg->prep.keep); cudaFree(reg->prep.block_count); cudaFree(reg->prep.block_offset);
        int* err = reg->prep.error;
        reg->prep = elm::ScanPrepScratch{};
        reg->prep.error = err;
        reg->prep_cap = 0;
        const size_t cap = (n + 4095) / 4096 * 4096, slots = elm::scan_prep_table_slots(cap), blocks = (cap + 255) / 256;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.tkeys), slots * sizeof(unsigned long long)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.tmin), slots * sizeof(uint32_t)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.keep), cap));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.block_count), blocks * sizeof(uint32_t)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.block_offset), blocks * sizeof(uint32_t)));
        reg->prep.table_slots = slots;
        reg->prep_cap = cap;
    }
    if (!reg->prep.error) ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->prep.error), sizeof(int)));
    return ELM_OK;
}
// FilterPointsByDistance's test on the host (pcm_matching.cpp:455-458, float arithmetic like the kernel)
bool host_passes_distance(const float* q, double max_dist) {
    if (!(max_dist > 0.0)) return true;
    const float d = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    return !(static_cast<double>(d) > max_dist);
}
}  // namespace

int elm_scan_pipeline_deskew(elm_scan_pipeline* p, const float* xyz, const float* point_time, size_t n, double stamp,
                             const elm_imu_queue* imu, const elm_odom_queue* odom, double* time_scan_cur, double* time_scan_end,
                             int32_t* deskew_ok) try {
    if (!p || !imu || !odom || !deskew_ok || (n && (!xyz || !point_time))) return fail(ELM_ERR_INVALID, "elm_scan_pipeline_deskew: bad argument");
    if (n > 0x7fffffffull / 8) return fail(ELM_ERR_INVALID, "scan too large");
    elm_registration* reg = p->reg;
    *deskew_ok = 0;
    p->deskewed = p->registered = false;
    p->n_raw = n;
    // time base of DeskewPointCloud (:473-486) from the first / last point of the FILTERED cloud
    size_t first = 0, last = n;
    while (first < n && !host_passes_distance(xyz + 3 * first, p->cfg.input_max_dist)) ++first;
    while (last > first && !host_passes_distance(xyz + 3 * (last - 1), p->cfg.input_max_dist)) --last;
    if (first >= last) return ELM_OK;  // nothing survives the filter
    elm::DeskewTableSet& t = p->tables;
    float offset = 0.f;
    t.time_scan_cur = stamp;
    t.time_scan_end = stamp + static_cast<double>(point_time[last - 1]);
    if (p->cfg.lidar_scan_time_end) {
        const double front_time = static_cast<double>(point_time[first]);
        t.time_scan_end = stamp;
        t.time_scan_cur = t.time_scan_end + front_time;
        offset = point_time[first];
    }
    if (time_scan_cur) *time_scan_cur = t.time_scan_cur;
    if (time_scan_end) *time_scan_end = t.time_scan_end;
    elm::imu_deskew_info(elm::ImuQueueView{imu->stamp, imu->gyro, imu->n}, t, nullptr);
    elm::odom_deskew_info(elm::OdomQueueView{odom->stamp, odom->pos, odom->quat_xyzw, odom->lin_vel, odom->ang_vel, odom->n}, t, nullptr);
    if (!t.imu_available || !t.odom_available) return ELM_OK;  // :493-495 "Deskew fail"
    ELM_CUDA(cudaSetDevice(reg->device));
    if (n > p->cap) {
        cudaFree(p->d_raw); cudaFree(p->d_flt); cudaFree(p->d_dsk); cudaFree(p->d_ds);
        p->d_raw = p->d_flt = p->d_dsk = p->d_ds = nullptr;
        p->cap = 0;
        const size_t cap = (n + 4095) / 4096 * 4096;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_raw), cap * 4 * sizeof(float)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_flt), cap * 4 * sizeof(float)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_dsk), cap * 3 * sizeof(float)));
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_ds), cap * 3 * sizeof(float)));
        p->cap = cap;
    }
    int rc = ensure_prep_scratch(reg, n);
    if (rc) return rc;
    // the ONE upload of the scan
    ELM_CUDA(cudaMemcpyAsync(p->d_raw, xyz, n * 3 * sizeof(float), cudaMemcpyHostToDevice, reg->stream));
    ELM_CUDA(cudaMemcpyAsync(p->d_raw + 3 * p->cap, point_time, n * sizeof(float), cudaMemcpyHostToDevice, reg->stream));
    // FilterPointsByDistance: stable compaction, the point times travel along; count -> d_counts[0]
    ELM_CUDA(elm::launch_scan_prep(p->d_raw, p->d_raw + 3 * p->cap, static_cast<int>(n), p->cfg.input_max_dist, 0.0, reg->prep, p->d_flt, p->d_flt + 3 * p->cap,
                                   nullptr, p->d_counts, reg->stream));
    // deskew tables -> device, then the point loop over the survivors (their number is read from HBM)
    const int entries = t.imu_pointer_cur + 1;
    if (entries > reg->dtable_cap) {
        cudaFree(reg->d_dtable); cudaFreeHost(reg->h_dtable);
        reg->d_dtable = nullptr; reg->h_dtable = nullptr;
        const int cap = (entries + 255) / 256 * 256;
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_dtable), static_cast<size_t>(cap) * 4 * sizeof(double)));
        ELM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&reg->h_dtable), static_cast<size_t>(cap) * 4 * sizeof(double)));
        reg->dtable_cap = cap;
    }
    elm::DeskewParams dp{};
    dp.imu_pointer_cur = t.imu_pointer_cur; dp.imu_available = 1; dp.odom_available = 1; dp.table_stride = reg->dtable_cap;
    dp.odom_incre_x = t.odom_incre[0]; dp.odom_incre_y = t.odom_incre[1]; dp.odom_incre_z = t.odom_incre[2];
    dp.time_scan_cur = t.time_scan_cur; dp.time_scan_end = t.time_scan_end;
    dp.n_dev = p->d_counts; dp.rel_time_offset = offset; dp.run_deskew = p->cfg.run_deskew ? 1 : 0;
    {   // (the pinned staging buffer is rewritten only after the previous scan was fetched: elm_scan_pipeline_fetch synchronises)
        const double* rows[4] = {t.imu_time.data(), t.imu_rot_x.data(), t.imu_rot_y.data(), t.imu_rot_z.data()};
        for (int r = 0; r < 4; ++r) std::memcpy(reg->h_dtable + static_cast<size_t>(r) * reg->dtable_cap, rows[r], entries * sizeof(double));
        ELM_CUDA(cudaMemcpyAsync(reg->d_dtable, reg->h_dtable, static_cast<size_t>(reg->dtable_cap) * 4 * sizeof(double), cudaMemcpyHostToDevice, reg->stream));
    }
    ELM_CUDA(elm::launch_deskew_points(p->d_flt, p->d_flt + 3 * p->cap, static_cast<int>(n), dp, reg->d_dtable, p->d_dsk, reg->num_sms, reg->stream));
    p->deskewed = true;
    *deskew_ok = 1;
    return ELM_OK;
} ELM_API_CATCH

int elm_scan_pipeline_register(elm_scan_pipeline* p, const elm_map* map, const double sync_lidar_pose[16], const elm_reg_config* cfg) try {
    if (!p || !map || !sync_lidar_pose || !cfg) return fail(ELM_ERR_INVALID, "elm_scan_pipeline_register: bad argument");
    if (!p->deskewed) return fail(ELM_ERR_STATE, "elm_scan_pipeline_register without a successful elm_scan_pipeline_deskew");
    elm_registration* reg = p->reg;
    ELM_CUDA(cudaSetDevice(reg->device));
    const int n = static_cast<int>(p->n_raw);
    const float* d_scan = p->d_dsk;
    const int* d_n = p->d_counts;
    if (p->cfg.input_voxel_ds_m > 0.0) {  // VoxelDownsample (:256-258) of the survivors, their number read from HBM
        ELM_CUDA(elm::launch_scan_prep(p->d_dsk, nullptr, n, 0.0, p->cfg.input_voxel_ds_m, reg->prep, p->d_ds, nullptr, nullptr, p->d_counts + 1, reg->stream,
                                       p->d_counts));
        d_scan = p->d_ds;
        d_n = p->d_counts + 1;
    }
    // the only read-back before the result: the point counts (the ICP launches are sized on the host)
    ELM_CUDA(cudaMemcpyAsync(p->h_counts, p->d_counts, sizeof(int), cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaMemcpyAsync(p->h_counts + 1, d_n, sizeof(int), cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaMemcpyAsync(p->h_counts + 2, reg->prep.error, sizeof(int), cudaMemcpyDeviceToHost, reg->stream));
    ELM_CUDA(cudaStreamSynchronize(reg->stream));
    if (p->h_counts[2]) return fail(ELM_ERR_RANGE, "scan point not finite or beyond +-2^20 voxels of the down-sampling grid");
    p->n_filter = p->h_counts[0];
    p->n_reg = p->h_counts[1];
    p->max_fitness = cfg->max_fitness_score;
    const int rc = elm_register_enqueue(reg, map, d_scan, static_cast<size_t>(p->n_reg), sync_lidar_pose, cfg);
    if (rc) return rc;
    ELM_CUDA(cudaEventRecord(p->ev_icp, reg->stream));
    p->registered = true;
    return ELM_OK;
} ELM_API_CATCH

int elm_scan_pipeline_ekf_update(elm_scan_pipeline* p, elm_ekf* ekf) try {
    if (!p || !ekf) return fail(ELM_ERR_INVALID, "elm_scan_pipeline_ekf_update: bad argument");
    if (!p->registered) return fail(ELM_ERR_STATE, "elm_scan_pipeline_ekf_update without elm_scan_pipeline_register");
    if (!ekf->ring_on) return fail(ELM_ERR_STATE, "elm_scan_pipeline_ekf_update needs elm_ekf_enable_state_ring");
    if (ekf->device != p->reg->device) return fail(ELM_ERR_INVALID, "filter and registration live on different devices");
    ELM_CUDA(cudaSetDevice(ekf->device));
    if (ekf->stream != p->reg->stream) ELM_CUDA(cudaStreamWaitEvent(ekf->stream, p->ev_icp, 0));
    elm::EkfIcpParams ip{};
    ip.stamp = p->tables.time_scan_end;  // PublishPcmOdom(icp_ego_pose, ros::Time(d_time_scan_end_), ...) :299
    std::memcpy(ip.T_lidar_to_ego, p->T_lidar_to_ego, sizeof ip.T_lidar_to_ego);
    ip.max_fitness = p->max_fitness;
    ip.trivial = p->reg->trivial ? 1 : 0;
    ELM_CUDA(elm::launch_ekf_update_from_icp(ekf->d_state, ekf->cfg, p->reg->d_state, ip, ekf->ring, ekf->stream));
    return ELM_OK;
} ELM_API_CATCH

int elm_scan_pipeline_fetch(elm_scan_pipeline* p, elm_scan_result* out) try {
    if (!p || !out) return fail(ELM_ERR_INVALID, "elm_scan_pipeline_fetch: bad argument");
    if (!p->registered) return fail(ELM_ERR_STATE, "elm_scan_pipeline_fetch without elm_scan_pipeline_register");
    std::memset(out, 0, sizeof *out);
    int32_t ok = 0, it = 0;
    double fit = 0.0;
    const int rc = elm_register_fetch(p->reg, out->T_lidar, &ok, &fit, out->local_cov, &it);
    if (rc) return rc;
    out->is_success = ok; out->iterations = it; out->fitness_score = fit;
    out->n_raw = static_cast<int32_t>(p->n_raw); out->n_after_filter = p->n_filter; out->n_registered = p->n_reg; out->deskew_ok = 1;
    out->time_scan_cur = p->tables.time_scan_cur; out->time_scan_end = p->tables.time_scan_end;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double a = 0.0;
            for (int k = 0; k < 4; ++k) a += out->T_lidar[4 * i + k] * p->T_lidar_to_ego[4 * k + j];
            out->T_ego[4 * i + j] = a;
        }
    if (ok) {
        const double R[9] = {out->T_ego[0], out->T_ego[1], out->T_ego[2], out->T_ego[4], out->T_ego[5], out->T_ego[6], out->T_ego[8], out->T_ego[9], out->T_ego[10]};
        elm::shape_pcm_covariance_hd(R, out->local_cov, fit, out->pose_cov);
    }
    p->registered = false;
    return ELM_OK;
} ELM_API_CATCH

int elm_comm_unique_id(uint8_t unique_id[128]) try {
    if (!unique_id) return fail(ELM_ERR_INVALID, "null id");
    if (!g_nccl.load()) return fail(ELM_ERR_NCCL, "libnccl.so.2 could not be loaded");
    const int e = g_nccl.GetUniqueId(unique_id);
    if (e != 0) return fail(ELM_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(e));
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_set_comm(elm_registration* reg, const uint8_t unique_id[128], int rank, int world_size) try {
    if (!reg || !unique_id || world_size < 1 || rank < 0 || rank >= world_size) return fail(ELM_ERR_INVALID, "elm_registration_set_comm: bad argument");
    if (!g_nccl.load()) return fail(ELM_ERR_NCCL, "libnccl.so.2 could not be loaded");
    ELM_CUDA(cudaSetDevice(reg->device));
    if (reg->comm) { g_nccl.CommDestroy(reg->comm); reg->comm = nullptr; }
    NcclApi::Uid id;
    std::memcpy(id.b, unique_id, 128);
    const int e = g_nccl.CommInitRank(&reg->comm, world_size, id, rank);
    if (e != 0) { reg->comm = nullptr; return fail(ELM_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(e)); }
    reg->rank = rank;
    reg->world = world_size;
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_peer_export(elm_registration* reg, uint8_t handle[64]) try {
    if (!reg || !handle) return fail(ELM_ERR_INVALID, "elm_registration_peer_export: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    ELM_CUDA(cudaSetDevice(reg->device));
    if (!reg->d_mailbox) {
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_mailbox), sizeof(elm::PeerMailbox)));
        ELM_CUDA(cudaMemset(reg->d_mailbox, 0, sizeof(elm::PeerMailbox)));
    }
    cudaIpcMemHandle_t h;
    ELM_CUDA(cudaIpcGetMemHandle(&h, reg->d_mailbox));
    std::memcpy(handle, &h, 64);
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_peer_attach(elm_registration* reg, const uint8_t* handles, int rank, int world_size) try {
    if (!reg || world_size < 1 || world_size > elm::kMaxPeers || rank < 0 || rank >= world_size || (world_size > 1 && !handles))
        return fail(ELM_ERR_INVALID, "elm_registration_peer_attach: bad argument (1 <= world_size <= 8)");
    ELM_CUDA(cudaSetDevice(reg->device));
    if (!reg->d_mailbox) {
        ELM_CUDA(cudaMalloc(reinterpret_cast<void**>(&reg->d_mailbox), sizeof(elm::PeerMailbox)));
    }
    ELM_CUDA(cudaStreamSynchronize(reg->stream));
    ELM_CUDA(cudaMemset(reg->d_mailbox, 0, sizeof(elm::PeerMailbox)));
    for (void*& p : reg->peer_opened) { if (p) cudaIpcCloseMemHandle(p); p = nullptr; }
    elm::PeerComm pc{};
    pc.rank = rank;
    pc.world = world_size;
    for (int r = 0; r < world_size; ++r) {
        if (r == rank) { pc.box[r] = reg->d_mailbox; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + 64 * r, 64);
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ELM_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
        }
        reg->peer_opened[r] = ptr;
        pc.box[r] = static_cast<elm::PeerMailbox*>(ptr);
    }
    reg->peer = pc;
    reg->rank = rank;
    reg->world = world_size;
    return ELM_OK;
} ELM_API_CATCH

int elm_registration_peer_detach(elm_registration* reg) try {
    if (!reg) return fail(ELM_ERR_INVALID, "null registration");
    ELM_CUDA(cudaSetDevice(reg->device));
    ELM_CUDA(cudaStreamSynchronize(reg->stream));
    for (void*& p : reg->peer_opened) { if (p) cudaIpcCloseMemHandle(p); p = nullptr; }
    reg->peer = elm::PeerComm{};
    return ELM_OK;
} ELM_API_CATCH

}  // extern "C"
