// Device-side data layout of the registration hot path (sm_100a).
//
// HBM layout of the map (built by host_map.cpp, uploaded by device_map.cu):
//   dslots  uint4[2 B]           neighbourhood directory: 2-choice cuckoo table, B = 2^k buckets of two 16-B slots
//                                {key_lo, key_hi, first point, counts} of every voxel key whose 27-neighbourhood holds a
//                                stored point; first/counts describe the centre z-column (x, y, z-1..z+1);
//                                counts = n(z-1) | n(z) << 10 | n(z+1) << 20; key = 3 x 21-bit biased coordinates
//   drows   uint32[80 * 2 B]     one 320-B row per SLOT (layout: voxel_key.hpp): header {key, VGICP candidate run, occupancy
//                                mask, flags} + nine 32-B column records {first point, counts, three octant words} of the
//                                z-columns (x+dx, y+dy), dx outer
//   pts     float4[P]            stored points, voxels in canonical order (sorted by (x,y,z)): the voxels (x,y,z-1..z+1) of one
//                                column are ONE contiguous run; inside a voxel sorted by OCTANT (z-half major), insertion order
//                                inside an octant; w = bits of the point's canonical index (voxel order, then insertion order)
//                                = its rank in the reference's visit order, which decides exact distance ties
//   prec    double[16 P]         GICP record per stored point (same order as pts): mean[3] cov[9] normal[3] pad  (128 B, one line)
//   vrec    double[16 V]         VGICP/AVGICP: one 128-byte line per voxel (canonical voxel order) {mean[3], cov[9], pad[4]}: the
//                                accumulation reads mean + covariance of a correspondence in ONE DRAM burst
//   vcand8  uint64[C]            VGICP candidates: for every directory entry the non-empty voxels of its 27-neighbourhood in
//                                visit order, 8 bytes each {3 x 13-bit mean relative to the entry's key, 25-bit voxel index}
//                                (voxel_key.hpp); the row header holds {first candidate, count} and the 27-bit occupancy mask
//   dir7    int32[8 S]           AVGICP: voxel indices of {c, +x, -x, +y, -y, +z, -z} of every directory slot or -1
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace elm {

struct MapView {
    const uint4* dslots;
    const uint32_t* drows;
    const float4* pts;
    const double* prec;
    const double* vrec;                 // VGICP/AVGICP: 16 doubles (one 128-byte line) per voxel: mean[3] cov[9] pad[4]
    const unsigned long long* vcand8;   // VGICP: 8-byte candidate records (voxel_key.hpp: pack_vcand)
    const int* dir7;  // AVGICP: 8 ints per directory slot: voxel indices of {c, +x, -x, +y, -y, +z, -z} or -1
    uint32_t bmask;  // directory buckets - 1
    double voxel_size;
    double inv_voxel_size;  // 1 / voxel_size when that is exact (voxel_size a power of two: p * inv == p / vs bit for bit), else 0
};

constexpr int kAcc = 32;       // accumulator vector: 21 JtJ (upper, row-major) + 6 Jtr + residual + n_corr + n_total + pad
constexpr int kAccUsed = 30;
constexpr int kIdxJtr = 21, kIdxRes = 27, kIdxNcorr = 28, kIdxNtotal = 29;

// Peer-memory exchange of the accumulators (multi-GPU): every rank owns one mailbox in its HBM, mapped into the other
// ranks' address spaces (CUDA IPC over NVLink).  Rank r writes its 32 sums straight into slot [parity][r] of EVERY
// mailbox; each rank waits until all slots of its own mailbox carry the current sequence number and sums them in rank
// order — bit-identical on all ranks, no broadcast, no separate collective launch.  The slots are self-validating
// ("LL" style): every fp64 value travels as two 8-byte words {32 data bits, 32-bit sequence tag}; an 8-byte store is
// atomic, so a word whose tag matches carries valid data — no fence, no separate flag write, ONE NVLink hop.
constexpr int kMaxAsyncIterations = 1024;  // concurrent-refresh iterations per call that own a pair of tile counters (beyond: pair mode)
constexpr int kMaxPeers = 8;
struct PeerMailbox {
    unsigned long long word[2][kMaxPeers][kAcc][2];  // [parity][source rank][accumulator][low / high half]: data | tag << 32
    unsigned long long seq;  // exchanges completed by the owner (all ranks advance in lockstep)
};
struct PeerComm {
    PeerMailbox* box[kMaxPeers];  // box[rank] is the local one
    int rank;
    int world;                    // 0: no peer exchange
};

struct IcpParams {
    int method;
    int n;              // scan points of THIS rank
    int queries_per_warp;
    double max_dist2;   // max_search_dist^2
    double th;          // trans_th == max_search_dist (reg.cpp:360-373)
    double lm_lambda;
    double term_thr;
    double min_overlap;
    double warm_margin;  // metres added to the previous match's distance when a warm search refreshes its remembered runs
    PeerComm peer;
    unsigned long long* stats;  // optional: [0] += map points visited by the search, [1] += queries (NULL = off)
};

// Per-registration device scratch of the ICP loop (owned by elm_registration, sized for the largest scan seen).
struct IcpWork {
    int* match;             // [n] device index of the matched map point (P2P/GICP) / voxel slot (VGICP); -1 = none
    float4* win;            // [n] P2P/GICP: the matched map point itself {x, y, z, bits of its device index}; none ->
                            //     {0, 0, 0, 0xffffffff} = the reference's default-constructed neighbour at the origin (Q2), so the
                            //     accumulation STREAMS its targets instead of gathering pts[match[i]]
    uint4* memo;            // warm start of the next iteration's search, two planes of memo_stride elements:
                            //   [0] {directory row, key_lo, key_hi, device index of the match}
                            //   [1] {q0.x, q0.y, q0.z, R} (floats): the candidate list holds every point of the 27 voxels within R of q0
    size_t memo_stride;     // = elements per plane of memo, cand, cidx
    uint32_t* ncand;        // [n] length of the query's candidate list; ~0: unusable
    float4* cand;           // candidate j of query i at cand[(i / 256) * cand_cap * 256 + j * 256 + i % 256] (tile-interleaved: the lists
                            // of 256 consecutive queries form one block, candidate-major inside): a COPY of the stored point
                            // {x, y, z, bits of its device index}
    int cand_cap;
    uint32_t* refresh_list;   // queries the reuse kernel of a warm iteration hands to its refresh kernel: one segment of 256 entries
                              // per tile of 256 queries, filled in thread order
    uint32_t* refresh_count;  // [tiles] entries used in each segment
    double* partials;       // [blocks][kAcc] per-block sums of one linearisation
    unsigned int* ticket;   // blocks finished
    // concurrent refresh (icp_warm_refresh_async_kernel runs BESIDE the reuse kernel of the same iteration):
    unsigned long long* tile_flag;   // [tiles] {epoch << 32 | stragglers of the tile; all ones in the low word = loop already left}, published by the
                                     // reuse kernel as soon as the tile's work list is complete
    double* tile_rows;               // [chunks of 16 tiles][kAcc] sums of the chunk's refreshed correspondences
    unsigned long long* tile_ticket; // THIS iteration's pair of counters: [0] chunks handed out, [1] chunk rows completed (icp_begin_kernel zeroes
                                     // the pairs of all kMaxAsyncIterations iterations of a call)
    unsigned int epoch;              // this iteration's epoch (0: no flags are published)
};

// Lives in HBM for the whole ICP loop; the host reads it back once at the end.
struct IcpState {
    double T[16];       // last_icp_pose
    double Tinv[16];    // last_icp_pose.inverse()
    double Rinv[9];     // last_icp_pose.block<3,3>(0,0).inverse()
    double acc[kAcc];   // reduced accumulators of the current iteration (allreduced across ranks)
    double JTJ[36];     // last linearisation, full symmetric (test hook)
    double JTr[6];
    double residual_sum;
    double n_corr;
    double fitness;     // Registration::d_fitness_score_ (persists across calls, reg.hpp:229)
    double local_cov[36];
    int iterations;     // AlignClouds* calls executed
    int done;           // loop left (termination, overlap failure)
    int overlap_fail;   // reg.cpp:352-356
    int comm_error;     // the peer exchange timed out (a rank never arrived)
};

}  // namespace elm
