// Host-side (C++) builder of the voxel map: the product's equivalent of the reference's
// VoxelHashMap::AddPoints / CalVoxelCovAll / CalPointCovAll
//   (pcm_matching/src/voxel_hash_map.cpp:270-285, include/voxel_hash_map.hpp:106-148,183-257)
// re-designed for a device-resident layout: instead of an unordered_map of per-voxel vectors the map is
// kept CANONICALLY SORTED — voxels ordered by key (x, then y, then z), points in insertion order inside a
// voxel — which makes the three z-neighbours of a voxel column one contiguous run in memory, and an
// open-addressed hash table (64-bit packed key -> {first point, count}) is built over it for upload.
// The order-dependent AddPointWithSpacing rule is preserved exactly: a stable sort by voxel keeps insertion
// order, and voxels are independent of each other, so the per-voxel sequential filter runs in parallel.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "voxel_key.hpp"

namespace elm {

// One open-addressed slot, 16 bytes: {key lo, key hi, first stored point, stored count}.
struct Slot { uint32_t key_lo, key_hi, start, count; };

// Neighbourhood directory (what the P2P/GICP search reads): one entry per voxel key whose 27-neighbourhood holds at
// least one stored point (occupied voxels + their one-voxel halo).  Entry = a 16-byte slot in a 2-choice cuckoo table
// with 2-slot buckets {key, first point and counts of the centre z-column} + a 320-byte row, indexed by the SLOT position,
// of nine column records {first point, n(z-1) | n(z) << 10 | n(z+1) << 20, octant words} for the columns (x+dx, y+dy), dx
// outer (layout: voxel_key.hpp).
// A lookup is two independent 32-byte loads (no probe chains); a miss means "no candidate at all".
struct DirSlot { uint32_t key_lo, key_hi, first, counts; };
struct DirDesc { uint32_t first, counts; };

struct HostMap {
    double voxel_size = 1.0;
    int cap = 30;

    // canonical sorted storage
    std::vector<uint64_t> vkey;     // V packed keys, ascending
    std::vector<uint32_t> vstart;   // V+1 prefix of stored points
    std::vector<float> pxyz;        // 3P
    std::vector<uint32_t> porig;    // P: running insertion index of the raw point (diagnostics)
    uint64_t n_raw_seen = 0;

    // covariances (row-major 3x3)
    bool has_vcov = false, has_pcov = false;
    std::vector<double> vmean, vcov;            // 3V, 9V
    std::vector<double> pmean, pcov, pnormal;   // 3P, 9P, 3P

    // open-addressed table over vkey (load <= 0.5); slot i <-> voxel slot_voxel[i] (or -1)
    std::vector<Slot> slots;
    std::vector<int32_t> slot_voxel;
    uint32_t mask = 0;

    // neighbourhood directory (see above); dir_slots.size() == 2 * (dir_bmask + 1), dir_rows.size() == kRowWords * dir_slots.size()
    std::vector<DirSlot> dir_slots;
    std::vector<uint32_t> dir_rows;
    DirDesc row_column(size_t slot, int c) const { const size_t w = row_col_word(slot, c); return DirDesc{dir_rows[w], dir_rows[w + 1]}; }
    // Device order of the stored points: inside every voxel sorted by octant (voxel_key.hpp), stable in the canonical order.
    // dev_order[d] = canonical index of the point at device position d (a permutation inside each voxel's range);
    // voct = the 8 octant-word bytes of every voxel.  The canonical arrays above (what export / save / the covariance passes
    // see) never change order; only the upload is permuted.
    std::vector<uint32_t> dev_order;
    std::vector<uint8_t> voct;
    bool octants = false;  // octant words valid (cap <= 255)
    uint32_t dir_bmask = 0;
    size_t dir_entries = 0;
    // VGICP / AVGICP candidates (built by cal_voxel_cov): for every directory entry the non-empty voxels of its
    // 27-neighbourhood in the reference's visit order (x outer, y, z inner), one 8-byte record each (pack_vcand, voxel_key.hpp:
    // the voxel's mean relative to the entry's key in 13-bit fixed point — enough for the search's pre-filter, the exact fp64
    // mean decides near ties — and the voxel's index).  Row header of the entry: {first candidate, count}, 27-bit occupancy
    // mask (bit 9 (dx+1) + 3 (dy+1) + (dz+1)).
    std::vector<uint64_t> vcand8;
    // AVGICP: per directory SLOT the voxel indices of {centre, +x, -x, +y, -y, +z, -z} (vhm.cpp:224-230) or -1, padded to 8
    std::vector<int32_t> dir7;

    size_t V() const { return vkey.size(); }
    size_t P() const { return pxyz.size() / 3; }

    // returns "" on success, else an error message (range violation)
    std::string add_points(const float* xyz, size_t n);
    // the same result delivered by the GPU builder (map_build.cu) for an EMPTY map: takes the canonical arrays over and builds the
    // derived tables (voxel table, neighbourhood directory, octant order) exactly as add_points does after its own compaction
    std::string adopt_canonical(std::vector<uint64_t>& nkey, std::vector<uint32_t>& nstart, std::vector<float>& nxyz, std::vector<uint32_t>& norig, size_t n_raw);
    // covariances computed elsewhere (GPU builder): take them over; cal_voxel_cov's tail (candidate lists)
    void adopt_voxel_cov(std::vector<double>& mean, std::vector<double>& cov);
    void adopt_point_cov(std::vector<double>& mean, std::vector<double>& cov, std::vector<double>& normal);
    void cal_voxel_cov();
    void cal_point_cov(double search_dist);
    void build_table();
    void build_voxel_candidates();
    void build_octants();
    // returns "" on success
    std::string build_directory();
    // slot index of a centre key in the directory or -1 (host mirror of the device lookup; tests)
    int64_t dir_find(uint64_t key) const;
    // voxel index of a packed key or -1
    int64_t find(uint64_t key) const;
    // VoxelHashMap::FindGroundHeight (voxel_hash_map.hpp:285-322): mean z of the (up to) five lowest stored points within 5 m
    // of (x, y) in the plane; false when three or fewer points are in range.  Walks only the voxel columns that can reach
    // the disc instead of copying the whole map as the reference does.
    bool find_ground_height(double x, double y, double& ground_z) const;
    // Built-map file (the reference reloads the .pcd and rebuilds the whole map at every start,
    // pcm_matching.cpp:69-101): every array above, little-endian, behind a header with magic / version / sizes.
    // Both return "" on success, else an error message.
    std::string save(const std::string& path) const;
    std::string load(const std::string& path);
};

// Symmetric 3x3 "plane regularisation" used by both covariance passes (voxel_hash_map.hpp:141-144, 241-244):
// returns I - (1 - 1e-3) n n^T with n the unit eigenvector of the smallest eigenvalue (== U diag(1,1,1e-3) V^T of the
// reference's JacobiSVD for a symmetric PSD input).  Degenerate smallest pair: convention documented in DESIGN.md.
void plane_regularize(const double cov[9], double out[9], double normal[3]);

}  // namespace elm
