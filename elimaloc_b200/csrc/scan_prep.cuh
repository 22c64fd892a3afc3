// Launch interface of scan_prep.cu (distance filter + first-in-voxel down-sampling of a scan, stable compaction)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace elm {

struct ScanPrepScratch {
    unsigned long long* tkeys;  // table_slots packed voxel keys (all ones = empty)
    uint32_t* tmin;             // table_slots: smallest input index of the voxel
    size_t table_slots;         // power of two >= 2 n
    uint8_t* keep;              // n survivor flags
    uint32_t* block_count;      // ceil(n / 256)
    uint32_t* block_offset;     // ceil(n / 256)
    int* error;                 // set to 1 when a point's voxel key is not packable (not finite / beyond +-2^20 voxels)
};

size_t scan_prep_table_slots(size_t n);

// max_dist <= 0: no distance filter; voxel_size <= 0: no down-sampling.  aux / aux_out (one float per point, e.g. the
// relative time stamp) and index_out (input index of every survivor) may be NULL.  n_out: device int.
// n_dev (may be NULL): the real number of input points sits in HBM (the count an earlier compaction left there) and n is only
// its upper bound — the grids are sized for n, the kernels read *n_dev: a chain of stages needs no host round trip for the counts.
cudaError_t launch_scan_prep(const float* xyz, const float* aux, int n, double max_dist, double voxel_size, const ScanPrepScratch& w,
                             float* xyz_out, float* aux_out, int* index_out, int* n_out, cudaStream_t s, const int* n_dev = nullptr);

}  // namespace elm
