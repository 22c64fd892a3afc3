// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.hpp header).
// Parity status: PINNED on oracle/_ref — the reference's own sources compiled against stand-in Eigen/oneTBB headers
// (tests/test_reference_build.py); Eigen's arithmetic kernels themselves stay restated (smallmat.hpp).
// Restates /root/reference/src/app/localization/pcm_matching/src/voxel_hash_map.cpp and the inline
// bodies of .../include/voxel_hash_map.hpp; each function cites the lines it follows.
#include "voxel_map.hpp"

#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// vhm.hpp:106-113 — keep iff below the cap and no stored point closer than map_resolution (norm, strict <).
void VoxelHashMap::VoxelBlock::AddPointWithSpacing(const PointStruct& point) {
    if (points.size() >= static_cast<size_t>(num_points)) return;
    for (const auto& vp : points)
        if (norm(vp.pose - point.pose) < map_resolution) return;
    points.push_back(point);
}

// vhm.hpp:114-148
void VoxelHashMap::VoxelBlock::CalVoxelCov() {
    const int n = static_cast<int>(points.size());
    covariance.cov = M3::Identity();
    covariance.mean = V3();
    if (n == 0) return;
    if (n == 1) { covariance.mean = points[0].pose; return; }
    V3 sum;
    for (int j = 0; j < n; ++j) sum = sum + points[j].pose;
    const V3 mean(sum.x / n, sum.y / n, sum.z / n);
    M3 cov;
    for (int j = 0; j < n; ++j) {
        const V3 d = points[j].pose - mean;
        const double dv[3] = {d.x, d.y, d.z};
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) cov(a, b) += dv[a] * dv[b];
    }
    for (int i = 0; i < 9; ++i) cov.m[i] /= (n - 1);
    covariance.cov = plane_regularize(cov);  // vhm.hpp:141-144
    covariance.mean = mean;
}

// vhm.cpp:270-285 — sequential, order-dependent.  Insert key = static_cast<int>(p / vs): truncation toward zero (Q1).
void VoxelHashMap::AddPoints(const std::vector<PointStruct>& points) {
    if (points.empty()) return;
    const double map_resolution = std::sqrt(voxel_size_ * voxel_size_ / max_points_per_voxel_);
    for (const auto& point : points) {
        const Voxel voxel{static_cast<int>(point.pose.x / voxel_size_), static_cast<int>(point.pose.y / voxel_size_),
                          static_cast<int>(point.pose.z / voxel_size_)};
        auto search = map_.find(voxel);
        if (search != map_.end()) {
            search->second.AddPointWithSpacing(point);
        } else {
            VoxelBlock vb;
            vb.points.push_back(point);
            vb.num_points = max_points_per_voxel_;
            vb.map_resolution = map_resolution;
            map_.insert({voxel, std::move(vb)});
        }
    }
}

// vhm.hpp:183-193
void VoxelHashMap::CalVoxelCovAll() {
    std::vector<VoxelBlock*> blocks;
    blocks.reserve(map_.size());
    for (auto& kv : map_) blocks.push_back(&kv.second);
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < static_cast<long>(blocks.size()); ++i) blocks[i]->CalVoxelCov();
}

// vhm.hpp:195-257 — neighbours = {self} U {every stored point of the 27 voxels with d^2 <= r^2}; the
// point itself has d = 0 so it is counted twice (Q5) and the n == 1 branch (vhm.hpp:225-228) is dead.
void VoxelHashMap::CalPointCovAll(double d_search_dist) {
    const double r2 = d_search_dist * d_search_dist;
    std::vector<VoxelBlock*> blocks;
    blocks.reserve(map_.size());
    for (auto& kv : map_) blocks.push_back(&kv.second);
    // Two passes so that no thread reads a covariance another thread is writing (only poses are read).
#pragma omp parallel for schedule(dynamic, 64)
    for (long bi = 0; bi < static_cast<long>(blocks.size()); ++bi) {
        std::vector<V3> neighbors;
        for (auto& point : blocks[bi]->points) {
            neighbors.clear();
            neighbors.push_back(point.pose);
            Voxel adj[27];
            const int na = GetAdjacentVoxels(point.pose, 2, adj);
            for (int a = 0; a < na; ++a) {
                auto it = map_.find(adj[a]);
                if (it == map_.end()) continue;
                for (const auto& np : it->second.points)
                    if (sqnorm(np.pose - point.pose) <= r2) neighbors.push_back(np.pose);
            }
            if (neighbors.size() == 1) {
                point.covariance.cov = M3::Identity();
                point.covariance.mean = point.pose;
            } else {
                const double n = static_cast<double>(neighbors.size());
                V3 sum;
                for (const auto& q : neighbors) sum = sum + q;
                const V3 mean(sum.x / n, sum.y / n, sum.z / n);
                M3 cov;
                for (const auto& q : neighbors) {
                    const V3 d = q - mean;
                    const double dv[3] = {d.x, d.y, d.z};
                    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) cov(a, b) += dv[a] * dv[b];
                }
                for (int i = 0; i < 9; ++i) cov.m[i] /= (n - 1.0);
                point.covariance.cov = plane_regularize(cov);  // vhm.hpp:241-244
                point.covariance.mean = mean;
            }
        }
    }
}

// vhm.cpp:208-243 — range 2: 27 voxels, x outer / y / z inner; range 1: centre,+x,-x,+y,-y,+z,-z; range 0: centre.
int VoxelHashMap::GetAdjacentVoxels(const V3& pose, int range, Voxel out[27]) const {
    const Voxel v = PointToVoxel(pose, voxel_size_);
    if (range == 0) { out[0] = v; return 1; }
    if (range == 1) {
        out[0] = Voxel{v.x, v.y, v.z};
        out[1] = Voxel{v.x + 1, v.y, v.z};
        out[2] = Voxel{v.x - 1, v.y, v.z};
        out[3] = Voxel{v.x, v.y + 1, v.z};
        out[4] = Voxel{v.x, v.y - 1, v.z};
        out[5] = Voxel{v.x, v.y, v.z + 1};
        out[6] = Voxel{v.x, v.y, v.z - 1};
        return 7;
    }
    int n = 0;
    for (int i = v.x - 1; i < v.x + 2; ++i)
        for (int j = v.y - 1; j < v.y + 2; ++j)
            for (int k = v.z - 1; k < v.z + 2; ++k) out[n++] = Voxel{i, j, k};
    return n;
}

namespace {
template <class Result, class Body>
Result chunked_reduce(size_t n, int nthreads, Body body) {
    // tbb::parallel_reduce stand-in: contiguous chunks, per-chunk result, joined left to right (vhm.cpp:74-84).
    if (nthreads < 1) nthreads = 1;
    std::vector<Result> parts(nthreads);
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
    for (int t = 0; t < nthreads; ++t) {
        const size_t b = n * t / nthreads, e = n * (t + 1) / nthreads;
        body(b, e, parts[t]);
    }
    Result out = std::move(parts[0]);
    for (int t = 1; t < nthreads; ++t) {
        out.first.insert(out.first.end(), std::make_move_iterator(parts[t].first.begin()),
                         std::make_move_iterator(parts[t].first.end()));
        out.second.insert(out.second.end(), std::make_move_iterator(parts[t].second.begin()),
                          std::make_move_iterator(parts[t].second.end()));
    }
    return out;
}
}  // namespace

// vhm.cpp:31-88 — P2P and GICP search.
std::tuple<std::vector<PointStruct>, std::vector<PointStruct>> VoxelHashMap::GetCorrespondencePoints(
    const std::vector<PointStruct>& pts, double max_dist, int nthreads) const {
    const double max2 = max_dist * max_dist;
    using R = std::pair<std::vector<PointStruct>, std::vector<PointStruct>>;
    R res = chunked_reduce<R>(pts.size(), nthreads, [&](size_t b, size_t e, R& r) {
        for (size_t i = b; i < e; ++i) {
            const PointStruct& point = pts[i];
            Voxel adj[27];
            const int na = GetAdjacentVoxels(point.pose, 2, adj);
            PointStruct closest;  // default: pose 0, cov (I, 0)  — Q2 (vhm.cpp:37)
            double best = std::numeric_limits<double>::max();
            for (int a = 0; a < na; ++a) {
                auto it = map_.find(adj[a]);
                if (it == map_.end()) continue;
                for (const auto& nb : it->second.points) {
                    const double d2 = sqnorm(nb.pose - point.pose);
                    if (d2 < best) { closest = nb; best = d2; }  // strict <, full struct copy (vhm.cpp:45-48)
                }
            }
            if (sqnorm(closest.pose - point.pose) < max2) {  // vhm.cpp:66
                r.first.emplace_back(point);
                r.second.emplace_back(closest);
            }
        }
    });
    return std::make_tuple(std::move(res.first), std::move(res.second));
}

// vhm.cpp:90-151 — VGICP search: nearest voxel MEAN among the non-empty voxels of the 27.
std::tuple<std::vector<PointStruct>, std::vector<CovStruct>> VoxelHashMap::GetCorrespondencesCov(
    const std::vector<PointStruct>& pts, double max_dist, int nthreads) const {
    const double max2 = max_dist * max_dist;
    using R = std::pair<std::vector<PointStruct>, std::vector<CovStruct>>;
    R res = chunked_reduce<R>(pts.size(), nthreads, [&](size_t b, size_t e, R& r) {
        for (size_t i = b; i < e; ++i) {
            const PointStruct& point = pts[i];
            Voxel adj[27];
            const int na = GetAdjacentVoxels(point.pose, 2, adj);
            CovStruct closest;  // default (I, 0) — Q2 (vhm.cpp:104)
            double best = std::numeric_limits<double>::max();
            for (int a = 0; a < na; ++a) {
                auto it = map_.find(adj[a]);
                if (it == map_.end() || it->second.points.empty()) continue;
                const CovStruct& c = it->second.covariance;
                const double d2 = sqnorm(c.mean - point.pose);
                if (d2 < best) { closest = c; best = d2; }
            }
            if (sqnorm(closest.mean - point.pose) < max2) {  // vhm.cpp:129
                r.first.emplace_back(point);
                r.second.emplace_back(closest);
            }
        }
    });
    return std::make_tuple(std::move(res.first), std::move(res.second));
}

// vhm.cpp:153-206 — AVGICP search: EVERY non-empty voxel of the 7-neighbourhood within range (Q6).
std::tuple<std::vector<PointStruct>, std::vector<CovStruct>> VoxelHashMap::GetCorrespondencesAllCov(
    const std::vector<PointStruct>& pts, double max_dist, int nthreads) const {
    const double max2 = max_dist * max_dist;
    using R = std::pair<std::vector<PointStruct>, std::vector<CovStruct>>;
    R res = chunked_reduce<R>(pts.size(), nthreads, [&](size_t b, size_t e, R& r) {
        for (size_t i = b; i < e; ++i) {
            const PointStruct& point = pts[i];
            Voxel adj[27];
            const int na = GetAdjacentVoxels(point.pose, 1, adj);
            for (int a = 0; a < na; ++a) {
                auto it = map_.find(adj[a]);
                if (it == map_.end() || it->second.points.empty()) continue;
                const CovStruct& c = it->second.covariance;
                if (sqnorm(c.mean - point.pose) < max2) {  // vhm.cpp:183
                    r.first.emplace_back(point);
                    r.second.emplace_back(c);
                }
            }
        }
    });
    return std::make_tuple(std::move(res.first), std::move(res.second));
}

// vhm.hpp:260-283
std::vector<PointStruct> VoxelHashMap::VoxelDownsample(const std::vector<PointStruct>& pts, double voxel_size) const {
    std::unordered_map<Voxel, size_t, VoxelHash> grid;
    grid.reserve(pts.size());
    std::vector<PointStruct> out;
    for (size_t i = 0; i < pts.size(); ++i) {
        const Voxel v = PointToVoxel(pts[i].pose, voxel_size);
        if (grid.find(v) == grid.end()) { grid.insert({v, i}); out.push_back(pts[i]); }
    }
    return out;
}

size_t VoxelHashMap::NumPoints() const {
    size_t n = 0;
    for (const auto& kv : map_) n += kv.second.points.size();
    return n;
}

}  // namespace orc
