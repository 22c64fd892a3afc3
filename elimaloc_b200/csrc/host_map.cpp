// Host-side voxel-map builder (see host_map.hpp).  Reference behaviour followed:
//   AddPoints            pcm_matching/src/voxel_hash_map.cpp:270-285
//   AddPointWithSpacing  pcm_matching/include/voxel_hash_map.hpp:106-113
//   CalVoxelCov          pcm_matching/include/voxel_hash_map.hpp:114-148
//   ProcessVoxelBlock    pcm_matching/include/voxel_hash_map.hpp:195-250
#include "host_map.hpp"

#include "cov_math.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <atomic>
#include <thread>
#include <chrono>
#include <cstdlib>

namespace elm {

namespace {

template <class F>
void parallel_for(size_t n, size_t grain, F f) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    size_t nt = std::min<size_t>(hw, (n + grain - 1) / std::max<size_t>(grain, 1));
    if (nt <= 1) { f(0, n); return; }
    std::vector<std::thread> th;
    th.reserve(nt);
    for (size_t t = 0; t < nt; ++t) {
        const size_t b = n * t / nt, e = n * (t + 1) / nt;
        th.emplace_back([=]() { f(b, e); });
    }
    for (auto& x : th) x.join();
}

// mean + sample covariance /(n-1) of a multiset of positions, then the regularisation.
void mean_cov_regularized(const double* pts, size_t n, double mean[3], double cov_out[9], double normal[3]) {
    double s[3] = {0, 0, 0};
    for (size_t i = 0; i < n; ++i) { s[0] += pts[3 * i]; s[1] += pts[3 * i + 1]; s[2] += pts[3 * i + 2]; }
    const double dn = static_cast<double>(n);
    mean[0] = s[0] / dn; mean[1] = s[1] / dn; mean[2] = s[2] / dn;
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; ++i) {
        const double d[3] = {pts[3 * i] - mean[0], pts[3 * i + 1] - mean[1], pts[3 * i + 2] - mean[2]};
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) c[a * 3 + b] += d[a] * d[b];
    }
    for (int i = 0; i < 9; ++i) c[i] /= (dn - 1.0);
    plane_regularize(c, cov_out, normal);
}

}  // namespace

void plane_regularize(const double cov[9], double out[9], double normal[3]) { plane_regularize_hd(cov, out, normal); }

namespace {
// ELM_BUILD_TIMING=1: stage times of the host builder on stderr (developer aid)
struct StageTimer {
    bool on = std::getenv("ELM_BUILD_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[host map] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};
}  // namespace

// Parallel, bit-identical restatement of the sequential insert: points are partitioned into x-slabs of voxels (the packed
// key's most significant field, so slab order is key order), every slab is sorted by (voxel, arrival) and filtered
// independently — voxels never interact (AddPointWithSpacing looks only at the points already kept in the SAME voxel), and
// inside a voxel the arrival order is preserved — and the kept points of the slabs are concatenated.  Records carry their
// coordinates, so after the partition pass every stage streams through memory.
std::string HostMap::add_points(const float* xyz, size_t n) {
    if (n == 0) return "";
    StageTimer timer;
    if (cap > static_cast<int>(kDirCountMask)) return "max_points_per_voxel above 1023 does not fit the column descriptors";
    const size_t P0 = P();
    const size_t total = P0 + n;
    if (total >= (1ull << 32)) return "more than 2^32 points";
    struct Rec { uint64_t key; uint32_t idx; float x, y, z; };  // idx = arrival rank: stored points first (they arrived earlier)
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    const size_t T = std::max<size_t>(1, std::min<size_t>(hw, (total + (1u << 16) - 1) >> 16));

    // 1. keys.  Stored points keep their voxel; new points: static_cast<int>(p / voxel_size) — truncation toward zero (vhm.cpp:275).
    std::vector<uint64_t> keys(total);
    std::atomic<bool> bad{false};
    const double vs = voxel_size;
    parallel_for(V(), 1 << 12, [&](size_t vb, size_t ve) {
        for (size_t v = vb; v < ve; ++v) for (uint32_t p = vstart[v]; p < vstart[v + 1]; ++p) keys[p] = vkey[v];
    });
    parallel_for(n, 1 << 16, [&](size_t b0, size_t e0) {
        for (size_t i = b0; i < e0; ++i) {
            const double qx = static_cast<double>(xyz[3 * i]) / vs, qy = static_cast<double>(xyz[3 * i + 1]) / vs,
                         qz = static_cast<double>(xyz[3 * i + 2]) / vs;
            // (two voxels of margin so that every centre key whose neighbourhood reaches a stored voxel is itself packable)
            const double lim = static_cast<double>(kKeyBias - 2);
            if (!(std::fabs(qx) < lim && std::fabs(qy) < lim && std::fabs(qz) < lim)) { bad = true; keys[P0 + i] = 0; continue; }
            keys[P0 + i] = pack_key(static_cast<int32_t>(qx), static_cast<int32_t>(qy), static_cast<int32_t>(qz));
        }
    });
    if (bad) return "map point outside +-2^20 voxels per axis (or not finite)";
    timer.lap("keys");

    // 2. slabs: bucket = (biased x field - min) >> shift, at most 2^14 buckets
    uint64_t xmin = ~0ull, xmax = 0;
    {
        std::vector<uint64_t> lo(T, ~0ull), hi(T, 0);
        std::vector<std::thread> th;
        for (size_t t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                uint64_t a = ~0ull, c = 0;
                for (size_t i = total * t / T; i < total * (t + 1) / T; ++i) { const uint64_t x = keys[i] >> (2 * kKeyBits); a = std::min(a, x); c = std::max(c, x); }
                lo[t] = a; hi[t] = c;
            });
        for (auto& x : th) x.join();
        for (size_t t = 0; t < T; ++t) { xmin = std::min(xmin, lo[t]); xmax = std::max(xmax, hi[t]); }
    }
    int shift = 0;
    while (((xmax - xmin) >> shift) >= (1u << 14)) ++shift;
    const size_t B = static_cast<size_t>((xmax - xmin) >> shift) + 1;
    auto bucket_of = [&](uint64_t key) { return static_cast<size_t>(((key >> (2 * kKeyBits)) - xmin) >> shift); };
    // per-thread histograms over contiguous arrival chunks; offsets bucket-major then thread-major keep arrival order inside a bucket
    std::vector<uint32_t> hist(T * B, 0);
    {
        std::vector<std::thread> th;
        for (size_t t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                uint32_t* h = &hist[t * B];
                for (size_t i = total * t / T; i < total * (t + 1) / T; ++i) ++h[bucket_of(keys[i])];
            });
        for (auto& x : th) x.join();
    }
    std::vector<uint32_t> bstart(B + 1, 0);
    {
        uint32_t run = 0;
        for (size_t b0 = 0; b0 < B; ++b0) {
            bstart[b0] = run;
            for (size_t t = 0; t < T; ++t) { const uint32_t c = hist[t * B + b0]; hist[t * B + b0] = run; run += c; }
        }
        bstart[B] = run;
    }
    std::vector<Rec> recs(total);
    {
        std::vector<std::thread> th;
        for (size_t t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                uint32_t* h = &hist[t * B];
                for (size_t i = total * t / T; i < total * (t + 1) / T; ++i) {
                    const float* c = i < P0 ? &pxyz[3 * i] : &xyz[3 * (i - P0)];
                    recs[h[bucket_of(keys[i])]++] = Rec{keys[i], static_cast<uint32_t>(i), c[0], c[1], c[2]};
                }
            });
        for (auto& x : th) x.join();
    }
    std::vector<uint64_t>().swap(keys);
    timer.lap("partition into x-slabs");

    // 3. per slab: sort by (voxel, arrival), sequential spacing filter per voxel (AddPointWithSpacing, vhm.hpp:106-113), kept
    //    points compacted in place at the front of the slab; voxel keys / counts collected per slab
    const double map_resolution = std::sqrt(voxel_size * voxel_size / cap);
    // sqrt(d2) < map_resolution  <=>  d2 < d2_limit with d2_limit = the smallest double whose correctly rounded square root
    // is >= map_resolution (sqrt is monotone): the same decisions as the reference's test, without a square root per pair
    double d2_limit = map_resolution * map_resolution;
    while (std::sqrt(d2_limit) >= map_resolution) d2_limit = std::nextafter(d2_limit, 0.0);
    while (std::sqrt(d2_limit) < map_resolution) d2_limit = std::nextafter(d2_limit, HUGE_VAL);
    std::vector<std::vector<uint64_t>> slab_keys(B);
    std::vector<std::vector<uint32_t>> slab_counts(B);
    std::atomic<size_t> next{0};
    {
        std::vector<std::thread> th;
        for (size_t t = 0; t < T; ++t)
            th.emplace_back([&]() {
                std::vector<double> kept;
                std::vector<Rec> tmp;
                std::vector<uint32_t> cnt;
                for (;;) {
                    const size_t b0 = next.fetch_add(1);
                    if (b0 >= B) break;
                    Rec* r = recs.data() + bstart[b0];
                    const size_t m = bstart[b0 + 1] - bstart[b0];
                    // stable LSD counting sort by key, one pass per 21-bit coordinate field (z, then y, then x) with as many bins as
                    // the field's range inside the slab (fields wider than 2^12 values take two passes; constant fields none).
                    // The slab is in arrival order and every pass is stable, so the result is ordered by (voxel, arrival).
                    if (tmp.size() < m) tmp.resize(m);
                    {
                        Rec* src = r;
                        Rec* dst = tmp.data();
                        const uint64_t fmask = (1ull << kKeyBits) - 1;
                        for (int f = 0; f < 3 && m > 1; ++f) {
                            const int fsh = kKeyBits * f;  // 0: z, 21: y, 42: x
                            uint64_t lo = ~0ull, hi = 0;
                            for (size_t i = 0; i < m; ++i) { const uint64_t v = (src[i].key >> fsh) & fmask; lo = std::min(lo, v); hi = std::max(hi, v); }
                            const uint64_t range = hi - lo + 1;
                            const int passes = range == 1 ? 0 : (range <= 4096 ? 1 : 2);
                            for (int ps = 0; ps < passes; ++ps) {
                                const int dsh = (passes == 2 && ps == 1) ? 11 : 0;
                                const uint64_t dmask = passes == 2 ? 2047u : ~0ull;
                                const size_t bins = passes == 2 ? 2048 : static_cast<size_t>(range);
                                cnt.assign(bins + 1, 0);
                                for (size_t i = 0; i < m; ++i) ++cnt[static_cast<size_t>(((((src[i].key >> fsh) & fmask) - lo) >> dsh) & dmask) + 1];
                                for (size_t k = 0; k < bins; ++k) cnt[k + 1] += cnt[k];
                                for (size_t i = 0; i < m; ++i) dst[cnt[static_cast<size_t>(((((src[i].key >> fsh) & fmask) - lo) >> dsh) & dmask)]++] = src[i];
                                std::swap(src, dst);
                            }
                        }
                        if (src != r) std::memcpy(r, src, m * sizeof(Rec));
                    }
                    size_t out = 0;
                    for (size_t i = 0; i < m;) {
                        const uint64_t key = r[i].key;
                        kept.clear();
                        size_t j = i;
                        for (; j < m && r[j].key == key; ++j) {
                            const double x = r[j].x, y = r[j].y, z = r[j].z;
                            bool ok = true;
                            if (r[j].idx >= P0 && !kept.empty()) {      // first point of a voxel is always kept (vhm.cpp:281-283)
                                if (kept.size() / 3 >= static_cast<size_t>(cap)) ok = false;  // full: nothing later can enter
                                for (size_t k = 0; ok && k < kept.size(); k += 3) {
                                    const double dx = kept[k] - x, dy = kept[k + 1] - y, dz = kept[k + 2] - z;
                                    if ((dx * dx + dy * dy) + dz * dz < d2_limit) ok = false;
                                }
                            }
                            if (ok) { kept.push_back(x); kept.push_back(y); kept.push_back(z); r[out++] = r[j]; }
                        }
                        slab_keys[b0].push_back(key);
                        slab_counts[b0].push_back(static_cast<uint32_t>(kept.size() / 3));
                        i = j;
                    }
                }
            });
        for (auto& x : th) x.join();
    }
    timer.lap("sort + spacing filter per slab");

    // 4. concatenate the slabs (slab order == key order) into the canonical arrays
    std::vector<size_t> vbase(B + 1, 0), pbase(B + 1, 0);
    for (size_t b0 = 0; b0 < B; ++b0) {
        vbase[b0 + 1] = vbase[b0] + slab_keys[b0].size();
        size_t c = 0;
        for (uint32_t k : slab_counts[b0]) c += k;
        pbase[b0 + 1] = pbase[b0] + c;
    }
    const size_t G = vbase[B], P1 = pbase[B];
    std::vector<uint64_t> nkey(G);
    std::vector<uint32_t> nstart(G + 1, 0);
    std::vector<float> nxyz(3 * P1);
    std::vector<uint32_t> norig(P1);
    parallel_for(B, 1, [&](size_t bb, size_t be) {
        for (size_t b0 = bb; b0 < be; ++b0) {
            size_t p = pbase[b0];
            for (size_t g = 0; g < slab_keys[b0].size(); ++g) {
                nkey[vbase[b0] + g] = slab_keys[b0][g];
                nstart[vbase[b0] + g] = static_cast<uint32_t>(p);
                p += slab_counts[b0][g];
            }
            const Rec* r = recs.data() + bstart[b0];
            for (size_t i = 0, o = pbase[b0]; o < pbase[b0 + 1]; ++i, ++o) {
                nxyz[3 * o] = r[i].x; nxyz[3 * o + 1] = r[i].y; nxyz[3 * o + 2] = r[i].z;
                norig[o] = r[i].idx < P0 ? porig[r[i].idx] : static_cast<uint32_t>(n_raw_seen + (r[i].idx - P0));
            }
        }
    });
    nstart[G] = static_cast<uint32_t>(P1);
    vkey.swap(nkey); vstart.swap(nstart); pxyz.swap(nxyz); porig.swap(norig);
    n_raw_seen += n;
    has_vcov = has_pcov = false;
    vmean.clear(); vcov.clear(); pmean.clear(); pcov.clear(); pnormal.clear();
    vcand8.clear(); dir7.clear();
    timer.lap("compaction");
    build_table();
    timer.lap("voxel table");
    const std::string e = build_directory();
    timer.lap("neighbourhood directory");
    return e;
}

std::string HostMap::adopt_canonical(std::vector<uint64_t>& nkey, std::vector<uint32_t>& nstart, std::vector<float>& nxyz, std::vector<uint32_t>& norig,
                                     size_t n_raw) {
    if (!vkey.empty()) return "adopt_canonical: the map is not empty";
    if (cap > static_cast<int>(kDirCountMask)) return "max_points_per_voxel above 1023 does not fit the column descriptors";
    StageTimer timer;
    vkey.swap(nkey); vstart.swap(nstart); pxyz.swap(nxyz); porig.swap(norig);
    n_raw_seen += n_raw;
    has_vcov = has_pcov = false;
    vmean.clear(); vcov.clear(); pmean.clear(); pcov.clear(); pnormal.clear();
    vcand8.clear(); dir7.clear();
    build_table();
    timer.lap("voxel table");
    const std::string e = build_directory();
    timer.lap("neighbourhood directory");
    return e;
}

void HostMap::adopt_voxel_cov(std::vector<double>& mean, std::vector<double>& cov) {
    vmean.swap(mean); vcov.swap(cov);
    has_vcov = true;
    build_voxel_candidates();
}

void HostMap::adopt_point_cov(std::vector<double>& mean, std::vector<double>& cov, std::vector<double>& normal) {
    pmean.swap(mean); pcov.swap(cov); pnormal.swap(normal);
    has_pcov = true;
}

// ---- neighbourhood directory ------------------------------------------------------------------------------------
namespace {

// keys (sorted, unique) dilated by one voxel along one axis: {k - unit, k, k + unit}, still sorted and unique
std::vector<uint64_t> dilate_axis(const std::vector<uint64_t>& in, int axis) {
    const int shift = kKeyBits * (2 - axis);
    const uint64_t unit = 1ull << shift, fmask = (1ull << kKeyBits) - 1;
    std::vector<uint64_t> lo, hi;
    lo.reserve(in.size()); hi.reserve(in.size());
    for (uint64_t k : in) {
        const uint64_t f = (k >> shift) & fmask;
        if (f > 0) lo.push_back(k - unit);
        if (f < fmask) hi.push_back(k + unit);
    }
    std::vector<uint64_t> t(lo.size() + in.size()), out(lo.size() + in.size() + hi.size());
    std::merge(lo.begin(), lo.end(), in.begin(), in.end(), t.begin());
    std::merge(t.begin(), t.end(), hi.begin(), hi.end(), out.begin());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
}

}  // namespace

// Octant order of the stored points for the device (voxel_key.hpp): per voxel a stable counting sort by octant code.
void HostMap::build_octants() {
    const size_t nv = V(), np = P();
    octants = cap <= kOctantCapMax;
    dev_order.resize(np);
    voct.assign(8 * nv, 0);
    const double vs = voxel_size;
    parallel_for(nv, 1 << 12, [&](size_t vb, size_t ve) {
        uint8_t code[kOctantCapMax + 1];
        for (size_t v = vb; v < ve; ++v) {
            const uint32_t s = vstart[v], n = vstart[v + 1] - s;
            if (!octants) {  // whole voxels only: identity order, no octant words
                for (uint32_t i = 0; i < n; ++i) dev_order[s + i] = s + i;
                continue;
            }
            int32_t cx, cy, cz;
            unpack_key(vkey[v], cx, cy, cz);
            uint32_t cnt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (uint32_t i = 0; i < n; ++i) {
                const float* p = &pxyz[3 * static_cast<size_t>(s + i)];
                const int o = (axis_half(static_cast<double>(p[2]) / vs, cz) << 2) | (axis_half(static_cast<double>(p[1]) / vs, cy) << 1) |
                              axis_half(static_cast<double>(p[0]) / vs, cx);
                code[i] = static_cast<uint8_t>(o);
                ++cnt[o + 1];
            }
            for (int k = 0; k < 8; ++k) cnt[k + 1] += cnt[k];  // cnt[k] = points in octants < k
            uint8_t* w = &voct[8 * v];
            for (int k = 1; k <= 7; ++k) w[k - 1] = static_cast<uint8_t>(cnt[k]);
            w[7] = static_cast<uint8_t>(n);
            uint32_t at[8];
            for (int k = 0; k < 8; ++k) at[k] = cnt[k];
            for (uint32_t i = 0; i < n; ++i) dev_order[s + at[code[i]]++] = s + i;
        }
    });
}

std::string HostMap::build_directory() {
    dir_slots.clear(); dir_rows.clear(); dir_bmask = 0; dir_entries = 0;
    dev_order.clear(); voct.clear();
    if (vkey.empty()) return "";
    StageTimer timer;
    build_octants();
    timer.lap("  octant order");
    // 1. centre keys: occupied voxels and their one-voxel halo (separable dilation z, y, x of the sorted key list)
    std::vector<uint64_t> E = dilate_axis(dilate_axis(dilate_axis(vkey, 2), 1), 0);
    dir_entries = E.size();
    if (E.size() >= (1ull << 31)) return "map too large for the neighbourhood directory";
    timer.lap("  dilation");
    // 2. 2-choice cuckoo placement, 2 slots per bucket, load <= 0.75; deterministic random walk, table doubled on failure
    size_t nb = 2;
    while (nb * 2 * 3 < E.size() * 4) nb <<= 1;
    std::vector<int64_t> owner;  // slot -> entry index or -1
    for (;; nb <<= 1) {
        if (nb * 2 >= (1ull << 32)) return "map too large for the neighbourhood directory";
        const uint32_t bm = static_cast<uint32_t>(nb - 1);
        owner.assign(nb * 2, -1);
        uint64_t rng = 0x9E3779B97F4A7C15ull;
        bool ok = true;
        for (size_t e0 = 0; e0 < E.size() && ok; ++e0) {
            int64_t cur = static_cast<int64_t>(e0);
            uint32_t avoid = 0xffffffffu;  // bucket the current key was just evicted from
            for (int kick = 0;; ++kick) {
                uint32_t b1, b2;
                dir_buckets(E[cur], bm, b1, b2);
                int64_t* s1 = &owner[2 * static_cast<size_t>(b1)];
                int64_t* s2 = &owner[2 * static_cast<size_t>(b2)];
                if (s1[0] < 0) { s1[0] = cur; break; }
                if (s1[1] < 0) { s1[1] = cur; break; }
                if (s2[0] < 0) { s2[0] = cur; break; }
                if (s2[1] < 0) { s2[1] = cur; break; }
                if (kick >= 2000) { ok = false; break; }
                rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                const uint32_t b = (b1 == avoid) ? b2 : ((b2 == avoid) ? b1 : ((rng & 2) ? b1 : b2));
                int64_t* victim = &owner[2 * static_cast<size_t>(b) + (rng & 1)];
                std::swap(cur, *victim);
                avoid = b;
            }
        }
        if (ok) { dir_bmask = bm; break; }
    }
    timer.lap("  cuckoo placement");
    // 3. slots + rows (row r belongs to slot r).  Entries are walked in KEY order with one forward cursor per column into the
    //    sorted voxel list (shifting a key by (dx, dy) keeps the order), so the nine runs of an entry cost O(1) amortised
    //    instead of a binary search each; the rows are written to the entry's slot.
    const size_t S = owner.size();
    dir_slots.assign(S, DirSlot{0xffffffffu, 0xffffffffu, 0, 0});
    dir_rows.assign(S * kRowWords, 0u);
    std::vector<uint32_t> slot_of(E.size());
    for (size_t s = 0; s < S; ++s) if (owner[s] >= 0) slot_of[static_cast<size_t>(owner[s])] = static_cast<uint32_t>(s);
    const size_t nV = vkey.size();
    parallel_for(E.size(), 1 << 14, [&](size_t eb, size_t ee) {
        size_t cur[9];
        uint64_t last[9];
        bool primed[9] = {false, false, false, false, false, false, false, false, false};
        for (size_t e = eb; e < ee; ++e) {
            const uint64_t key = E[e];
            int32_t x, y, z;
            unpack_key(key, x, y, z);
            const size_t s = slot_of[e];
            uint32_t* hdr = &dir_rows[row_word(s, 0)];
            hdr[kRowKeyLo] = static_cast<uint32_t>(key); hdr[kRowKeyHi] = static_cast<uint32_t>(key >> 32);
            hdr[kRowFlags] = octants ? kRowFlagOctants : 0u;
            for (int c = 0; c < 9; ++c) {
                const int32_t cx = x + c / 3 - 1, cy = y + c % 3 - 1;
                if (!key_in_range(cx) || !key_in_range(cy)) continue;
                const int32_t zlo = std::max(z - 1, -kKeyBias), zhi = std::min(z + 1, kKeyBias - 1);
                const uint64_t klo = pack_key(cx, cy, zlo), khi = pack_key(cx, cy, zhi);
                if (!primed[c] || klo < last[c]) {  // first use in this chunk (or a non-monotone step at the range border)
                    cur[c] = static_cast<size_t>(std::lower_bound(vkey.begin(), vkey.end(), klo) - vkey.begin());
                    primed[c] = true;
                } else {
                    size_t v = cur[c], step = 0;
                    while (v < nV && vkey[v] < klo && step < 64) { ++v; ++step; }
                    if (v < nV && vkey[v] < klo) v = static_cast<size_t>(std::lower_bound(vkey.begin() + v, vkey.end(), klo) - vkey.begin());
                    cur[c] = v;
                }
                last[c] = klo;
                uint32_t first = 0, counts = 0;
                bool any = false;
                uint32_t* col = &dir_rows[row_col_word(s, c)];
                for (size_t v = cur[c]; v < nV && vkey[v] <= khi; ++v) {
                    int32_t vx, vy, vz;
                    unpack_key(vkey[v], vx, vy, vz);
                    if (!any) { first = vstart[v]; any = true; }
                    const uint32_t dz = static_cast<uint32_t>(vz - (z - 1));
                    counts |= (vstart[v + 1] - vstart[v]) << (kDirCountBits * dz);
                    std::memcpy(&col[2 + 2 * dz], &voct[8 * v], 8);
                }
                col[0] = first; col[1] = counts;
            }
            const DirDesc centre = row_column(s, 4);
            dir_slots[s] = DirSlot{static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), centre.first, centre.counts};
        }
    });
    timer.lap("  slots + rows");
    return "";
}

// ---- built-map file ------------------------------------------------------------------------------------------------
namespace {

constexpr char kMapMagic[8] = {'E', 'L', 'M', 'B', '2', '0', '0', 'M'};
constexpr uint32_t kMapVersion = 4;

// checksum of everything written / read so far: 8 interleaved FNV-1a lanes over 64-bit words (tail bytes one by one)
struct Checksum {
    uint64_t h[8] = {0xcbf29ce484222325ull, 0x84222325cbf29ce4ull, 0x9ce484222325cbf2ull, 0x2325cbf29ce48422ull,
                     0xf29ce484222325cbull, 0x222325cbf29ce484ull, 0xe484222325cbf29cull, 0x25cbf29ce4842223ull};
    void add(const void* p, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        size_t i = 0;
        for (; i + 64 <= n; i += 64)
            for (int l = 0; l < 8; ++l) { uint64_t w; std::memcpy(&w, b + i + 8 * l, 8); h[l] = (h[l] ^ w) * 0x100000001b3ull; }
        for (; i < n; ++i) h[i & 7] = (h[i & 7] ^ b[i]) * 0x100000001b3ull;
    }
    uint64_t value() const { uint64_t v = 0; for (int l = 0; l < 8; ++l) v = (v ^ h[l]) * 0x100000001b3ull; return v; }
};
struct Writer {
    std::FILE* f; Checksum c; bool ok = true;
    void put(const void* p, size_t n) { if (ok && n) { ok = std::fwrite(p, 1, n, f) == n; c.add(p, n); } }
};
struct Reader {
    std::FILE* f; Checksum c; bool ok = true;
    void get(void* p, size_t n) { if (ok && n) { ok = std::fread(p, 1, n, f) == n; if (ok) c.add(p, n); } }
};

struct FileCloser { void operator()(std::FILE* f) const { if (f) std::fclose(f); } };

template <class T>
void write_vec(Writer& w, const std::vector<T>& v) {
    const uint64_t n = v.size();
    w.put(&n, sizeof n);
    w.put(v.data(), n * sizeof(T));
}
template <class T>
void read_vec(Reader& r, std::vector<T>& v, uint64_t max_elems) {
    uint64_t n = 0;
    r.get(&n, sizeof n);
    if (!r.ok || n > max_elems) { r.ok = false; return; }
    v.resize(n);
    r.get(v.data(), n * sizeof(T));
}

}  // namespace

// The file holds the CANONICAL arrays only (sorted voxel keys, point prefix, stored points, insertion indices, covariances).
// Everything the kernels dereference — voxel table, neighbourhood directory, octant order, candidate lists — is rebuilt from
// them on load, so an internally inconsistent (or crafted) file cannot make the device read out of bounds.
std::string HostMap::save(const std::string& path) const {
    std::unique_ptr<std::FILE, FileCloser> f(std::fopen(path.c_str(), "wb"));
    if (!f) return "cannot open " + path + " for writing";
    const uint32_t flags = (has_vcov ? 1u : 0u) | (has_pcov ? 2u : 0u);
    const uint64_t counts[4] = {V(), P(), dir_entries, n_raw_seen};
    Writer w{f.get()};
    w.put(kMapMagic, 8); w.put(&kMapVersion, 4); w.put(&flags, 4); w.put(&voxel_size, 8); w.put(&cap, 4); w.put(&mask, 4);
    w.put(&dir_bmask, 4); w.put(counts, sizeof counts);
    write_vec(w, vkey); write_vec(w, vstart); write_vec(w, pxyz); write_vec(w, porig); write_vec(w, vmean); write_vec(w, vcov);
    write_vec(w, pmean); write_vec(w, pcov); write_vec(w, pnormal);
    const uint64_t sum = w.c.value();
    bool ok = w.ok && std::fwrite(&sum, 8, 1, f.get()) == 1;
    if (!ok || std::fflush(f.get()) != 0) return "short write to " + path;
    return "";
}

std::string HostMap::load(const std::string& path) {
    std::unique_ptr<std::FILE, FileCloser> f(std::fopen(path.c_str(), "rb"));
    if (!f) return "cannot open " + path;
    char magic[8];
    uint32_t version = 0, flags = 0, file_mask = 0, file_bmask = 0;
    uint64_t counts[4] = {0, 0, 0, 0};
    HostMap m;
    Reader r{f.get()};
    r.get(magic, 8); r.get(&version, 4);
    if (!r.ok || std::memcmp(magic, kMapMagic, 8) != 0) return path + " is not an elimaloc_b200 map file";
    if (version != kMapVersion) return path + ": unsupported map file version " + std::to_string(version);
    r.get(&flags, 4); r.get(&m.voxel_size, 8); r.get(&m.cap, 4); r.get(&file_mask, 4); r.get(&file_bmask, 4); r.get(counts, sizeof counts);
    const uint64_t lim = 1ull << 34;
    read_vec(r, m.vkey, lim); read_vec(r, m.vstart, lim); read_vec(r, m.pxyz, lim); read_vec(r, m.porig, lim); read_vec(r, m.vmean, lim);
    read_vec(r, m.vcov, lim); read_vec(r, m.pmean, lim); read_vec(r, m.pcov, lim); read_vec(r, m.pnormal, lim);
    uint64_t sum = 0;
    bool ok = r.ok && std::fread(&sum, 8, 1, f.get()) == 1 && sum == r.c.value();
    if (!ok) return path + ": truncated or corrupt map file";
    m.has_vcov = (flags & 1u) != 0;
    m.has_pcov = (flags & 2u) != 0;
    m.n_raw_seen = counts[3];
    // structural consistency of the canonical arrays (a file from another build / a damaged file must not reach the kernels)
    const size_t nv = m.vkey.size(), np = m.pxyz.size() / 3;
    ok = counts[0] == nv && counts[1] == np && m.pxyz.size() == 3 * np && m.vstart.size() == nv + 1 && m.porig.size() == np &&
         np < (1ull << 32) && (nv == 0 || (m.vstart.front() == 0 && m.vstart.back() == np)) && (nv != 0 || np == 0) &&
         m.voxel_size > 0.0 && m.cap >= 1 && m.cap <= static_cast<int>(kDirCountMask) &&
         (!m.has_vcov || (m.vmean.size() == 3 * nv && m.vcov.size() == 9 * nv)) &&
         (!m.has_pcov || (m.pmean.size() == 3 * np && m.pcov.size() == 9 * np && m.pnormal.size() == 3 * np));
    for (size_t v = 0; ok && v < nv; ++v) {
        int32_t x, y, z;
        unpack_key(m.vkey[v], x, y, z);
        const int32_t lim_k = kKeyBias - 2;
        ok = (m.vkey[v] >> (3 * kKeyBits)) == 0 && std::abs(x) < lim_k && std::abs(y) < lim_k && std::abs(z) < lim_k &&
             m.vstart[v] < m.vstart[v + 1] && m.vstart[v + 1] - m.vstart[v] <= static_cast<uint32_t>(m.cap) &&
             (v + 1 == nv || m.vkey[v] < m.vkey[v + 1]);
    }
    if (!ok) return path + ": inconsistent map file";
    m.build_table();
    const std::string e = m.build_directory();
    if (!e.empty()) return path + ": " + e;
    if (m.has_vcov) m.build_voxel_candidates();
    if (m.dir_entries != counts[2] || m.mask != file_mask || m.dir_bmask != file_bmask) return path + ": inconsistent map file";
    *this = std::move(m);
    return "";
}

int64_t HostMap::dir_find(uint64_t key) const {
    if (dir_slots.empty()) return -1;
    uint32_t b1, b2;
    dir_buckets(key, dir_bmask, b1, b2);
    const uint32_t lo = static_cast<uint32_t>(key), hi = static_cast<uint32_t>(key >> 32);
    for (uint32_t b : {b1, b2})
        for (uint32_t j = 0; j < 2; ++j) {
            const DirSlot& s = dir_slots[2 * static_cast<size_t>(b) + j];
            if (s.key_lo == lo && s.key_hi == hi) return static_cast<int64_t>(2 * static_cast<size_t>(b) + j);
        }
    return -1;
}

void HostMap::build_table() {
    size_t capacity = 16;
    while (capacity < 2 * V()) capacity <<= 1;
    mask = static_cast<uint32_t>(capacity - 1);
    slots.assign(capacity, Slot{0xffffffffu, 0xffffffffu, 0, 0});
    slot_voxel.assign(capacity, -1);
    for (size_t v = 0; v < V(); ++v) {
        uint32_t h = home_slot(vkey[v]) & mask;
        while (slot_voxel[h] >= 0) h = (h + 1) & mask;
        slots[h] = Slot{static_cast<uint32_t>(vkey[v]), static_cast<uint32_t>(vkey[v] >> 32), vstart[v], vstart[v + 1] - vstart[v]};
        slot_voxel[h] = static_cast<int32_t>(v);
    }
}

int64_t HostMap::find(uint64_t key) const {
    if (slots.empty()) return -1;
    uint32_t h = home_slot(key) & mask;
    for (;;) {
        const int32_t v = slot_voxel[h];
        if (v < 0) return -1;
        if (vkey[v] == key) return v;
        h = (h + 1) & mask;
    }
}

bool HostMap::find_ground_height(double x, double y, double& ground_z) const {
    const double range = 5.0, range2 = range * range;                     // vhm.hpp:286-287
    if (vkey.empty()) return false;
    // stored keys truncate toward zero: a point with p / vs in (c - 1, c + 1) may sit under key c, so widen by one voxel
    const double vs = voxel_size;
    const int32_t x0 = static_cast<int32_t>(std::max(std::floor((x - range) / vs) - 1.0, static_cast<double>(-kKeyBias))),
                  x1 = static_cast<int32_t>(std::min(std::floor((x + range) / vs) + 1.0, static_cast<double>(kKeyBias - 1)));
    const int32_t y0 = static_cast<int32_t>(std::max(std::floor((y - range) / vs) - 1.0, static_cast<double>(-kKeyBias))),
                  y1 = static_cast<int32_t>(std::min(std::floor((y + range) / vs) + 1.0, static_cast<double>(kKeyBias - 1)));
    std::vector<double> zs;
    for (int32_t kx = x0; kx <= x1; ++kx)
        for (int32_t ky = y0; ky <= y1; ++ky) {
            // all voxels of the column (kx, ky, *) are one contiguous run of the sorted key list
            const uint64_t lo = pack_key(kx, ky, -kKeyBias), hi = pack_key(kx, ky, kKeyBias - 1);
            for (size_t v = static_cast<size_t>(std::lower_bound(vkey.begin(), vkey.end(), lo) - vkey.begin()); v < vkey.size() && vkey[v] <= hi; ++v)
                for (uint32_t p = vstart[v]; p < vstart[v + 1]; ++p) {
                    const double dx = static_cast<double>(pxyz[3 * p]) - x, dy = static_cast<double>(pxyz[3 * p + 1]) - y;
                    if (dx * dx + dy * dy <= range2) zs.push_back(pxyz[3 * p + 2]);   // vhm.hpp:292-296
                }
        }
    if (zs.size() <= 3) return false;                                       // vhm.hpp:298-300
    const size_t n = std::min<size_t>(5, zs.size());                        // vhm.hpp:302-306
    std::partial_sort(zs.begin(), zs.begin() + static_cast<std::ptrdiff_t>(n), zs.end());
    double sum = 0.0;
    for (size_t i = 0; i < n; ++i) sum += zs[i];
    ground_z = sum / static_cast<double>(n);                                // vhm.hpp:318-319
    return true;
}

// CalVoxelCovAll: n == 0 -> (I, 0) [cannot occur: a voxel exists only with >= 1 point]; n == 1 -> (I, p);
// n >= 2 -> sample covariance /(n-1), regularised, with the mean.
void HostMap::cal_voxel_cov() {
    const size_t nv = V();
    vmean.assign(3 * nv, 0.0);
    vcov.assign(9 * nv, 0.0);
    parallel_for(nv, 4096, [&](size_t b, size_t e) {
        std::vector<double> buf;
        for (size_t v = b; v < e; ++v) {
            const uint32_t s = vstart[v], cnt = vstart[v + 1] - s;
            double* m = &vmean[3 * v];
            double* c = &vcov[9 * v];
            if (cnt == 1) {
                m[0] = pxyz[3 * s]; m[1] = pxyz[3 * s + 1]; m[2] = pxyz[3 * s + 2];
                c[0] = c[4] = c[8] = 1.0;
                continue;
            }
            buf.resize(3 * cnt);
            for (uint32_t i = 0; i < 3 * cnt; ++i) buf[i] = pxyz[3 * s + i];
            mean_cov_regularized(buf.data(), cnt, m, c, nullptr);
        }
    });
    has_vcov = true;
    build_voxel_candidates();
}

void HostMap::build_voxel_candidates() {
    const size_t S = dir_slots.size();
    vcand8.clear();
    if (S == 0) return;
    std::vector<int32_t> voxel_slot(V(), -1);
    for (size_t s = 0; s < slot_voxel.size(); ++s) if (slot_voxel[s] >= 0) voxel_slot[slot_voxel[s]] = static_cast<int32_t>(s);
    // The nine column descriptors of the entry's row already say which of the 27 voxels exist (count > 0) and where their
    // points start; the voxels of one column are consecutive in the sorted voxel list, so one search per non-empty column
    // (first point -> voxel index) replaces 27 table look-ups.
    // pass 1: occupancy mask + count per entry; pass 2: fill (two passes so that the candidate array is laid out in slot order)
    std::vector<uint32_t> first(S + 1, 0);
    std::vector<uint32_t> masks(S, 0);
    parallel_for(S, 1 << 14, [&](size_t sb, size_t se) {
        for (size_t s = sb; s < se; ++s) {
            const DirSlot& sl = dir_slots[s];
            if ((sl.key_lo & sl.key_hi) == 0xffffffffu) continue;
            uint32_t mask = 0;
            for (int c = 0; c < 9; ++c) {
                const uint32_t counts = row_column(s, c).counts;
                for (int dz = 0; dz < 3; ++dz) if ((counts >> (kDirCountBits * dz)) & kDirCountMask) mask |= 1u << (3 * c + dz);
            }
            masks[s] = mask;
        }
    });
    for (size_t s = 0; s < S; ++s) first[s + 1] = first[s] + static_cast<uint32_t>(__builtin_popcount(masks[s]));
    vcand8.assign(static_cast<size_t>(first[S]) + 4, 0ull);  // (+4 records of padding: aligned 32-byte loads)
    dir7.assign(8 * S, -1);
    parallel_for(S, 1 << 13, [&](size_t sb, size_t se) {
        int64_t v27[27];
        for (size_t s = sb; s < se; ++s) {
            uint32_t* hdr = &dir_rows[row_word(s, 0)];
            hdr[kRowCandFirst] = first[s]; hdr[kRowCandCount] = first[s + 1] - first[s]; hdr[kRowOccMask] = masks[s];
            if (!masks[s]) continue;
            for (int c = 0; c < 9; ++c) {
                v27[3 * c] = v27[3 * c + 1] = v27[3 * c + 2] = -1;
                if (!((masks[s] >> (3 * c)) & 7u)) continue;
                // voxel that owns the column's first point: vstart[v] <= first < vstart[v + 1]
                int64_t v = static_cast<int64_t>(std::upper_bound(vstart.begin(), vstart.end(), row_column(s, c).first) - vstart.begin()) - 1;
                for (int dz = 0; dz < 3; ++dz) if ((masks[s] >> (3 * c + dz)) & 1u) v27[3 * c + dz] = v++;
            }
            static const int kL7[7] = {13, 22, 4, 16, 10, 14, 12};  // centre, +x, -x, +y, -y, +z, -z
            for (int j = 0; j < 7; ++j) if (v27[kL7[j]] >= 0) dir7[8 * s + j] = static_cast<int32_t>(v27[kL7[j]]);
            int32_t key[3];
            unpack_key((static_cast<uint64_t>(dir_slots[s].key_hi) << 32) | dir_slots[s].key_lo, key[0], key[1], key[2]);
            uint64_t* out = &vcand8[first[s]];
            for (int L = 0; L < 27; ++L) {
                if (v27[L] < 0) continue;
                const size_t v = static_cast<size_t>(v27[L]);
                *out++ = pack_vcand(vmean[3 * v] / voxel_size - key[0], vmean[3 * v + 1] / voxel_size - key[1], vmean[3 * v + 2] / voxel_size - key[2],
                                    static_cast<uint32_t>(v));
            }
        }
    });
}

// CalPointCovAll: per stored point, neighbours = {self} U {stored points of the 27 voxels around FLOOR(p / vs)
// with d^2 <= r^2}; self is in that set too, so it is counted twice and n >= 2 always.
void HostMap::cal_point_cov(double search_dist) {
    const size_t np = P();
    const double r2 = search_dist * search_dist;
    pmean.assign(3 * np, 0.0);
    pcov.assign(9 * np, 0.0);
    pnormal.assign(3 * np, 0.0);
    const double vs = voxel_size;
    parallel_for(np, 8192, [&](size_t b, size_t e) {
        std::vector<double> nb;
        for (size_t p = b; p < e; ++p) {
            const double x = pxyz[3 * p], y = pxyz[3 * p + 1], z = pxyz[3 * p + 2];
            nb.clear();
            nb.push_back(x); nb.push_back(y); nb.push_back(z);
            const int32_t kx = static_cast<int32_t>(std::floor(x / vs)), ky = static_cast<int32_t>(std::floor(y / vs)),
                          kz = static_cast<int32_t>(std::floor(z / vs));
            // Squared gap (metres, under-estimated) between the point and anything STORED under key c along one axis: insert
            // keys truncate toward zero, so key c holds p / vs in [c, c+1) for c > 0, (c-1, c] for c < 0 and (-1, 1) for c == 0.
            // A voxel whose box is farther than the search radius cannot contribute a neighbour: skipping it leaves the
            // neighbour list — and its order, hence every rounding of the sums — exactly as the full 27-voxel visit builds it.
            auto gap2 = [vs](double p, int32_t c) {
                const double q = p / vs;
                const double lo = static_cast<double>(c <= 0 ? c - 1 : c), hi = static_cast<double>(c >= 0 ? c + 1 : c);
                const double g = std::max(std::max(lo - q, q - hi), 0.0) * vs * (1.0 - 1e-9);
                return g * g;
            };
            const double r2_skip = r2 * (1.0 + 1e-9);
            for (int i = kx - 1; i <= kx + 1; ++i)
                for (int j = ky - 1; j <= ky + 1; ++j)
                    for (int k = kz - 1; k <= kz + 1; ++k) {
                        if (!key_in_range(i) || !key_in_range(j) || !key_in_range(k)) continue;
                        if (gap2(x, i) + gap2(y, j) + gap2(z, k) > r2_skip) continue;
                        const int64_t v = find(pack_key(i, j, k));
                        if (v < 0) continue;
                        for (uint32_t q = vstart[v]; q < vstart[v + 1]; ++q) {
                            const double dx = pxyz[3 * q] - x, dy = pxyz[3 * q + 1] - y, dz = pxyz[3 * q + 2] - z;
                            if ((dx * dx + dy * dy) + dz * dz <= r2) { nb.push_back(pxyz[3 * q]); nb.push_back(pxyz[3 * q + 1]); nb.push_back(pxyz[3 * q + 2]); }
                        }
                    }
            mean_cov_regularized(nb.data(), nb.size() / 3, &pmean[3 * p], &pcov[9 * p], &pnormal[3 * p]);
        }
    });
    has_pcov = true;
}

}  // namespace elm
