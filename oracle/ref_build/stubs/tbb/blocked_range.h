// Stand-in for oneTBB (absent from this image).  TBB only schedules the reference's loops: here a range is cut into
// stub_threads() contiguous chunks, one std::thread each, and parallel_reduce joins the partial results left to right — the
// order oneTBB's join guarantees as well — so the vectors come out in scan order whatever the thread count.
#pragma once
#include <algorithm>
#include <cstddef>
#include <thread>
#include <vector>
namespace tbb {
inline int& stub_threads() { static int n = 1; return n; }  // set through ref_set_threads(); 1 = run inline

template <typename Value>
class blocked_range {
public:
    using const_iterator = Value;
    blocked_range(Value b, Value e, std::size_t grain = 1) : b_(b), e_(e) { (void)grain; }
    Value begin() const { return b_; }
    Value end() const { return e_; }
    bool empty() const { return !(b_ < e_); }
    std::size_t size() const { return static_cast<std::size_t>(e_ - b_); }
private:
    Value b_, e_;
};

namespace detail {
// fn(chunk index, first, last) over `chunks` contiguous pieces of [0, n)
template <typename Fn>
void run_chunks(std::size_t n, int chunks, const Fn& fn) {
    std::vector<std::thread> th;
    th.reserve(static_cast<std::size_t>(chunks));
    for (int c = 0; c < chunks; ++c) {
        const std::size_t lo = n * static_cast<std::size_t>(c) / static_cast<std::size_t>(chunks);
        const std::size_t hi = n * static_cast<std::size_t>(c + 1) / static_cast<std::size_t>(chunks);
        th.emplace_back([&fn, c, lo, hi] { fn(c, lo, hi); });
    }
    for (auto& t : th) t.join();
}
inline int chunks_for(std::size_t n) { return static_cast<int>(std::max<std::size_t>(1, std::min<std::size_t>(static_cast<std::size_t>(stub_threads()), n))); }
}  // namespace detail
}  // namespace tbb
