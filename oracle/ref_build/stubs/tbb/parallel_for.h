#pragma once
#include "blocked_range.h"
namespace tbb {
template <typename Range, typename Body>
void parallel_for(const Range& range, const Body& body) {
    const std::size_t n = range.empty() ? 0 : range.size();
    const int chunks = detail::chunks_for(n);
    if (chunks <= 1) { body(range); return; }
    detail::run_chunks(n, chunks, [&](int, std::size_t lo, std::size_t hi) { body(Range(range.begin() + lo, range.begin() + hi)); });
}
}  // namespace tbb
