#pragma once
#include "blocked_range.h"
namespace tbb {
template <typename Range, typename Value, typename Body, typename Join>
Value parallel_reduce(const Range& range, const Value& identity, const Body& body, const Join& join) {
    const std::size_t n = range.empty() ? 0 : range.size();
    const int chunks = detail::chunks_for(n);
    if (chunks <= 1) return body(range, identity);
    std::vector<Value> part(static_cast<std::size_t>(chunks), identity);
    detail::run_chunks(n, chunks, [&](int c, std::size_t lo, std::size_t hi) {
        part[static_cast<std::size_t>(c)] = body(Range(range.begin() + lo, range.begin() + hi), identity);
    });
    Value acc = std::move(part[0]);
    for (int c = 1; c < chunks; ++c) acc = join(std::move(acc), part[static_cast<std::size_t>(c)]);  // left to right
    return acc;
}
}  // namespace tbb
