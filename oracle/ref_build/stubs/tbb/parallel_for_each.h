#pragma once
#include "blocked_range.h"
namespace tbb {
template <typename It, typename Body>
void parallel_for_each(It first, It last, const Body& body) {
    if (stub_threads() <= 1) { for (; first != last; ++first) body(*first); return; }
    std::vector<It> items;  // forward iterators (unordered_map): index them first
    for (; first != last; ++first) items.push_back(first);
    const int chunks = detail::chunks_for(items.size());
    detail::run_chunks(items.size(), chunks, [&](int, std::size_t lo, std::size_t hi) { for (std::size_t i = lo; i < hi; ++i) body(*items[i]); });
}
}  // namespace tbb
