#pragma once
#include <thread>
namespace tbb {
namespace this_task_arena {
inline int max_concurrency() { const unsigned n = std::thread::hardware_concurrency(); return n ? static_cast<int>(n) : 1; }
}  // namespace this_task_arena
}  // namespace tbb
