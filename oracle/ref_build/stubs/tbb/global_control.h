#pragma once
#include "blocked_range.h"
namespace tbb {
struct global_control {
    enum parameter { max_allowed_parallelism };
    global_control(parameter, std::size_t n) { stub_threads() = n > 0 ? static_cast<int>(n) : 1; }
};
}  // namespace tbb
