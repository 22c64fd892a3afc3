// Shadows localization_interface/localization_functions.hpp for the oracle/_ref build.  The real header pulls in ROS, tf,
// OpenCV and PCL (all absent here); the two translation units on the registration path need only three things from it:
// a handful of standard headers, the oneTBB headers, and the names of its ANSI colour strings.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <iostream>
#include <limits>
#include <string>

#include "Eigen/Dense"
#include "localization_struct.hpp"
#include "tbb/blocked_range.h"
#include "tbb/parallel_for.h"
#include "tbb/parallel_for_each.h"
#include "tbb/parallel_reduce.h"

namespace elm_ref_stub {
inline std::string sgr(int code) { return "\x1b[" + std::to_string(code) + "m"; }
}  // namespace elm_ref_stub
static const std::string RESET = elm_ref_stub::sgr(0), RED = elm_ref_stub::sgr(31), GREEN = elm_ref_stub::sgr(32),
                         YELLOW = elm_ref_stub::sgr(33), BLUE = elm_ref_stub::sgr(34), MAGENTA = elm_ref_stub::sgr(35),
                         CYAN = elm_ref_stub::sgr(36), WHITE = elm_ref_stub::sgr(37);
