// PCD v0.7 reader (see pcd_reader.cpp)
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace elm {

// Fills xyz (3 floats per point, points with a non-finite coordinate dropped and counted).  Returns "" or an error message.
std::string read_pcd_xyz(const std::string& path, std::vector<float>& xyz, size_t* dropped);

}  // namespace elm
