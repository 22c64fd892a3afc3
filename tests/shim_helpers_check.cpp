// Exercises the host-side helper members of shim/registration_shim.hpp that the node calls around RunRegister —
// VoxelDownsample and the two TransformPoints overloads — on points read from a file; prints the surviving input indices
// (carried in `intensity`) and the transformed coordinates.  Compiled by tests/test_node_on_shim.py against the stand-in Eigen.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "registration_shim.hpp"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const size_t n = static_cast<size_t>(std::atol(argv[2]));
    const double voxel = std::atof(argv[3]);
    std::vector<float> xyz(3 * n);
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(xyz.data(), sizeof(float), 3 * n, f) != 3 * n) return 3;
    std::fclose(f);
    std::vector<PointStruct> pts(n);
    for (size_t i = 0; i < n; ++i) {
        pts[i].pose = pts[i].local = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        pts[i].intensity = static_cast<double>(i);
    }
    VoxelHashMap map;  // no Init: the helper needs no handle
    const std::vector<PointStruct> ds = map.VoxelDownsample(pts, voxel);
    std::printf("%zu\n", ds.size());
    for (const PointStruct& p : ds) std::printf("%d\n", static_cast<int>(p.intensity));
    Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
    T(0, 0) = 0.0; T(0, 1) = -1.0; T(1, 0) = 1.0; T(1, 1) = 0.0;  // 90 degrees about z
    T(0, 3) = 1.5; T(1, 3) = -2.5; T(2, 3) = 0.25;
    Registration reg;
    std::vector<PointStruct> moved;
    reg.TransformPoints(T, ds, moved);
    std::vector<PointStruct> in_place = ds;
    reg.TransformPoints(T, in_place);
    for (size_t i = 0; i < moved.size() && i < 5; ++i)
        std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %d\n", moved[i].pose.x(), moved[i].pose.y(), moved[i].pose.z(), in_place[i].pose.x(),
                    moved[i].local.x(), ds[i].pose.x(), static_cast<int>(moved[i].intensity));
    return 0;
}
