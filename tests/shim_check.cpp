// Compiles shim/registration_shim.hpp (against tests/mock_eigen when the image has no Eigen) and drives the reference-side
// call sequence of pcm_matching.cpp:82-101, 280-282 through it.  TEST INFRASTRUCTURE.
//   shim_check <map.f32> <n_map> <scan.f32> <n_scan> <T0 16 doubles, COLUMN-major like Eigen> <method> <out.txt>
// Writes: status line, 16 doubles of the returned pose (column-major storage order, i.e. Eigen's data()), is_success,
// fitness, 36 doubles of local_cov (column-major), number of map points seen through Pointcloud().
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifndef ELM_SHIM_DEVICE
#define ELM_SHIM_DEVICE 0
#endif
#include "registration_shim.hpp"

static std::vector<PointStruct> read_points(const char* path, size_t n) {
    std::vector<float> xyz(3 * n);
    std::FILE* f = std::fopen(path, "rb");
    if (!f || std::fread(xyz.data(), sizeof(float), 3 * n, f) != 3 * n) { std::fprintf(stderr, "cannot read %s\n", path); std::exit(2); }
    std::fclose(f);
    std::vector<PointStruct> pts(n);
    for (size_t i = 0; i < n; ++i) pts[i].pose = pts[i].local = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);  // Pcl2PointStruct
    return pts;
}

int main(int argc, char** argv) {
    if (argc != 23) { std::fprintf(stderr, "usage\n"); return 2; }
    const std::vector<PointStruct> map_pts = read_points(argv[1], std::strtoull(argv[2], nullptr, 10));
    const std::vector<PointStruct> scan = read_points(argv[3], std::strtoull(argv[4], nullptr, 10));
    Eigen::Matrix4d T0;
    for (int k = 0; k < 16; ++k) T0.data()[k] = std::strtod(argv[5 + k], nullptr);
    RegistrationConfig cfg{};
    cfg.i_max_thread = 10;
    cfg.icp_method = static_cast<IcpMethod>(std::atoi(argv[21]));
    cfg.gicp_cov_search_dist = 0.4;
    cfg.use_radar_cov = false;
    cfg.max_iteration = 6;
    cfg.max_search_dist = 5.0;
    cfg.lm_lambda = 0.5;
    cfg.icp_termination_threshold_m = 0.0;
    cfg.min_overlap_ratio = 0.0;
    cfg.max_fitness_score = 1e30;
    cfg.b_debug_print = false;

    VoxelHashMap local_map_;                       // pcm_matching.cpp:86-101
    local_map_.Init(1.0, 30);
    local_map_.AddPoints(map_pts);
    if (cfg.icp_method == VGICP || cfg.icp_method == AVGICP) local_map_.CalVoxelCovAll();
    if (cfg.icp_method == GICP) local_map_.CalPointCovAll(cfg.gicp_cov_search_dist);
    Registration registration_;                    // pcm_matching.cpp:82
    registration_.Init(cfg);

    bool is_success = true;
    double fitness_score = -7.0;                   // must survive a failure untouched (registration.cpp:415)
    Eigen::Matrix6d local_cov;
    const Eigen::Matrix4d T = registration_.RunRegister(scan, local_map_, T0, cfg, is_success, fitness_score, local_cov);  // :280-282

    std::FILE* o = std::fopen(argv[22], "w");
    std::fprintf(o, "%d\n", local_map_.Empty() ? 1 : 0);
    for (int k = 0; k < 16; ++k) std::fprintf(o, "%.17g ", T.data()[k]);
    std::fprintf(o, "\n%d\n%.17g\n", is_success ? 1 : 0, fitness_score);
    for (int k = 0; k < 36; ++k) std::fprintf(o, "%.17g ", local_cov.data()[k]);
    std::fprintf(o, "\n%zu\n", local_map_.Pointcloud().size());
    std::fclose(o);
    return 0;
}
