/* elimaloc_b200 — C ABI of the B200-native pcm_matching registration hot path.
 *
 * Drop-in boundary for jaeyoungjo99/ELiMaLoc's per-scan registration.  The reference has no FFI layer:
 * the boundary is the C++ member call
 *     Eigen::Matrix4d Registration::RunRegister(const std::vector<PointStruct>&, const VoxelHashMap&,
 *             const Eigen::Matrix4d&, RegistrationConfig, bool&, double&, Eigen::Matrix6d&)
 *     (src/app/localization/pcm_matching/include/registration.hpp:122-124, src/registration.cpp:274-418)
 * plus the map-side calls the node makes once at start-up (src/pcm_matching.cpp:82-101).
 * shim/registration_shim.hpp re-declares those C++ types on top of this header so pcm_matching.cpp
 * compiles unchanged; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every matrix is ROW-major double (Eigen is column-major: the shim transposes);
 *   - every function returns an elm_status (0 = ok); a non-zero status never throws and leaves outputs untouched
 *     unless stated; elm_last_error() gives the message of the calling thread's last failure;
 *   - thread-agnostic: any thread may call; calls on one handle must be serialised by the caller (the node
 *     already does: both call sites hold mutex_pcl_, pcm_matching.cpp:199,357);
 *   - the product never falls back to a CPU path: without a CUDA device every compute entry point returns
 *     ELM_ERR_CUDA.
 * Reference paths below are relative to src/app/localization/ of the reference repository.
 */
#ifndef ELIMALOC_B200_H
#define ELIMALOC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum elm_status {
    ELM_OK = 0,
    ELM_ERR_INVALID = 1,     /* bad argument (NULL, negative size, unknown method ...) */
    ELM_ERR_CUDA = 2,        /* CUDA runtime error or no device */
    ELM_ERR_NCCL = 3,        /* NCCL error or NCCL library not loadable */
    ELM_ERR_UNSUPPORTED = 4, /* valid in the reference but out of scope here (use_radar_cov = 1) */
    ELM_ERR_RANGE = 5,       /* map extent beyond +-(2^20 - 2) voxels per axis, max_points_per_voxel > 1023 */
    ELM_ERR_STATE = 6,       /* call order (e.g. GICP without elm_map_cal_point_cov) */
    ELM_ERR_IO = 7           /* map file cannot be opened / is not a map file / is truncated or inconsistent */
} elm_status;

/* IcpMethod, pcm_matching/include/registration.hpp:60 */
enum { ELM_P2P = 0, ELM_GICP = 1, ELM_VGICP = 2, ELM_AVGICP = 3 };

/* RegistrationConfig, pcm_matching/include/registration.hpp:62-85 — the fields RunRegister reads
 * (registration.cpp:302-412).  voxel_search_method, doppler_trans_lambda, gicp_cov_search_dist and the ego_to_*
 * members are parsed by the node but never read by the solver and are therefore not part of the ABI.
 * Defaults: config/localization.ini:80-109. */
typedef struct elm_reg_config {
    int32_t icp_method;    /* ELM_P2P .. ELM_AVGICP */
    int32_t max_iteration;
    int32_t max_thread;    /* i_max_thread: CPU thread cap of the reference; ignored on the GPU */
    int32_t use_radar_cov; /* must be 0 (ELM_ERR_UNSUPPORTED otherwise) */
    int32_t debug_print;   /* b_debug_print: per-phase timings to stdout */
    int32_t reserved0;
    double max_search_dist;
    double lm_lambda;
    double icp_termination_threshold_m;
    double min_overlap_ratio;
    double max_fitness_score;
    double range_variance_m;
    double azimuth_variance_deg;
    double elevation_variance_deg;
} elm_reg_config;

typedef struct elm_map elm_map;                   /* VoxelHashMap, voxel_hash_map.hpp:89-335 (device resident) */
typedef struct elm_registration elm_registration; /* Registration, registration.hpp:101-230 (+ device scratch) */

const char* elm_last_error(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int elm_device_count(void);

/* ---- VoxelHashMap ------------------------------------------------------------------------------------------ */
/* VoxelHashMap::Init (voxel_hash_map.cpp:26-29).  device = CUDA ordinal the map lives on. */
int elm_map_create(elm_map** out, double voxel_size, int max_points_per_voxel, int device);
void elm_map_destroy(elm_map* map);
/* VoxelHashMap::AddPoints (voxel_hash_map.cpp:270-285) on packed float xyz[3*n] (Pcl2PointStruct, pcm_matching.hpp:
 * 205-220, widens float fields to double, so float is lossless).  Same order-dependent semantics: insert key =
 * truncation toward zero, first point of a voxel always kept, later ones iff below the cap and no stored point within
 * sqrt(voxel_size^2 / cap).  May be called repeatedly; each call re-publishes the device copy. */
int elm_map_add_points(elm_map* map, const float* xyz, size_t n);
/* VoxelHashMap::CalVoxelCovAll (voxel_hash_map.hpp:183-193): needed before VGICP / AVGICP. */
int elm_map_cal_voxel_cov(elm_map* map);
/* VoxelHashMap::CalPointCovAll (voxel_hash_map.hpp:252-257): needed before GICP. */
int elm_map_cal_point_cov(elm_map* map, double search_dist);
/* Where the three passes above run for a map that lives on a device: 1 (default) = on the GPU (map_build.cu: stable radix sort
 * by voxel + per-voxel replay of the spacing test in arrival order, covariance kernels; bit-identical to the host builder),
 * 0 = host builder.  elm_map_add_points uses the GPU builder for the first call on an empty map (the node's only call,
 * pcm_matching.cpp:87); later calls merge on the host.  The derived search tables are built on the host either way.
 * elm_map_build_times: milliseconds the last AddPoints / CalVoxelCovAll / CalPointCovAll spent in the builder proper. */
int elm_map_set_gpu_build(elm_map* map, int enable);
int elm_map_build_times(const elm_map* map, double ms[3]);
/* VoxelHashMap::Empty (voxel_hash_map.hpp:325) */
int elm_map_empty(const elm_map* map);
size_t elm_map_num_voxels(const elm_map* map);
size_t elm_map_num_points(const elm_map* map);
/* Canonical dump (test hook; stands in for Pointcloud()/Covariances(), voxel_hash_map.cpp:245-265): voxels sorted
 * by key (x, y, z), points in insertion order inside a voxel.  Any pointer may be NULL.
 * keys[3V] counts[V] vmean[3V] vcov[9V] pxyz[3P] pmean[3P] pcov[9P] */
int elm_map_export(const elm_map* map, int32_t* keys, int32_t* counts, double* vmean, double* vcov, float* pxyz,
                   double* pmean, double* pcov);

/* Result shaping: PcmMatching::PublishPcmOdom's pose covariance (pcm_matching.cpp:1082-1098 with
 * NormalizeCovariance / UpdateCovarianceField, pcm_matching.hpp:247-290).  R_ego: rotation of the published ego pose
 * (row-major 3x3); local_cov: RunRegister's 6x6 output; icp_pose_std_m: cfg_.d_icp_pose_std_m (= the fitness score,
 * pcm_matching.cpp:295).  Writes the translation block [0..2][0..2] and the rotation block [3..5][3..5] of the row-major
 * 6x6 `pose_cov`; the other 18 entries are left as they are, as the reference does.  Host arithmetic (a dozen flops). */
int elm_shape_pcm_covariance(const double R_ego[9], const double local_cov[36], double icp_pose_std_m, double pose_cov[36]);

/* VoxelHashMap::FindGroundHeight (voxel_hash_map.hpp:285-322; used by the RViz initial-pose click, pcm_matching.cpp:387):
 * mean z of the up to five lowest stored points within 5 m of (x, y) in the plane; found = 0 (ground_z untouched) when three
 * or fewer points are in range.  Host arithmetic on the sorted map (only the voxel columns that reach the disc are visited;
 * the reference copies the whole map per click). */
int elm_map_find_ground_height(const elm_map* map, double x, double y, double* ground_z, int32_t* found);

/* PCD input: the node fills its map from a .pcd through pcl::io::loadPCDFile<PointType>
 * (pcm_matching.cpp:69-79) and uses x, y, z only (Pcl2PointStruct, pcm_matching.hpp:205-220).  Dependency-free reader of PCD
 * v0.7 (DATA ascii / binary / binary_compressed; x, y, z of TYPE F/I/U, any SIZE).  Points with a non-finite coordinate are
 * dropped (counted in n_dropped).  elm_pcd_read_xyz: xyz may be NULL to query n_points first.
 * elm_map_add_points_pcd = read + elm_map_add_points. */
int elm_pcd_read_xyz(const char* path, float* xyz, size_t capacity_points, size_t* n_points, size_t* n_dropped);
int elm_map_add_points_pcd(elm_map* map, const char* path, size_t* n_points);

/* Built-map file.  The reference reloads its .pcd and rebuilds the whole voxel map at every start of the
 * node (pcm_matching.cpp:69-101); elm_map_save writes everything AddPoints / CalVoxelCovAll / CalPointCovAll produced
 * (canonical points, voxel table, neighbourhood directory, covariances) and elm_map_load brings it back — host arrays
 * validated, then uploaded to `device` (-1: host only) — without touching the raw cloud again. */
int elm_map_save(const elm_map* map, const char* path);
int elm_map_load(elm_map** out, const char* path, int device);

/* Self-check of the neighbourhood directory the P2P/GICP search reads (test hook, host only): every centre key whose
 * 27 voxels (GetAdjacentVoxels range 2, voxel_hash_map.cpp:232-241) hold a stored point must be found by the 2-bucket
 * lookup, its nine column descriptors must equal the canonical arrays, and keys outside that set must miss.
 * entries = centre keys stored, slots = table slots, mismatches = violations found (0 = consistent). */
int elm_map_directory_check(const elm_map* map, uint64_t* entries, uint64_t* slots, uint64_t* mismatches);

/* ---- Registration ------------------------------------------------------------------------------------------- */
/* stream: a cudaStream_t (as void*) every kernel of this handle is launched on; NULL = a private stream. */
int elm_registration_create(elm_registration** out, int device, void* stream);
void elm_registration_destroy(elm_registration* reg);

/* Registration::RunRegister (registration.cpp:274-418) — HOST buffers in, host results out, synchronous.
 *   src_xyz[3*n]          source_local: sensor-frame points (pose == local on entry, pcm_matching.hpp:213-214)
 *   T_init[16]            initial_guess
 *   T_out[16]             return value
 *   is_success            written on every return path exactly as the reference does
 *   fitness_score         written only on success (registration.cpp:415)
 *   local_cov[36]         Identity, overwritten by GICP only (registration.cpp:280,142)
 * n == 0 is undefined in the reference (0/0 overlap ratio); here: is_success = 0, T_out = T_init. */
int elm_run_register(elm_registration* reg, const elm_map* map, const float* src_xyz, size_t n,
                     const double T_init[16], const elm_reg_config* cfg, double T_out[16], int32_t* is_success,
                     double* fitness_score, double local_cov[36]);

/* Same registration with the scan already resident in HBM (d_src_xyz = device pointer to packed float xyz[3*n]).
 * Enqueues the whole ICP loop on the handle's stream and returns without synchronising; elm_register_fetch
 * synchronises and returns the results of the last enqueue.  This is what bench.py times as `value`. */
int elm_register_enqueue(elm_registration* reg, const elm_map* map, const float* d_src_xyz, size_t n,
                         const double T_init[16], const elm_reg_config* cfg);
int elm_register_fetch(elm_registration* reg, double T_out[16], int32_t* is_success, double* fitness_score,
                       double local_cov[36], int32_t* iterations_run);

/* One correspondence search + one AlignClouds* accumulation at a fixed pose, no solve (test hook for per-iteration
 * parity: registration.cpp:28-51 / 85-132 / 171-208).  Host buffers.  JTJ[36] full symmetric 6x6. */
int elm_linearize(elm_registration* reg, const elm_map* map, const float* src_xyz, size_t n, const double T[16],
                  const elm_reg_config* cfg, double JTJ[36], double JTr[6], double* residual_sum, int64_t* n_corr);

/* Correspondence dump (test hook for index-level parity of voxel_hash_map.cpp:31-206).  K = 7 for AVGICP else 1.
 * count[n]; target[n*K*3]: matched map point (P2P/GICP) or voxel mean (VGICP/AVGICP); unmatched entries are 0. */
int elm_correspondences(elm_registration* reg, const elm_map* map, const float* src_xyz, size_t n, const double T[16],
                        int method, double max_search_dist, int32_t* count, double* target);

/* The same dump after a SEQUENCE of poses T_seq[n_poses][16] searched one after the other exactly as the ICP loop does
 * (first pose: cold search, following poses: warm-started from the previous pose's matches); returns the correspondences
 * at the last pose.  Test hook: the warm-started search must return what GetCorrespondencePoints returns at that pose. */
int elm_correspondences_sequence(elm_registration* reg, const elm_map* map, const float* src_xyz, size_t n, const double* T_seq,
                                 int n_poses, int method, double max_search_dist, int32_t* count, double* target);

/* Counters of the last enqueue (kernel launches issued on the stream; used by bench.py's gpu_launches). */
int elm_registration_launch_count(const elm_registration* reg, int64_t* launches);

/* Per-kernel timing: when enabled, the search kernel and the accumulate(+reduce+solve) kernel of every ICP iteration
 * are bracketed by cudaEvents on the launch stream and elm_register_fetch accumulates the elapsed times.  bench.py's
 * roofline uses it; in the reference the matching instrumentation is the per-iteration chrono print of
 * registration.cpp:307-347. */
int elm_registration_set_profiling(elm_registration* reg, int enable);
int elm_registration_profile(const elm_registration* reg, double* search_ms, double* accumulate_ms, int64_t* iterations);
/* The same two sums split by kind of iteration: k = 0 cold iteration (search kernel, accumulate kernel), k = 1 first warm
 * iteration of a call (reuse kernel, refresh kernel doing the bulk refresh), k = 2 later warm iterations (reuse kernel,
 * refresh kernel).  ms[2 k] / ms[2 k + 1] = first / second kernel of the iteration, iterations[k] = iterations timed. */
int elm_registration_profile_by_kind(const elm_registration* reg, double ms[6], int64_t iterations[3]);

/* Counters of the P2P/GICP search (off by default; bench.py uses them to state the bytes the search actually has to
 * read): map points visited and queries searched since the last elm_registration_set_stats call. */
int elm_registration_set_stats(elm_registration* reg, int enable);
int elm_registration_stats(elm_registration* reg, uint64_t* map_points_visited, uint64_t* queries);
/* All 32 counters ([0] candidate points read by the searches, [1] searches, [20] warm searches handed to the refresh kernel,
 * [21] warm searches; the rest are developer cycle counters of -DELM_PHASE_TIMING builds). */
int elm_registration_stats_raw(elm_registration* reg, uint64_t counters[32]);

/* P2P / GICP kernel structure.  Default (0): a search kernel (match[] out) followed by an accumulate+reduce+solve kernel.
 * 1: search, linearisation, block/grid reduction and the 6x6 solve in ONE kernel per ICP iteration — same results to
 * rounding (the per-block summation order differs); measured slower on B200 (register pressure), kept as an option. */
int elm_registration_set_fused(elm_registration* reg, int enable);

/* Spatial binning of the scan (default OFF — measured slower on B200 for the pruned search): once per call the scan is counting-sorted on the device by the voxel each
 * point falls into under the initial guess, and the SEARCH walks it in that order so that neighbouring queries share
 * cache lines of the map.  The accumulation keeps the caller's order, so results do not depend on this switch. */
int elm_registration_set_binning(elm_registration* reg, int enable);

/* Search strategy of P2P/GICP.  Default (0): exact pruning — a voxel of the 27-neighbourhood is skipped when its bounding
 * box is provably farther than the best candidate already found, which cannot change the result.  1: visit all 27
 * voxels exactly like GetCorrespondencePoints (voxel_hash_map.cpp:40-51) — same answers, more bytes. */
int elm_registration_set_exhaustive(elm_registration* reg, int exhaustive);

/* Warm start of the P2P/GICP search (default 1).  From the second iteration of a call on, the point matched in the
 * previous iteration bounds the nearest distance before the map is read, and only the octants (half-voxel cells) of the
 * 27 voxels within that bound are visited — identical correspondences (GetCorrespondencePoints, voxel_hash_map.cpp:31-88),
 * a fraction of the bytes.  0: every iteration runs the cold search. */
int elm_registration_set_warm_start(elm_registration* reg, int enable);

/* ---- deskew --------------------------------------------------------------------------------------------------- */
/* Inputs of the per-point deskew = the member tables PcmMatching::ImuDeskewInfo / OdomDeskewInfo fill on the host
 * (pcm_matching.cpp:533-729; vec_d_imu_time_, vec_d_imu_rot_{x,y,z}_, i_imu_pointer_cur_, f_odom_incre_{x,y,z}_,
 * d_time_scan_cur_, d_time_scan_end_, b_is_imu_available_, b_is_odom_available_). */
typedef struct elm_deskew_tables {
    const double* imu_time;  /* host arrays, imu_pointer_cur + 1 valid entries each */
    const double* imu_rot_x;
    const double* imu_rot_y;
    const double* imu_rot_z;
    int32_t imu_pointer_cur;
    int32_t imu_available;
    int32_t odom_available;
    int32_t reserved0;
    float odom_incre_x, odom_incre_y, odom_incre_z, reserved1;
    double time_scan_cur, time_scan_end;
} elm_deskew_tables;

/* The tbb::parallel_for over DeskewPoint of PcmMatching::DeskewPointCloud (pcm_matching.cpp:499-511, 780-824) with
 * FindRotation / FindPosition (:731-778): xyz[3n] + rel_time[n] (PointXYZIT::time, seconds after scan start) ->
 * xyz_out[3n], float32 like the reference.  Host buffers, synchronous; runs on the registration handle's device/stream. */
int elm_deskew_points(elm_registration* reg, const float* xyz, const float* rel_time, size_t n, const elm_deskew_tables* tables,
                      float* xyz_out);
/* Same with device buffers (d_xyz_out may feed elm_register_enqueue directly); asynchronous on the handle's stream. */
int elm_deskew_points_device(elm_registration* reg, const float* d_xyz, const float* d_rel_time, size_t n,
                             const elm_deskew_tables* tables, float* d_xyz_out);

/* Scan pre-processing — the two serial steps the node runs on every scan before RunRegister, as one stable
 * compaction on the GPU:
 *   FilterPointsByDistance (pcm_matching.cpp:451-465): keep iff sqrt(x^2 + y^2 + z^2) <= max_dist (float32 arithmetic);
 *   VoxelDownsample (voxel_hash_map.hpp:260-283): keep the first point (input order) of every voxel floor(p / voxel_size).
 * max_dist <= 0 / voxel_size <= 0 switch a step off (the node filters the raw scan, deskews, then down-samples: call it
 * twice around elm_deskew_points_device).  Survivors keep their input order (the reference emits them in unordered_map
 * order, which is implementation-defined).  aux / aux_out: one float per point carried along (the relative time stamp);
 * index_out: input index of every survivor; both optional.  n_out: survivors. */
int elm_scan_preprocess(elm_registration* reg, const float* xyz, const float* aux, size_t n, double max_dist, double voxel_size,
                        float* xyz_out, float* aux_out, int32_t* index_out, size_t* n_out);
int elm_scan_preprocess_device(elm_registration* reg, const float* d_xyz, const float* d_aux, size_t n, double max_dist,
                               double voxel_size, float* d_xyz_out, float* d_aux_out, int32_t* d_index_out, size_t* n_out);

/* ---- EKF (ekf_localization) ------------------------------------------------------------------------------------ */
/* The 27-state filter of EkfAlgorithm (README: "24-DOF"; STATE_ORDER is 27, ekf_algorithm.hpp:41-69) with its state and
 * covariance resident in HBM.  In scope: Init, RunPredictionImu, RunGnssUpdate for the PCM / PCM_INIT sources,
 * UpdateEkfState, ComplementaryKalmanFilter, the Check* flags, GetCurrentState.  CAN / ZUPT / NavSat / IMU-mount
 * calibration branches are off in config/localization.ini and not provided. */
#define ELM_EKF_STATE_ORDER 27

/* EkfLocalizationConfig fields those functions read (ekf_localization_config.hpp:20-95; INI names in the comments) */
typedef struct elm_ekf_config {
    double imu_gravity;           /* imu_gravity */
    double ekf_init_x_m, ekf_init_y_m, ekf_init_z_m, ekf_init_roll_deg, ekf_init_pitch_deg, ekf_init_yaw_deg;
    double state_std_pos_m;       /* ekf_state_uncertainty_pos_m */
    double state_std_rot_deg;     /* ekf_state_uncertainty_rot_deg */
    double state_std_vel_mps;     /* ekf_state_uncertainty_vel_mps */
    double imu_std_gyro_dps;      /* ekf_imu_uncertainty_gyro_dps */
    double imu_std_acc_mps;       /* ekf_imu_uncertainty_acc_mps */
    double imu_bias_cov_gyro;     /* ekf_imu_bias_cov_gyro */
    double imu_bias_cov_acc;      /* ekf_imu_bias_cov_acc */
    int32_t imu_estimate_gravity; /* imu_estimate_gravity */
    int32_t use_complementary_filter;
    int32_t reserved[2];
} elm_ekf_config;

/* Members of EkfAlgorithm (ekf_algorithm.hpp:269-289: S_, P_, flags, prev_timestamp_, prev_ego_state_) plus the
 * function-static memory of ComplementaryKalmanFilter (ekf_algorithm.cpp:613-614).  Quaternions are (w, x, y, z);
 * P is row-major 27 x 27 in the S_X .. S_IMU_YAW order. */
typedef struct elm_ekf_state {
    double pos[3], rot[4], vel[3], gyro[3], acc[3], bg[3], ba[3], grav[3], imu_rot[4];
    double P[ELM_EKF_STATE_ORDER * ELM_EKF_STATE_ORDER];
    double prev_timestamp, prev_gnss_timestamp;
    double ckf_prev_vel_local_x, ckf_prev_time;
    double ego[26], ego_prev_timestamp;
    int32_t reset_for_init_prediction, state_initialized, yaw_initialized, rotation_stabilized, state_stabilized;
    int32_t pcm_init_on_going, pcm_update_count, ckf_has_prev, predictions, updates, reserved[2];
} elm_ekf_state;

/* EkfGnssMeasurement (localization_struct.hpp:146-153); source follows the GnssSource enum: 3 = PCM, 4 = PCM_INIT */
typedef struct elm_ekf_measurement {
    double timestamp, pos[3], rot[4], pos_cov[9], rot_cov[9];
    int32_t source, reserved;
} elm_ekf_measurement;

typedef struct elm_ekf elm_ekf;
/* EkfAlgorithm::EkfAlgorithm + Init (ekf_algorithm.cpp:13-66) */
int elm_ekf_create(elm_ekf** out, const elm_ekf_config* cfg, int device, void* stream);
void elm_ekf_destroy(elm_ekf* ekf);
/* RunPredictionImu (ekf_algorithm.cpp:167-316): one single-block kernel, asynchronous on the handle's stream */
int elm_ekf_predict_imu(elm_ekf* ekf, double timestamp, const double gyro[3], const double acc[3]);
/* RunGnssUpdate for the PCM / PCM_INIT sources (ekf_algorithm.cpp:318-432): one kernel, asynchronous */
int elm_ekf_update_pose(elm_ekf* ekf, const elm_ekf_measurement* meas);
/* raw state down / up (synchronises the stream) */
int elm_ekf_get_state(elm_ekf* ekf, elm_ekf_state* out);
int elm_ekf_set_state(elm_ekf* ekf, const elm_ekf_state* in);
/* GetCurrentState (ekf_algorithm.cpp:778-833): ego[26] = timestamp, x y z, roll pitch yaw, roll/pitch/yaw rate,
 * vx vy vz (local), ax ay az (local), x/y/z_cov_m (local), latitude/longitude/height std, roll/pitch/yaw cov, 0 */
int elm_ekf_get_current_state(elm_ekf* ekf, double ego[26]);

/* ---- deskew tables + the device-resident scan chain (config 5) --------------------------------------------------- */
/* The node's two message queues as plain arrays (what PcmMatching keeps in deq_imu_ / deq_odom_):
 *   IMU       stamp[n], gyro[3 n] = the angular velocity ImuAngular2RosAngular returns (pcm_matching.hpp)
 *   odometry  stamp[n], position[3 n], orientation (x, y, z, w)[4 n], twist.linear[3 n] (body frame), twist.angular[3 n] */
typedef struct elm_imu_queue { const double* stamp; const double* gyro; size_t n; } elm_imu_queue;
typedef struct elm_odom_queue {
    const double* stamp; const double* pos; const double* quat_xyzw; const double* lin_vel; const double* ang_vel; size_t n;
} elm_odom_queue;
#define ELM_IMU_QUEUE_LENGTH 2000 /* pcm_matching.hpp:113 */

/* PcmMatching::ImuDeskewInfo + OdomDeskewInfo (pcm_matching.cpp:533-729) — SURVEY row a18, host arithmetic.  Fills `tables`
 * (time_scan_cur / time_scan_end must be set by the caller: DeskewPointCloud :473-486); its four table pointers are set to
 * storage[0 .. 4 * ELM_IMU_QUEUE_LENGTH) (caller-owned).  imu_drop / odom_drop (may be NULL): messages the node pops from the
 * front of its queues.  Unavailable data is not an error: tables->imu_available / odom_available say so, as the members do. */
int elm_deskew_build_tables(const elm_imu_queue* imu, const elm_odom_queue* odom, elm_deskew_tables* tables, double* storage,
                            size_t* imu_drop, size_t* odom_drop);

/* The per-scan chain of PcmMatching::CallbackPointCloud (pcm_matching.cpp:235-299) with the point data resident in HBM from
 * the raw scan to the pose: ONE upload of the raw scan, then FilterPointsByDistance -> DeskewPointCloud's point loop ->
 * VoxelDownsample -> RunRegister on the registration's stream; the stages hand their point counts to each other through HBM.
 * The host reads back 12 bytes (the counts, needed to size the ICP launches) and the 1.2 KB result block.  With an elm_ekf, the
 * result is folded into the filter by a kernel that reads the IcpState where it lies (elm_scan_pipeline_ekf_update). */
typedef struct elm_scan_pipeline_config {
    double input_max_dist;      /* cfg_.d_input_max_dist (FilterPointsByDistance); <= 0: off */
    double input_voxel_ds_m;    /* cfg_.d_input_voxel_ds_m (VoxelDownsample); <= 0: off */
    int32_t run_deskew;         /* cfg_.b_run_deskew */
    int32_t lidar_scan_time_end;/* cfg_.b_lidar_scan_time_end: the message stamp is the scan END, point times are <= 0 */
    double tf_ego_to_lidar[16]; /* cfg_.tf_ego_to_lidar, row-major */
} elm_scan_pipeline_config;
typedef struct elm_scan_result {
    double T_lidar[16];   /* RunRegister's return value */
    double T_ego[16];     /* icp_lidar_pose * tf_ego_to_lidar^-1 (pcm_matching.cpp:298) */
    double fitness_score; /* valid when is_success */
    double local_cov[36];
    double pose_cov[36];  /* PublishPcmOdom's covariance blocks (elm_shape_pcm_covariance), valid when is_success */
    double time_scan_cur, time_scan_end;
    int32_t is_success, iterations, n_raw, n_after_filter, n_registered, deskew_ok;
} elm_scan_result;
typedef struct elm_scan_pipeline elm_scan_pipeline;
int elm_scan_pipeline_create(elm_scan_pipeline** out, elm_registration* reg, const elm_scan_pipeline_config* cfg);
void elm_scan_pipeline_destroy(elm_scan_pipeline* p);
/* Stage 1 (pcm_matching.cpp:235-243, DeskewPointCloud :467-531): uploads the raw scan (xyz[3n], point_time[n] = PointXYZIT::time),
 * derives the time base from the first / last point that passes the distance filter, builds the deskew tables on the host
 * (a18) and enqueues filter + deskew.  deskew_ok = 0 (nothing enqueued) when the IMU or odometry data does not cover the scan,
 * or no point survives the filter — the node drops such a scan.  time_scan_end: what GetInterpolatedPose is asked for next. */
int elm_scan_pipeline_deskew(elm_scan_pipeline* p, const float* xyz, const float* point_time, size_t n, double stamp,
                             const elm_imu_queue* imu, const elm_odom_queue* odom, double* time_scan_cur, double* time_scan_end,
                             int32_t* deskew_ok);
/* Stage 2 (pcm_matching.cpp:253-282): VoxelDownsample of the deskewed cloud and RunRegister from sync_lidar_pose, enqueued;
 * returns after reading the point counts (12 bytes). */
int elm_scan_pipeline_register(elm_scan_pipeline* p, const elm_map* map, const double sync_lidar_pose[16], const elm_reg_config* cfg);
/* CallbackPcmOdom (ekf_localization.cpp:147-179) + GnssTimeCompensation (:323-394) + RunGnssUpdate, fed from the IcpState in
 * HBM: success gate, ego pose, PublishPcmOdom's covariance shaping, time compensation against the ring of EgoStates the filter
 * keeps in HBM (elm_ekf_enable_state_ring), update.  Asynchronous; ordered after the registration by an event. */
int elm_scan_pipeline_ekf_update(elm_scan_pipeline* p, elm_ekf* ekf);
/* Synchronises and returns the result of the last elm_scan_pipeline_register. */
int elm_scan_pipeline_fetch(elm_scan_pipeline* p, elm_scan_result* out);
/* PublishInThread's deque of EgoStates (ekf_localization.cpp:398-410) kept in HBM: when enabled every elm_ekf_predict_imu is
 * followed by GetCurrentState + the deque's push rule on the device (1000 entries). */
int elm_ekf_enable_state_ring(elm_ekf* ekf, int enable);

/* ---- multi-GPU (one process per GPU; scan sharded over ranks, map replicated) --------------------------------- */
/* unique_id: 128 bytes.  Rank 0 fills it with elm_comm_unique_id and hands it to the other ranks (bench.py uses
 * torch.distributed for that); every rank then calls elm_registration_set_comm.  After that each RunRegister sums the
 * 29 accumulators (21 JtJ + 6 Jtr + residual + count) over ranks with one ncclAllReduce per ICP iteration and `n`
 * is this rank's shard. */
int elm_comm_unique_id(uint8_t unique_id[128]);
int elm_registration_set_comm(elm_registration* reg, const uint8_t unique_id[128], int rank, int world_size);

/* Peer-memory exchange (preferred on one NVLink/NVSwitch box): instead of ncclAllReduce + a separate solve launch, the
 * last block of the accumulation kernel writes its 32 sums into a mailbox of EVERY rank (CUDA IPC mapping, NVLink stores),
 * waits for the other ranks' flags and sums the mailbox slots in rank order, then solves — one kernel, no collective launch.
 *   1. every rank: elm_registration_peer_export -> 64-byte cudaIpcMemHandle of its mailbox
 *   2. exchange the handles (any host transport; the tests use torch.distributed all_gather)
 *   3. every rank: elm_registration_peer_attach(handles[world_size][64], rank, world_size), then a host barrier
 * world_size <= 8.  A rank that never arrives makes the others time out after ~3 s (ELM_ERR_NCCL) instead of hanging.
 * Takes precedence over elm_registration_set_comm while attached. */
int elm_registration_peer_export(elm_registration* reg, uint8_t handle[64]);
int elm_registration_peer_attach(elm_registration* reg, const uint8_t* handles, int rank, int world_size);
int elm_registration_peer_detach(elm_registration* reg);

#ifdef __cplusplus
}
#endif
#endif /* ELIMALOC_B200_H */
