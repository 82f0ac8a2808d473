/*
 * b2s.h -- C ABI of the B200-native scan-to-map registration and voxel-map fusion engine.
 *
 * This is the drop-in boundary behind open3d_slam's CloudRegistration / ScanToMapRegistration /
 * Submap interfaces.  The reference has NO C/FFI boundary today: its seam is a pair of abstract C++
 * classes plus factories (paths relative to /root/reference/open3d_slam/open3d_slam/):
 *     include/open3d_slam/CloudRegistration.hpp:19-27      CloudRegistration::registerClouds,
 *                                                          estimateNormalsOrCovariancesIfNeeded
 *     include/open3d_slam/ScanToMapRegistration.hpp:29-38  ScanToMapRegistration::processForScanMatchingAndMerging,
 *                                                          scanToMapRegistration, prepareInitialMap
 *     include/open3d_slam/Submap.hpp:38-45                 Submap::insertScan / insertScanDenseMap / getMapPointCloud
 * Each entry point below names the reference interface (file:line) it replaces.  The C++ subclasses a
 * maintainer adds on the reference side live in shim/ and are described in INTEGRATION.md.
 *
 * Conventions
 *   - plain C99, no torch / CUDA types in any signature (a CUDA stream crosses as void*).
 *   - every function returns int32_t status: B2S_OK (0) or a negative B2S_E_* code; the message of the
 *     last failure on the calling thread is available from b2s_last_error().  No exception crosses.
 *   - host point data is the reference's own memory layout: array-of-xyz doubles, 24-byte stride
 *     (std::vector<Eigen::Vector3d>, include/open3d_slam/typedefs.hpp:23), or float32 xyz with an arbitrary
 *     stride (sensor_msgs/PointCloud2 wire format, open3d_conversions.cpp:61-67).
 *   - 4x4 transforms are 16 doubles in ROW-MAJOR order (Eigen::Matrix4d is column-major: copy element-wise).
 *   - all arithmetic on the device is fp64 like the reference (ScanToMap ICP parity target 1e-4, achieved ~1e-12).
 *   - a handle owns one CUDA stream; calls on one handle are serialised by an internal mutex and are
 *     asynchronous with respect to the host unless they return data to host memory.  Use one handle per
 *     host thread for concurrency (the reference calls registerClouds from 3 threads, SlamWrapper.cpp:228-231).
 *   - the CUDA extension is the only implementation: there is no CPU fallback.
 */
#ifndef B2S_H_
#define B2S_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library itself is built with -fvisibility=hidden */
#endif

#define B2S_OK 0
#define B2S_E_INVALID (-1)      /* bad argument (reference: assert_* / LogError -> std::runtime_error) */
#define B2S_E_CUDA (-2)         /* CUDA runtime failure */
#define B2S_E_EMPTY (-3)        /* empty cloud where the reference asserts non-empty (ScanToMapRegistration.cpp:51-52,60) */
#define B2S_E_NO_NORMALS (-4)   /* point-to-plane target without normals ([O3D] RegistrationICP LogError) */
#define B2S_E_CAPACITY (-5)     /* a fixed-capacity device structure would overflow */
#define B2S_E_UNSUPPORTED (-6)

typedef struct b2s_handle b2s_handle;
typedef struct b2s_cloud b2s_cloud;     /* device-resident point cloud: xyz (+ normals) fp64 */
typedef struct b2s_submap b2s_submap;   /* device-resident sparse map cloud (+ optional dense voxel map) */

/* CroppingVolumeEnum / croppingVolumeFactory  (src/croppers.cpp:27-51, include/open3d_slam/croppers.hpp) */
enum { B2S_CROP_NONE = 0, B2S_CROP_MAX_RADIUS = 1, B2S_CROP_MIN_RADIUS = 2, B2S_CROP_MINMAX_RADIUS = 3, B2S_CROP_CYLINDER = 4 };

/* ScanCroppingParameters (include/open3d_slam/Parameters.hpp:51-57) + the CroppingVolume state
 * (isInvertVolume_, pose_ translation; src/croppers.cpp:57-67).  Only the translation of the pose is used. */
typedef struct b2s_cropper {
  int32_t kind;
  int32_t invert;
  double rmin, rmax, zmin, zmax;
  double center[3];
} b2s_cropper;

/* CloudRegistrationType (Parameters.hpp:37).  All three estimators run on the device; GeneralizedIcp derives the per-point
 * covariances from the clouds' normals like [O3D] does when no covariances are present. */
enum { B2S_REG_POINT_TO_PLANE = 0, B2S_REG_POINT_TO_POINT = 1, B2S_REG_GENERALIZED = 2 };

/* IcpParameters + ICPConvergenceCriteria (Parameters.hpp:66-71; src/CloudRegistration.cpp:58-66; [O3D] defaults
 * relative_fitness = relative_rmse = 1e-6) */
typedef struct b2s_icp_params {
  int32_t reg_type;
  int32_t max_iter;            /* icp.max_n_iter */
  double max_corr_dist;        /* icp.max_correspondence_dist */
  int32_t knn;                 /* icp.knn            (normal estimation) */
  double knn_radius;           /* icp.max_distance_knn */
  double rel_fitness, rel_rmse;
} b2s_icp_params;

/* ScanProcessingParameters + the two croppers ScanToMapIcp holds (Parameters.hpp:59-64, ScanToMapRegistration.cpp:29-33) */
typedef struct b2s_scan_params {
  double voxel_size;           /* scan_processing.voxel_size */
  double downsampling_ratio;   /* scan_processing.downsampling_ratio */
  uint32_t seed;               /* replaces std::random_device in [O3D] RandomDownSample */
  b2s_cropper map_builder_cropper;   /* params_.mapBuilder_.cropper_  : applied to the raw scan (preprocess) */
  b2s_cropper scan_matcher_cropper;  /* params_.scanProcessing_.cropper_ : narrow crop / map patch crop */
} b2s_scan_params;

typedef struct b2s_config {
  b2s_icp_params icp;
  b2s_scan_params scan;
  double map_voxel_size;       /* map_builder.map_voxel_size (Parameters.hpp:94) */
  double dense_voxel_size;     /* dense_map_builder.map_voxel_size */
  double nn_cell_size;         /* 0 = automatic (max_corr_dist / 4) : cell edge of the nearest-neighbour grid */
  int32_t icp_cluster_ctas;    /* 0 = automatic: CTAs (= SMs) one registration may spread over: 8 suits many concurrent registrations
                                * (throughput), 16 a single stream of scans (latency); rounded down to a power of two, at most 16 */
  int32_t reserved_;
} b2s_config;

/* open3d::pipelines::registration::RegistrationResult as read by the callers
 * (src/Mapper.cpp:151-159, src/Odometry.cpp:51-72, src/PlaceRecognition.cpp:118-149) */
typedef struct b2s_result {
  double T[16];                /* transformation_ (row-major) */
  double fitness;              /* fitness_ */
  double inlier_rmse;          /* inlier_rmse_ */
  int32_t n_corr;              /* correspondence_set_.size() */
  int32_t iters;               /* ICP updates applied */
} b2s_result;

void b2s_default_config(b2s_config* cfg);   /* Lua defaults, parameter_structure_definitions.lua:52-72,102, PointToPlaneIcp */

/* ---- engine life cycle : cloudRegistrationFactory / scanToMapRegistrationFactory
 *      (src/CloudRegistration.cpp:85-100, src/ScanToMapRegistration.cpp:91-103) ------------------------------- */
int32_t b2s_create(const b2s_config* cfg, int32_t device, void* cuda_stream_or_null, b2s_handle** out);
void b2s_destroy(b2s_handle* h);
int32_t b2s_set_config(b2s_handle* h, const b2s_config* cfg);    /* ScanToMapIcp::setParameters (ScanToMapRegistration.cpp:24-27) */
int32_t b2s_synchronize(b2s_handle* h);
const char* b2s_last_error(void);
const char* b2s_version(void);
int32_t b2s_device_count(void);
/* number of kernels launched by this handle since creation ("gpu_launches" evidence for bench.py) */
int64_t b2s_launch_count(const b2s_handle* h);

/* per-kernel-group device time, CUDA events on the handle's stream (the reference prints per-stage wall times with
 * o3d_slam::Timer, src/time.cpp:35-78).  kinds: 0 icp, 1 normals, 2 radix sort, 3 NN-grid build, 4 voxel keys+means,
 * 5 fusion, 6 select, 7 crop.  b2s_profile_read synchronises, returns the sums since the last read and resets them. */
#define B2S_PROFILE_KINDS 8
int32_t b2s_profile_enable(b2s_handle* h, int32_t on);
int32_t b2s_profile_read(b2s_handle* h, double* ms_by_kind, int64_t* count_by_kind, int32_t n_kinds);
/* debug aid (1024 words): [0..255] clock64 stamps {start, search end, reduce end, solve end} x 64 evaluations of the ICP kernel (CTA 0),
 * [256..511] per-CTA phase times, [512 + 8 e + 0..3] search statistics of evaluation e < 32 (candidates scanned in phase 1, point
 * evaluations, points queued for phase 2, candidates scanned in phase 2) */
int32_t b2s_debug_icp_clocks(b2s_handle* h, int32_t enable, long long* out_1024);

/* ---- clouds (open3d::geometry::PointCloud points_/normals_) ------------------------------------------------ */
int32_t b2s_cloud_create(b2s_handle* h, b2s_cloud** out);
void b2s_cloud_destroy(b2s_cloud* c);
int32_t b2s_cloud_upload_f64(b2s_handle* h, b2s_cloud* c, const double* xyz, const double* normals_or_null, size_t n);
int32_t b2s_cloud_upload_f32(b2s_handle* h, b2s_cloud* c, const void* xyz, size_t n, size_t stride_bytes);
int32_t b2s_cloud_size(b2s_handle* h, const b2s_cloud* c, size_t* n, int32_t* has_normals);   /* synchronises */
int32_t b2s_cloud_download(b2s_handle* h, const b2s_cloud* c, double* xyz, double* normals_or_null, size_t capacity, size_t* n);
int32_t b2s_cloud_copy(b2s_handle* h, const b2s_cloud* src, b2s_cloud* dst);

/* ---- stages of the hot path (SURVEY.md section 8a row ids) ---------------------------------------------- */
/* P1  CroppingVolume::crop                                   src/croppers.cpp:76-106 */
int32_t b2s_crop(b2s_handle* h, const b2s_cloud* in, const b2s_cropper* cropper, b2s_cloud* out);
/* P2  o3d_slam::voxelize -> [O3D] VoxelDownSample             src/helpers.cpp:107-113 */
int32_t b2s_voxel_down_sample(b2s_handle* h, const b2s_cloud* in, double voxel_size, b2s_cloud* out);
/* P3  estimateNormalsOrCovariancesIfNeeded                    src/CloudRegistration.cpp:49-56 */
int32_t b2s_estimate_normals(b2s_handle* h, b2s_cloud* cloud, int32_t knn, double radius);
/* P4  [O3D] RandomDownSample (seeded)                         src/ScanToMapRegistration.cpp:39 */
int32_t b2s_random_down_sample(b2s_handle* h, const b2s_cloud* in, double ratio, uint32_t seed, b2s_cloud* out);
/* F0  o3d_slam::transform (keeps the near-identity duplication quirk)   src/helpers.cpp:273-305 */
int32_t b2s_transform(b2s_handle* h, const b2s_cloud* in, const double T[16], b2s_cloud* out);
/* S1  ScanToMapIcp::processForScanMatchingAndMerging          src/ScanToMapRegistration.cpp:42-54
 *     raw scan -> merge_ (wide) and match_ (narrow); B2S_E_EMPTY when either is empty (checked lazily) */
int32_t b2s_process_scan(b2s_handle* h, const b2s_cloud* raw, b2s_cloud* merge, b2s_cloud* match);
/* R1  CloudRegistration::registerClouds (point-to-plane)      src/CloudRegistration.cpp:44-48 */
int32_t b2s_register(b2s_handle* h, const b2s_cloud* source, const b2s_cloud* target, const double init[16], b2s_result* out);
/* config 4: n independent registrations in one launch (PlaceRecognition.cpp:71,111 iterates them serially) */
int32_t b2s_register_batch(b2s_handle* h, int32_t n, const b2s_cloud* const* sources, const b2s_cloud* const* targets,
                           const double* inits /* n x 16 */, b2s_result* out /* n */);
/* R3  the correspondence search of [O3D] GetRegistrationResultAndCorrespondences (KDTreeFlann::SearchHybrid(q, r, 1)) on its own:
 *     for every point of `queries`, moved by T first when T is given, the index of its nearest point of `target` with d^2 < r^2
 *     (-1 = none; exact, ties -> lower index) and the squared distance (-1 when none).  The correspondence_set_ of a
 *     RegistrationResult is this call at the result's transformation.  Host arrays hold `capacity` >= query count entries. */
int32_t b2s_nearest_neighbors(b2s_handle* h, const b2s_cloud* queries, const b2s_cloud* target, double max_correspondence_distance,
                              const double T_or_null[16], int32_t* index_out, double* d2_out_or_null, size_t capacity, size_t* n_queries);
/* host-pointer convenience form of R1 used by the C++ shim: uploads, registers, returns. */
int32_t b2s_register_host(b2s_handle* h, const double* src_xyz, size_t n_src, const double* tgt_xyz, const double* tgt_normals,
                          size_t n_tgt, const double init[16], b2s_result* out);

/* ---- submap (Submap::mapCloud_ / denseMap_)   include/open3d_slam/Submap.hpp:38-45 ----------------------- */
int32_t b2s_submap_create(b2s_handle* h, size_t capacity_points, b2s_submap** out);
void b2s_submap_destroy(b2s_submap* sm);
/* F1  Submap::insertScan without carving: transform, append, voxelizeWithinCroppingVolume around the sensor
 *     src/Submap.cpp:39-75, src/helpers.cpp:115-183 */
int32_t b2s_submap_insert(b2s_handle* h, b2s_submap* sm, const b2s_cloud* preprocessed_scan, const double map_to_sensor[16]);
/* C1  Submap::carve of the sparse map (space carving)   src/Submap.cpp:55-60,109-123, src/helpers.cpp:235-271,
 *     src/Voxel.cpp:123-149.  The caller keeps the reference's schedule (nScansInsertedMap_ % carveSpaceEveryNscans_ == 1,
 *     before the scan is appended) and passes the pose the map-builder cropper was LAST set to (the previous insertion:
 *     src/Submap.cpp:59 runs before :71).  raw_scan is in the sensor frame.  n_removed may be NULL (no synchronisation). */
typedef struct b2s_carving_params {     /* SpaceCarvingParameters, include/open3d_slam/Parameters.hpp:85-92 */
  double voxel_size;                    /* voxelSize_ = 0.1 */
  double max_raytracing_length;         /* maxRaytracingLength_ = 20.0 */
  double truncation_distance;           /* truncationDistance_ = 0.1 */
  double min_dot_product_with_normal;   /* minDotProductWithNormal_ = 0.5 */
  double neighborhood_radius_dense_map; /* neighborhoodRadiusDenseMap_ = 0.1 (dense map only) */
} b2s_carving_params;
int32_t b2s_submap_carve(b2s_handle* h, b2s_submap* sm, const b2s_cloud* raw_scan, const double map_to_sensor[16],
                         const double cropper_pose[16], const b2s_carving_params* params, size_t* n_removed);
/* D1  ConstantVelocityMotionCompensation::undistortInputPointCloud (src/MotionCompensation.cpp:64-139): per-point motion
 *     compensation by azimuth phase (computePhase), for the linear / angular (roll-pitch-yaw) velocity the host estimated from
 *     its pose buffer (estimateLinearAndAngularVelocity, :33-57).  in != out; the output carries no normals. */
int32_t b2s_undistort(b2s_handle* h, const b2s_cloud* in, const double linear_velocity[3], const double angular_velocity_rpy[3],
                      double scan_duration, int32_t is_spinning_clockwise, b2s_cloud* out);
/* L1  the two steps either side of the loop-closure ICP (src/PlaceRecognition.cpp:103-111,148; src/constraint_builders.cpp:54,71)
 *     computeIndicesOfOverlappingPoints + SelectByIndex   src/helpers.cpp:307-332 : the points of source / target whose
 *     voxel (edge voxel_size, source moved by source_to_target) holds >= min_points_per_voxel points of BOTH clouds, in
 *     their original order (the reference's index lists come in hash-map order; only the sets matter to its callers).
 *     [O3D] GetInformationMatrixFromPointClouds(source, target, max_correspondence_distance, transformation): 6x6, row-major. */
int32_t b2s_overlap(b2s_handle* h, const b2s_cloud* source, const b2s_cloud* target, const double source_to_target[16], double voxel_size,
                    int32_t min_points_per_voxel, b2s_cloud* source_overlap, b2s_cloud* target_overlap);
int32_t b2s_information_matrix(b2s_handle* h, const b2s_cloud* source, const b2s_cloud* target, double max_correspondence_distance,
                               const double transformation[16], double info_out[36]);
/* C2  Submap::carve of the DENSE map   src/Submap.cpp:86-89,125-136, src/helpers.cpp:347-377, src/VoxelHashMap.cpp:13-45,
 *     src/Voxel.cpp:162-192.  `scan` is used in the frame the caller hands over (the reference passes the raw scan together with
 *     the map-frame sensor position); the every-N-scans schedule stays with the caller.  Uses voxel = dense_voxel_size,
 *     neighborhood_radius_dense_map, truncation_distance, max_raytracing_length.  n_removed may be NULL (no synchronisation). */
int32_t b2s_dense_carve(b2s_handle* h, b2s_submap* sm, const b2s_cloud* scan, const double sensor_position[3],
                        const b2s_carving_params* params, size_t* n_removed);
/* F3  Submap::insertScanDenseMap -> VoxelizedPointCloud::insert           src/Submap.cpp:77-92, src/Voxel.cpp:66-88 */
int32_t b2s_submap_insert_dense(b2s_handle* h, b2s_submap* sm, const b2s_cloud* raw_scan, const double map_to_sensor[16],
                                const b2s_cropper* dense_cropper);
/* F2  VoxelHashMap<Voxel> query interface (include/open3d_slam/VoxelHashMap.hpp:104-158) on the dense map, batched over
 *     the points of a device cloud:  hasVoxelContainingPoint / getVoxelContainingPointPtr -> counts[i] (0 = no voxel) and,
 *     optionally, the aggregated position (AggregatedVoxel::getAggregatedPosition, src/Voxel.cpp:35-40) in means_xyz[3i..];
 *     removeKey(getKey(p)) for every point; size(); clear().  Host arrays must hold `capacity` >= cloud size entries. */
int32_t b2s_dense_query(b2s_handle* h, const b2s_submap* sm, const b2s_cloud* points, int32_t* counts, double* means_xyz, size_t capacity);
int32_t b2s_dense_remove(b2s_handle* h, b2s_submap* sm, const b2s_cloud* points);
int32_t b2s_dense_size(b2s_handle* h, const b2s_submap* sm, size_t* n_voxels);
int32_t b2s_dense_clear(b2s_handle* h, b2s_submap* sm);
/* Submap::transform (loop-closure correction of a whole submap)              src/Submap.cpp:94-107
 *     mapCloud_.Transform(T) ([O3D] PointCloud::Transform: points T p / w, normals R n; no duplication quirk),
 *     denseMap_.transform(T) (src/Voxel.cpp:49-64: applied to the voxel SUMS, keys unchanged -- kept as it is),
 *     mapToRangeSensor_ = mapToRangeSensor_ * T (the pose state of b2s_submap_set_pose / b2s_submap_get_pose). */
int32_t b2s_submap_transform(b2s_handle* h, b2s_submap* sm, const double T[16]);
/* Submap::getMapPointCloud (copy-out)                                       src/Submap.cpp:184-191 */
int32_t b2s_submap_size(b2s_handle* h, const b2s_submap* sm, size_t* n);
int32_t b2s_submap_download(b2s_handle* h, const b2s_submap* sm, double* xyz, double* normals, size_t capacity, size_t* n);
/* Submap::getMapPointCloudCopy (src/Submap.cpp:187-191) without leaving the device: the map cloud as a b2s_cloud (what place
 * recognition, the overlap selection and the voxel map of the revisit check read, src/PlaceRecognition.cpp:69,96, src/Submap.cpp:236) */
int32_t b2s_submap_to_cloud(b2s_handle* h, const b2s_submap* sm, b2s_cloud* out);
int32_t b2s_submap_dense_download(b2s_handle* h, const b2s_submap* sm, double* xyz, double* normals, int32_t* keys, size_t capacity,
                                  size_t* n);
int32_t b2s_submap_set_cloud(b2s_handle* h, b2s_submap* sm, const b2s_cloud* cloud);   /* load / initial map */
/* S2  ScanToMapIcp::scanToMapRegistration: crop the map patch around the sensor, then R1
 *     src/ScanToMapRegistration.cpp:55-62 */
int32_t b2s_register_to_submap(b2s_handle* h, const b2s_cloud* scan, const b2s_submap* sm, const double map_to_sensor[16],
                               const double init[16], b2s_result* out);

/* ---- fused device-side chain for throughput runs: S1 -> S2 -> fitness gate -> F1, no host round trip.
 *      Restates the steady-state branch of Mapper::addRangeMeasurement (src/Mapper.cpp:139-177) with the pose
 *      prediction supplied by the caller.  The result is written to device memory and fetched with
 *      b2s_scan_result_fetch after b2s_synchronize. ------------------------------------------------------------ */
/* the pose state mapToRangeSensor_ lives on the device next to the map (Mapper.hpp mapToRangeSensor_/mapToRangeSensorPrev_) */
int32_t b2s_submap_set_pose(b2s_handle* h, b2s_submap* sm, const double map_to_sensor[16]);
int32_t b2s_submap_get_pose(b2s_handle* h, const b2s_submap* sm, double map_to_sensor[16]);
/* one scan: guess = pose * odometry_motion (Mapper.cpp:130-137); S1; S2 around the pose state; gate
 * (fitness < min_refinement_fitness rejects unless ignore_min_fitness, Mapper.cpp:151); accepted -> pose = result, F1. */
int32_t b2s_mapper_step_async(b2s_handle* h, b2s_submap* sm, const b2s_cloud* raw_scan, const double odometry_motion[16],
                              double min_refinement_fitness, int32_t ignore_min_fitness, int32_t slot /* 0..255 */);
int32_t b2s_scan_result_fetch(b2s_handle* h, int32_t slot, b2s_result* out);   /* synchronises */
/* the same chain end to end with HOST buffers: float32 xyz (wire format; pinned memory keeps the copy asynchronous) in,
 * RegistrationResult out -- upload + S1 + S2 + gate + F1 + read-back in one call (Mapper::addRangeMeasurement as the
 * ROS callback sees it, rospkg/src/OnlineRangeDataProcessorRos.cpp:38-43 -> core/src/Mapper.cpp:101-181) */
int32_t b2s_mapper_step_host(b2s_handle* h, b2s_submap* sm, const void* xyz_f32, size_t n, size_t stride_bytes,
                             const double odometry_motion[16], double min_refinement_fitness, int32_t ignore_min_fitness,
                             b2s_result* out);
/* the same without waiting: the upload, the chain and the copy of the result into *out_pinned (page-locked host memory,
 * valid after the next b2s_synchronize) are only enqueued -- one host thread can keep many mappers (handles) busy */
int32_t b2s_mapper_step_host_async(b2s_handle* h, b2s_submap* sm, const void* xyz_f32, size_t n, size_t stride_bytes,
                                   const double odometry_motion[16], double min_refinement_fitness, int32_t ignore_min_fitness,
                                   b2s_result* out_pinned);
/* CUDA-graph replay of b2s_mapper_step_async for this submap: after two eager steps the ~45 launches of one scan are
 * captured once and replayed with a single cudaGraphLaunch.  Every scan must be uploaded (b2s_cloud_upload_f32/_f64) or
 * copied (b2s_cloud_copy) into the returned fixed-capacity staging cloud, which is then passed as raw_scan; the slot
 * argument must equal (number of graph steps so far) % 256.  Falls back to eager launches when the chain cannot be
 * captured (e.g. a cropper without a maximum radius needs a host round trip). */
int32_t b2s_mapper_graph_enable(b2s_handle* h, b2s_submap* sm, size_t raw_capacity_points, double min_refinement_fitness,
                                int32_t ignore_min_fitness, b2s_cloud** staging_out);

/* ---- options of the device-resident Mapper chain (b2s_mapper_step_async / _host / _host_async) -----------------------
 * What Mapper::addRangeMeasurement and SubmapCollection::insertScan do around S1/S2/F1 with their default wiring:
 *   - the minimum-motion gate in front of the map insertion                       src/Mapper.cpp:170-176
 *   - Submap::insertScan(..., isPerformCarving = true): space carving of the sparse map every carveSpaceEveryNscans_
 *     insertions (nScansInsertedMap_ % N == 1, map not empty), with the cropper at the pose of the LAST insertion
 *                                                                                 src/SubmapCollection.cpp:178,205, src/Submap.cpp:55-60,109-123
 *   - insertScanDenseMap(raw scan, mapToRangeSensor, carving = true) for every scan addRangeMeasurement accepted
 *                                                                                 src/SlamWrapper.cpp:318-327,363-376, src/Submap.cpp:77-92,125-136
 * All decisions are taken ON THE DEVICE from device-resident counters (the host never learns whether a scan passed
 * the fitness gate before it reads the result), so the chain stays free of host round trips and graph-replayable.
 * Options are per submap; changing them drops a captured graph (it is re-captured on the next step). */
typedef struct b2s_mapper_options {
  double min_movement_between_mapping_steps;   /* MapperParameters::minMovementBetweenMappingSteps_ (Parameters.hpp:161), default 0 */
  int32_t carve_enabled;                       /* isPerformCarving of Submap::insertScan */
  int32_t carve_every_n_scans;                 /* SpaceCarvingParameters::carveSpaceEveryNscans_ (Parameters.hpp:89), default 10 */
  b2s_carving_params carving;                  /* mapBuilder_.carving_ */
  int32_t dense_enabled;                       /* feed the dense map with every accepted raw scan */
  int32_t dense_carve_every_n_scans;           /* denseMapBuilder_.carving_.carveSpaceEveryNscans_; 0 = no dense carving */
  b2s_carving_params dense_carving;            /* denseMapBuilder_.carving_ */
  b2s_cropper dense_cropper;                   /* denseMapBuilder_.cropper_ (applied in the sensor frame, Submap.cpp:78-79) */
} b2s_mapper_options;
void b2s_default_mapper_options(b2s_mapper_options* o);
int32_t b2s_submap_set_mapper_options(b2s_handle* h, b2s_submap* sm, const b2s_mapper_options* o);
/* device-side bookkeeping of the chain, read back (synchronises): what the reference keeps in Mapper / Submap members */
typedef struct b2s_mapper_counters {
  int64_t steps;                 /* scans that went through S1 + S2 on this submap */
  int64_t accepted;              /* passed the fitness gate (addRangeMeasurement returned true) */
  int64_t inserted_map;          /* Submap::nScansInsertedMap_ */
  int64_t inserted_dense;        /* Submap::nScansInsertedDenseMap_ */
  int64_t carve_runs;            /* how often the sparse carving actually ran */
  int64_t carved_points_total;   /* map points removed by it */
  int64_t dense_carve_runs;
  int64_t carved_voxels_total;   /* dense voxels emptied */
} b2s_mapper_counters;
int32_t b2s_submap_get_mapper_counters(b2s_handle* h, const b2s_submap* sm, b2s_mapper_counters* out);

/* ---- F4  o3d_slam::VoxelMap (include/open3d_slam/Voxel.hpp:19-36, src/Voxel.cpp:123-160): voxel -> per-layer lists of point indices.
 *      Keys are getVoxelIdx(p, 1 / voxelSize) = floor(p * inv) per axis (VoxelHashMap.hpp:43-50).  Layers are integers
 *      0..B2S_VOXEL_MAP_LAYERS-1 (the shim maps the reference's layer names, e.g. Submap::voxelMapLayer).  The users on this path:
 *      the revisit check of SubmapCollection::isSwitchingSubmapsConsistant (src/SubmapCollection.cpp:352-364) =
 *      b2s_voxel_map_has_voxel over the scan moved by mapToRangeSensor; Submap::computeFeatures fills it with
 *      voxelMap_.clear(); voxelMap_.insertCloud(voxelMapLayer, mapCloud_) (src/Submap.cpp:235-236). ------------------------------ */
#define B2S_VOXEL_MAP_LAYERS 4
typedef struct b2s_voxel_map b2s_voxel_map;
int32_t b2s_voxel_map_create(b2s_handle* h, const double voxel_size[3], size_t capacity_voxels, b2s_voxel_map** out);   /* VoxelMap(voxelSize) */
void b2s_voxel_map_destroy(b2s_voxel_map* vm);
int32_t b2s_voxel_map_clear(b2s_handle* h, b2s_voxel_map* vm);                                                        /* clear() */
int32_t b2s_voxel_map_insert_cloud(b2s_handle* h, b2s_voxel_map* vm, int32_t layer, const b2s_cloud* cloud);          /* insertCloud(layer, cloud) */
int32_t b2s_voxel_map_size(b2s_handle* h, const b2s_voxel_map* vm, size_t* n_voxels);                                 /* size() */
/* hasVoxelContainingPoint for every point of `points` (moved by the isometry T first when T is given): flags (optional, one per
 * point, `capacity` entries) and the number of hits */
int32_t b2s_voxel_map_has_voxel(b2s_handle* h, const b2s_voxel_map* vm, const b2s_cloud* points, const double T_or_null[16],
                                int32_t* flags_or_null, size_t capacity, size_t* n_hits);
/* getIndicesInVoxel(layer, p) for every point: CSR answer, offsets[n + 1] and the concatenated index lists (each sorted ascending =
 * insertion order of insertCloud(layer, cloud)).  indices may be NULL to query the sizes only. */
int32_t b2s_voxel_map_indices_in_voxel(b2s_handle* h, const b2s_voxel_map* vm, int32_t layer, const b2s_cloud* points, int32_t* offsets,
                                       size_t offsets_capacity, int32_t* indices, size_t indices_capacity, size_t* n_indices);
/* copies of the clouds the last mapper step produced on this handle (ProcessedScans of Mapper.cpp:139; SubmapCollection keeps
 * the merge_ cloud in its overlap buffer, src/SubmapCollection.cpp:83-92,180).  Either output may be NULL. */
int32_t b2s_mapper_processed_scan(b2s_handle* h, b2s_cloud* merge_out, b2s_cloud* match_out);

/* ---- device-to-device hand-over of a cloud's arrays (a submap that is the registration target on
 *      several GPUs is built once by its owner and broadcast over NVLink by the host side -- torch.distributed / NCCL own
 *      the transfer, this library only copies between its cloud and the caller's device buffers on the handle's stream).
 *      xyz / normals are 3 x f64 per point, like every cloud.  export: synchronises (the count is returned). -------------- */
int32_t b2s_cloud_export_device(b2s_handle* h, const b2s_cloud* c, void* xyz_dev, void* normals_dev_or_null, size_t capacity_points, size_t* n);
int32_t b2s_cloud_import_device(b2s_handle* h, b2s_cloud* c, const void* xyz_dev, const void* normals_dev_or_null, size_t n);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* B2S_H_ */
