// common.cuh -- shared device helpers and the host-side runtime types of the b2s engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b2s.h"

namespace b2s {

// ------------------------------------------------------------------------------------------------
// error plumbing: no exception crosses the C ABI; the message is kept per host thread
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define B2S_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::b2s::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return B2S_E_CUDA;                                                                          \
    }                                                                                             \
  } while (0)

#define B2S_TRY(expr)              \
  do {                             \
    int32_t _s = (expr);           \
    if (_s != B2S_OK) return _s;   \
  } while (0)

#define B2S_REQUIRE(cond, code, ...)     \
  do {                                   \
    if (!(cond)) {                       \
      ::b2s::set_error(__VA_ARGS__);     \
      return (code);                     \
    }                                    \
  } while (0)

// ------------------------------------------------------------------------------------------------
// grow-only device buffer
// ------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool tracked = true;   // a re-allocation invalidates captured graphs (false for clouds the caller owns: a graph never holds them)
  int32_t ensure(size_t bytes, cudaStream_t s, bool preserve = false);
  void release();
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// device-side status word: kernels OR error bits into it, the host checks it when it synchronises
enum : uint32_t { ST_KEY_OVERFLOW = 1u, ST_CAPACITY = 2u, ST_EMPTY = 4u, ST_HASH_FULL = 8u };

struct GridHeader {      // one per nearest-neighbour grid, lives in device memory
  double origin[3];
  double cell, inv_cell;
  int32_t dims[3];
  int32_t ncell;
  int32_t n;             // points indexed
  int32_t pad;
};

// dense-grid nearest-neighbour index over a point set (K-index): points counting-sorted by cell
struct GridIndex {
  DevBuf hdr;            // GridHeader
  DevBuf bbox;           // 6 x uint64 ordered-double min/max
  DevBuf cell_start;     // int32 [cap_cells + 1]
  DevBuf rank;           // int32 per point: rank within its cell (-1 = not indexed)
  DevBuf pts;            // double4 per indexed point: x,y,z, bits(original index)
  DevBuf nrm;            // double4 per indexed point: nx,ny,nz,0
  int32_t cap_cells = 0;
  void release();
};

struct ScanScratch {     // decoupled look-back scan state
  DevBuf state;          // uint64 per tile + int32 tile counter
  int32_t cap_tiles = 0;
};

struct ScanJob {         // one array of a batched scan (scan_exclusive_i32_batch)
  const int32_t* in; int32_t* out; const int32_t* d_n; unsigned long long* state; int32_t* counter;
};

struct SortScratch {
  DevBuf hist;           // 256 x nblocks int32
  DevBuf keys_alt, vals_alt;
};

}  // namespace b2s

struct b2s_cloud {
  b2s_handle* h = nullptr;
  int device = 0;
  b2s::DevBuf xyz;       // 3 x f64 per point (the reference's AoS layout)
  b2s::DevBuf nrm;       // 3 x f64 per point
  b2s::DevBuf dn;        // int32 device-side point count
  size_t n_max = 0;      // host-side upper bound of the count (launch sizing)
  size_t fixed_cap = 0;  // non-zero: n_max is pinned to this capacity (graph replay needs constant launch dimensions)
  long long n_known = 0; // exact count when the host knows it, -1 otherwise
  bool has_normals = false;
};

struct b2s_submap {
  b2s_handle* h = nullptr;
  int device = 0;
  b2s_cloud* cloud[2] = {nullptr, nullptr};  // ping-pong map cloud (mapCloud_)
  int cur = 0;
  size_t capacity = 0;
  // dense map (VoxelizedPointCloud): open-addressing hash of running sums
  b2s::DevBuf dense_keys;    // uint64 packed key, EMPTY = ~0
  b2s::DevBuf dense_sum;     // 6 x f64 per slot
  b2s::DevBuf dense_cnt;     // int32 per slot
  b2s::DevBuf dense_used;    // int32 occupied-slot counter
  size_t dense_cap = 0;
  double dense_voxel = 0.0;
  bool dense_has_normals = false;
  b2s::DevBuf pose;          // 4 x (4x4 f64): [0] mapToRangeSensor_ state, [1] insertion pose, [2] odometry motion, [3] initial guess
  // asynchronous read-back of the map size (keeps the host-side launch bound tight without ever synchronising)
  int32_t* pinned_cnt = nullptr;
  cudaEvent_t cnt_ev = nullptr;
  bool cnt_pending = false;
  size_t adds_after_readback = 0;
  // CUDA-graph replay of the per-scan chain (b2s_mapper_graph_enable): every launch dimension is derived from fixed
  // capacities, the per-step inputs (odometry motion, result slot) come from a ring indexed by a device-side counter
  bool graph_mode = false;
  int graph_warm = 0;                 // eager steps still to run before the capture (sizes every scratch buffer)
  cudaGraphExec_t gexec = nullptr;
  int64_t graph_kernels = 0;          // kernels per replay (for the launch counter)
  b2s_cloud* staging = nullptr;       // fixed-capacity input cloud the caller uploads each scan into
  double* odom_ring = nullptr;        // pinned, device-mapped: 64 x (4x4) odometry motions
  long long host_step = 0;
  b2s::DevBuf gstate;                 // int32 [0] device step counter, [1] current result slot
  double g_min_fitness = 0.0;
  int g_ignore_fitness = 0;
  // persistent voxel hash of the map cloud (K-fuse, fuse.cu): map-voxel key -> chain of the map points inside that voxel
  b2s::DevBuf vkeys;         // uint64 [vcap] packed voxel key, EMPTY = ~0
  b2s::DevBuf vhead;         // int32 [vcap] first member of the voxel's chain (-1 = none); members >= FUSE_STAGE_BASE are staged scan points
  b2s::DevBuf vstamp;        // int32 [vcap] stamp of the last insertion that touched the voxel
  b2s::DevBuf vnext;         // int32 [capacity] chain link of every map point
  b2s::DevBuf pstamp;        // int32 [capacity] stamp of the insertion that last rewrote the point
  b2s::DevBuf stage_xyz, stage_nrm, stage_next, stage_in;   // the transformed scan of the insertion in flight
  b2s::DevBuf touched;       // int32 voxels (table slots) the insertion in flight touched
  b2s::DevBuf dups;          // int32 [2][FUSE_DUP_CAP] voxels holding more than one map point (ping-pong)
  size_t vcap = 0;
  size_t stage_cap = 0;
  // Mapper / SubmapCollection wiring of the device chain (b2s_mapper_options) and its device-side state words (MS_*)
  b2s_mapper_options opts;
  b2s::DevBuf mstate;                 // int32 [MS_WORDS]: gates and counters of the chain, see the MS_* indices below
  unsigned long long graph_alloc_gen = 0;   // value of b2s::g_alloc_generation when the graph was captured
  unsigned long long graph_cfg_gen = 0;     // value of b2s_handle::cfg_gen when the graph was captured
};

namespace b2s {
// device-side state words of the mapper chain (b2s_submap::mstate)
enum MapperStateWord {
  MS_ACCEPT = 0,      // this step passed the fitness gate (Mapper.cpp:151)
  MS_INSERT = 1,      // ... and the minimum-motion gate (Mapper.cpp:170-176): F1 runs
  MS_CARVE = 2,       // sparse-map carving runs in this step (Submap.cpp:111)
  MS_DENSE = 3,       // dense-map insertion runs in this step
  MS_DCARVE = 4,      // dense-map carving runs in this step (Submap.cpp:127)
  MS_NINS = 5,        // Submap::nScansInsertedMap_
  MS_NDENSE = 6,      // Submap::nScansInsertedDenseMap_
  MS_NSTEPS = 7, MS_NACCEPT = 8, MS_NCARVE = 9, MS_CARVED = 10, MS_NDCARVE = 11, MS_DCARVED = 12,
  MS_CARVE_N = 13,    // point count the carving compaction works on (0 when carving is skipped)
  MS_TMP = 14,        // [14] grid-wide ticket, [15] dense-carve removed count
  MS_NDEAD = 16,      // tombstones among the map slots (points merged away; xyz = NaN) -- the map holds dn - NDEAD points
  MS_STAMP = 17,      // stamp of the last committed insertion
  MS_NTOUCHED = 18,   // voxels touched by the insertion in flight
  MS_DUPSEL = 19,     // which half of `dups` is current
  MS_NDUP = 20,       // [20], [21]: entries of the two halves
  MS_VUSED = 22,      // occupied slots of the voxel table
  MS_TICKET2 = 23,
  MS_WORDS = 32
};
// bumped by every DevBuf re-allocation: a captured graph holds raw pointers of the scratch buffers, so a graph captured
// under an older generation is re-captured before it is replayed again
extern unsigned long long g_alloc_generation;
// per-kernel-group device timing with CUDA events on the launching stream (bench.py's roofline numbers)
enum ProfKind { PK_ICP = 0, PK_NORMALS, PK_SORT, PK_GRID, PK_VOXEL, PK_FUSE, PK_SELECT, PK_CROP, PK_COUNT };
struct ProfRec { int kind; cudaEvent_t a, b; };
}  // namespace b2s

namespace b2s {
// set while this host thread is inside cudaStreamBeginCapture/EndCapture: growing a device buffer is impossible there
extern thread_local bool g_capturing;
extern thread_local bool g_capture_broken;
}  // namespace b2s

struct b2s_handle {
  bool prof_enabled = false;
  long long* icp_dbg = nullptr;       // optional device buffer of clock64 stamps (b2s_debug_icp_clocks)
  std::vector<b2s::ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::recursive_mutex mu;            // recursive: the composite host entry points hold it across the calls they chain
  b2s_config cfg;
  unsigned long long cfg_gen = 1;     // bumped by b2s_set_config: captured graphs bake the configuration in
  int64_t launches = 0;

  b2s::DevBuf status;                 // uint32 device status word
  b2s::ScanScratch scan;
  b2s::SortScratch sort;
  b2s::GridIndex grid_a, grid_b;      // target index (ICP) / self index (normals)
  // generic scratch buffers, by role
  b2s::DevBuf keys, vals, flags, offs, tmp_i32, tmp_f64, misc, work_xyz;
  b2s::DevBuf problems;               // IcpProblem array (device)
  b2s::DevBuf results;                // b2s_result array (device)
  b2s::DevBuf slots;                  // per-slot results for the async mapper step
  b2s::DevBuf poses;                  // small device-resident transforms
  void* pinned = nullptr;             // pinned host staging
  size_t pinned_cap = 0;
  // temporaries for the fused chains
  b2s_cloud* t0 = nullptr; b2s_cloud* t1 = nullptr; b2s_cloud* t2 = nullptr; b2s_cloud* t3 = nullptr;
  std::vector<b2s::GridIndex*> batch_grids;
  b2s::DevBuf batch_jobs;             // GridJob + ScanJob tables and the scan tile states of a batched index build
  std::vector<unsigned char> batch_jobs_host;
};

namespace b2s {

struct ProfScope {   // records an event pair around the launches issued during its lifetime (no-op unless enabled)
  b2s_handle* h; int idx;
  ProfScope(b2s_handle* h_, int kind);
  ~ProfScope();
};
int32_t ensure_pinned(b2s_handle* h, size_t bytes);
int32_t check_status(b2s_handle* h);     // synchronises and converts device status bits into an error

// ---- primitives (scan.cu / radix_sort.cu / grid_index.cu / ...) : all asynchronous on h->stream ----
// exclusive scan of in[0..*d_n) into out[0..*d_n]; out[*d_n] and *d_total (optional) receive the total
int32_t scan_exclusive_i32(b2s_handle* h, const int32_t* in, int32_t* out, const int32_t* d_n, size_t n_max, int32_t* d_total);
// njobs independent scans in one launch; every job's tile state (scan_state_bytes(n_max), zeroed) is supplied by the caller
size_t scan_state_bytes(size_t n_max);
int32_t scan_exclusive_i32_batch(b2s_handle* h, const ScanJob* jobs_dev, int njobs, size_t n_max);
// stable LSD radix sort of (key, value) pairs, key_bits low bits significant; result ends in keys/vals
// (pointers are swapped so that keys/vals designate the sorted arrays on return, *_alt the scratch)
int32_t radix_sort_pairs_u32(b2s_handle* h, uint32_t*& keys, uint32_t*& vals, uint32_t*& keys_alt, uint32_t*& vals_alt,
                             const int32_t* d_n, size_t n_max, int key_bits);
int32_t radix_sort_pairs_u64(b2s_handle* h, uint64_t*& keys, uint32_t*& vals, uint64_t*& keys_alt, uint32_t*& vals_alt,
                             const int32_t* d_n, size_t n_max, int key_bits);
inline const int32_t* grid_starts(const GridIndex* g) { return g->cell_start.as<int32_t>() + g->cap_cells + 4; }

// K-index: build the NN grid over cloud points (optionally only those inside `patch`, centre read from device pose)
struct CropDev {      // cropper passed by value to kernels; centre may come from a device-resident 4x4 (row-major)
  int32_t kind, invert;
  double rmin, rmax, zmin, zmax;
  double cx, cy, cz;
  const double* pose_dev;   // if non-null the centre is (pose[3], pose[7], pose[11])
};
CropDev make_crop(const b2s_cropper* c, const double* pose_dev = nullptr);
int32_t grid_build(b2s_handle* h, GridIndex* g, const b2s_cloud* cloud, double cell, const CropDev* patch, bool with_normals);
// the same for n clouds at once (batched registration: every pair brings its own target), blockIdx.y = cloud
int32_t grid_build_batch(b2s_handle* h, GridIndex* const* g, const b2s_cloud* const* clouds, int n, double cell, bool with_normals);

int32_t pose_to_device(b2s_handle* h, const double* T, double* dst);   // host 4x4 -> device slot, no staging buffer (voxel.cu)
int32_t cloud_reserve(b2s_handle* h, b2s_cloud* c, size_t n, bool normals);
int32_t cloud_set_count(b2s_handle* h, b2s_cloud* c, size_t n);

// stages
int32_t op_crop(b2s_handle* h, const b2s_cloud* in, const CropDev& crop, b2s_cloud* out);
int32_t op_voxel_down_sample(b2s_handle* h, const b2s_cloud* in, const CropDev* crop, double voxel, b2s_cloud* out);
// flags (optional, one int per point of c): only flagged points get a normal
int32_t op_estimate_normals(b2s_handle* h, b2s_cloud* c, int knn, double radius, double cell_hint, const int32_t* flags = nullptr);
int32_t select_flags(b2s_handle* h, const b2s_cloud* in, double ratio, uint32_t seed);
int32_t select_compact(b2s_handle* h, const b2s_cloud* in, double ratio, b2s_cloud* out);
int32_t op_random_down_sample(b2s_handle* h, const b2s_cloud* in, double ratio, uint32_t seed, b2s_cloud* out);
int32_t op_transform(b2s_handle* h, const b2s_cloud* in, const double* T_host, b2s_cloud* out);

struct IcpProblem {
  const double* src_xyz;
  const int32_t* src_n;
  const GridHeader* ghdr;
  const int32_t* cell_start;
  const double* tgt_pts;    // double4
  const double* tgt_nrm;    // double4
  double* work_xyz;         // global working copy of the source (used when it does not fit in shared memory) ...
  int32_t* work_prev;       // ... and its per-point search state
  const double* init_dev;   // optional device-resident init (overrides init)
  double init[16];
  double max_corr;
  double rel_fitness, rel_rmse;
  int32_t max_iter;
  int32_t src_n_max;
  int32_t estimator;        // B2S_REG_POINT_TO_PLANE / B2S_REG_POINT_TO_POINT / EST_INFORMATION
  int32_t pad;
  b2s_result* out;
  double* info_out;         // EST_INFORMATION: 36 doubles, row-major 6x6
  const double* src_nrm;    // B2S_REG_GENERALIZED: source normals (3 x f64 per point, source order)
  double gicp_eps;          // TransformationEstimationForGeneralizedICP::epsilon_ (1e-3)
  int32_t* corr_index;      // optional: per source point the ORIGINAL index of its final correspondence (-1 = none) ...
  double* corr_d2;          // ... and its squared distance (correspondence_set_ of the last evaluation)
};
constexpr int EST_INFORMATION = 3;   // internal estimator code: a single evaluation that outputs [O3D]'s information matrix
constexpr int EST_CORRESPONDENCES = 4;   // a single evaluation whose only output is corr_index / corr_d2 (+ fitness, rmse)
// single_host != nullptr: one registration, the problem travels as a kernel argument (no copy, no sync)
int32_t icp_launch(b2s_handle* h, const IcpProblem* single_host, const IcpProblem* problems_dev, int n_problems, size_t max_src_points);

int32_t op_submap_insert(b2s_handle* h, b2s_submap* sm, const b2s_cloud* scan, const double* T_dev, const int32_t* gate_dev);
// K-fuse bookkeeping (fuse.cu): (re)build the persistent voxel hash from the map cloud (after set_cloud / carving / transform;
// enable_dev gates it on the device), allocate it, and the tombstone-free view of the map for readers that leave the device
constexpr int FUSE_STAGE_BASE = 1 << 30;   // chain members >= this are staged scan points (index - FUSE_STAGE_BASE)
constexpr int FUSE_DUP_CAP = 1 << 16;
int32_t fuse_reserve(b2s_handle* h, b2s_submap* sm);
int32_t fuse_rehash(b2s_handle* h, b2s_submap* sm, const int32_t* enable_dev = nullptr);
int32_t submap_compact_view(b2s_handle* h, b2s_submap* sm, b2s_cloud** view);   // -> sm->cloud[1] holding the live points in map order
// F2 VoxelHashMap queries on the dense map (fuse.cu)
int32_t op_dense_query(b2s_handle* h, const b2s_submap* sm, const b2s_cloud* pts, int32_t* count_dev, double* mean_dev);
int32_t op_dense_remove(b2s_handle* h, b2s_submap* sm, const b2s_cloud* pts);
int32_t op_dense_count(b2s_handle* h, const b2s_submap* sm, int32_t* out_dev);
// sensor_dev != nullptr: the sensor position is the translation of that device-resident 4x4; enable_dev (optional): the
// kernels return at once unless *enable_dev != 0 (device-side schedule of the mapper chain)
int32_t op_dense_carve(b2s_handle* h, b2s_submap* sm, const b2s_cloud* scan, const double* sensor, const double* sensor_dev, double radius,
                       double trunc, double max_len, int32_t* removed_dev, const int32_t* enable_dev = nullptr);
int32_t op_submap_transform(b2s_handle* h, b2s_submap* sm, const double* T_host);   // Submap::transform (voxel.cu)
// D1 constant-velocity de-skew (voxel.cu)
int32_t op_undistort(b2s_handle* h, const b2s_cloud* in, const double* lin_vel, const double* ang_vel_rpy, double scan_duration, int clockwise,
                     b2s_cloud* out);
// L1 overlap selection in front of the loop-closure ICP (overlap.cu); T_dev = sourceToTarget (device, row-major)
int32_t op_overlap(b2s_handle* h, const b2s_cloud* source, const b2s_cloud* target, const double* T_dev, double voxel, int min_pts,
                   b2s_cloud* source_overlap, b2s_cloud* target_overlap);
// C1 space carving of the sparse map (carve.cu); removed_dev (optional) receives the number of removed points
int32_t op_submap_carve(b2s_handle* h, b2s_submap* sm, const b2s_cloud* raw_scan, const double* T_dev, const CropDev& crop,
                        const b2s_carving_params& prm, int32_t* removed_dev, const int32_t* enable_dev = nullptr);
// T_dev != nullptr: device-resident pose (T_host ignored)
int32_t op_dense_insert(b2s_handle* h, b2s_submap* sm, const b2s_cloud* raw, const double* T_host, const double* T_dev, const b2s_cropper* crop,
                        const int32_t* enable_dev = nullptr);

// upper bound of the CTAs of a streaming kernel: 2 per SM (B2S_GRID_CAP overrides).  Measured at 16 concurrent chains: 148 .. 592
// CTAs give 9.1 - 9.2 k registrations/s, 2368 (the round-1 value) 8.1 k -- few fat CTAs leave the SMs to the other chains' kernels
int grid_cap();
// The stand-alone operators (voxel down-sample, normals, crop of ONE large cloud: config 3) are not sharing the GPU with other chains'
// kernels: for the duration of such a call the cap is lifted so that a 2^20-point cloud fills every SM.
struct WideGridScope {
  explicit WideGridScope(size_t n);
  ~WideGridScope();
  bool on;
};
inline int grid_for(size_t n, int threads, int max_blocks = 0) {
  if (max_blocks <= 0) max_blocks = grid_cap();
  size_t b = (n + (size_t)threads - 1) / (size_t)threads;
  if (b < 1) b = 1;
  if (b > (size_t)max_blocks) b = (size_t)max_blocks;
  return (int)b;
}

bool pdl_enabled();   // runtime.cu: true inside a PdlScope unless B2S_PDL=0 (A/B)
// Launches carry the attribute only inside the per-scan mapper chain (b2s_mapper_step_*), where it was measured (+2.2 % at 16 chains);
// everywhere else kernels launch exactly as before.
struct PdlScope {
  PdlScope();
  ~PdlScope();
};

#ifdef __CUDACC__
// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-stream-serialization attribute (see pdl_wait)
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);   // errors surface through cudaGetLastError / the next B2S_CUDA, like <<<>>>
}
#endif

}  // namespace b2s

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace b2s {

// Programmatic dependent launch: every kernel of the library starts with this wait and every launch carries the programmatic-stream-
// serialization attribute, so a kernel's launch (scheduling, CTA distribution, its own prologue up to here) overlaps the tail of its
// predecessor in the stream instead of starting after it -- a scan is a chain of 42 kernels of 3-50 us.  Past the wait the predecessor
// grid has completed and its writes are visible, so nothing else changes.  Without the launch attribute the instruction is a no-op.
// (An explicit early griddepcontrol.launch_dependents at the top of every kernel was measured too: the successors' CTAs become resident
// long before they can run and hold registers / shared memory that the other chains' kernels need -- 10.4 k -> 7.9 k registrations/s.)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// order-preserving map double -> uint64 so that atomicMin/atomicMax work on doubles
__host__ __device__ inline unsigned long long ord_encode(double v) {
  unsigned long long u;
#ifdef __CUDA_ARCH__
  u = (unsigned long long)__double_as_longlong(v);
#else
  memcpy(&u, &v, 8);
#endif
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double ord_decode(unsigned long long u) {
  u = (u & 0x8000000000000000ull) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double v; memcpy(&v, &u, 8); return v;
#endif
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// squared distance accumulated exactly like nanoflann's L2 adaptor and the oracle: (dx*dx + dy*dy) + dz*dz,
// with explicit round-to-nearest ops so that nvcc never contracts it into FMAs (keeps argmin / strict radius
// decisions bit-identical to the CPU oracle)
__device__ __forceinline__ double dist2_exact(double ax, double ay, double az, double bx, double by, double bz) {
  double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// reference croppers (core/src/croppers.cpp:121-165); only the translation of the pose is used
__device__ __forceinline__ bool crop_within(const CropDev& c, double x, double y, double z) {
  double cx = c.cx, cy = c.cy, cz = c.cz;
  if (c.pose_dev) { cx = c.pose_dev[3]; cy = c.pose_dev[7]; cz = c.pose_dev[11]; }
  double dx = __dsub_rn(x, cx), dy = __dsub_rn(y, cy), dz = __dsub_rn(z, cz);
  bool w = true;
  switch (c.kind) {
    case B2S_CROP_MAX_RADIUS: w = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))) <= c.rmax; break;
    case B2S_CROP_MIN_RADIUS: w = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))) >= c.rmin; break;
    case B2S_CROP_MINMAX_RADIUS: {
      double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
      w = d <= c.rmax && d >= c.rmin;
      break;
    }
    case B2S_CROP_CYLINDER: w = z >= c.zmin && z <= c.zmax && sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) <= c.rmax; break;
    default: w = true;
  }
  return c.invert ? !w : w;
}

// 3 x 10-bit (u32) and 3 x 21-bit (u64) Morton interleave
__device__ __forceinline__ uint32_t morton_part10(uint32_t x) {
  x &= 0x3FFu;
  x = (x | (x << 16)) & 0x030000FFu;
  x = (x | (x << 8)) & 0x0300F00Fu;
  x = (x | (x << 4)) & 0x030C30C3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}
__device__ __forceinline__ uint64_t morton_part21(uint64_t x) {
  x &= 0x1FFFFFull;
  x = (x | (x << 32)) & 0x1F00000000FFFFull;
  x = (x | (x << 16)) & 0x1F0000FF0000FFull;
  x = (x | (x << 8)) & 0x100F00F00F00F00Full;
  x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

}  // namespace b2s
#endif
