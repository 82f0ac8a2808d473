// fuse.cu -- K-fuse (F0+F1) and K-dense (F3): the map side of the hot path.
//
//   F1  Submap::insertScan (no carving)        core/src/Submap.cpp:39-75
//         transform(T, scan)                   core/src/helpers.cpp:273-305   (duplication quirk when T ~ identity)
//         mapCloud_ += scan
//         voxelizeWithinCroppingVolume(...)    core/src/helpers.cpp:115-183   (+ AccumulatedPoint :30-70)
//   F3  Submap::insertScanDenseMap -> VoxelizedPointCloud::insert   core/src/Submap.cpp:77-92, core/src/Voxel.cpp:66-88
//
// F1 on the device keeps the reference's semantics exactly: the transformed scan is appended to the map arrays, every
// point inside the map-builder cropper (centred on the sensor) is keyed by floor(p * (1/v)) on the GLOBAL-origin grid,
// a stable radix sort groups voxel members in map order, one thread per voxel averages them in that order (an old
// map point counts as ONE member; normals: mean of non-NaN then normalized()), points outside the cropper pass
// through untouched.  The result is committed back into the map arrays only when the (device-resident) gate is open,
// which is how the fitness gate of Mapper::addRangeMeasurement (core/src/Mapper.cpp:151) runs without a host sync.
// Output order: voxels in Morton order of (key - base), then pass-through points (the reference: pass-through first,
// then std::unordered_map order -- unspecified, nothing downstream depends on it).
#include "common.cuh"

namespace b2s {

constexpr int FZ_THREADS = 256;

int32_t pose_to_device(b2s_handle* h, const double* T, double* dst);  // voxel.cu

// append (optionally duplicated) transformed scan to the map arrays; writes the total into *d_tot
__global__ void __launch_bounds__(FZ_THREADS) fuse_append_kernel(double* __restrict__ mxyz, double* __restrict__ mnrm,
                                                                 const int32_t* __restrict__ d_nmap, const double* __restrict__ sxyz,
                                                                 const double* __restrict__ snrm, const int32_t* __restrict__ d_nscan,
                                                                 const double* __restrict__ Tdev, const int32_t* __restrict__ gate,
                                                                 size_t capacity, int32_t* d_tot, uint32_t* status) {
  const int nmap = *d_nmap;
  int ns = *d_nscan;
  const bool open = (gate == nullptr || *gate != 0) && ns > 0;  // Submap.cpp:41-43: empty scan -> nothing happens
  if (!open) { if (blockIdx.x == 0 && threadIdx.x == 0) *d_tot = 0; return; }
  double T[16];
#pragma unroll
  for (int i = 0; i < 16; i++) T[i] = Tdev[i];
  double mx = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) mx = fmax(mx, fabs(T[i] - ((i % 5 == 0) ? 1.0 : 0.0)));
  const bool ident = mx < 1e-4;  // helpers.cpp:275
  const size_t total = (size_t)nmap + (size_t)(ident ? 2 : 1) * (size_t)ns;
  if (total > capacity) { if (blockIdx.x == 0 && threadIdx.x == 0) { atomicOr(status, ST_CAPACITY); *d_tot = 0; } return; }
  if (blockIdx.x == 0 && threadIdx.x == 0) *d_tot = (int32_t)total;
  const size_t b0 = (size_t)nmap, b1 = b0 + (ident ? (size_t)ns : 0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
    const double px = sxyz[3 * i], py = sxyz[3 * i + 1], pz = sxyz[3 * i + 2];
    const double a = snrm[3 * i], b = snrm[3 * i + 1], c = snrm[3 * i + 2];
    if (ident) {
      mxyz[3 * (b0 + i)] = px; mxyz[3 * (b0 + i) + 1] = py; mxyz[3 * (b0 + i) + 2] = pz;
      mnrm[3 * (b0 + i)] = a; mnrm[3 * (b0 + i) + 1] = b; mnrm[3 * (b0 + i) + 2] = c;
    }
    const double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], px), __dmul_rn(T[1], py)), __dmul_rn(T[2], pz)), T[3]);
    const double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], px), __dmul_rn(T[5], py)), __dmul_rn(T[6], pz)), T[7]);
    const double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], px), __dmul_rn(T[9], py)), __dmul_rn(T[10], pz)), T[11]);
    const double w = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[12], px), __dmul_rn(T[13], py)), __dmul_rn(T[14], pz)), T[15]);
    const size_t o = b1 + i;
    mxyz[3 * o] = __ddiv_rn(x, w); mxyz[3 * o + 1] = __ddiv_rn(y, w); mxyz[3 * o + 2] = __ddiv_rn(z, w);
    mnrm[3 * o] = __dadd_rn(__dadd_rn(__dmul_rn(T[0], a), __dmul_rn(T[1], b)), __dmul_rn(T[2], c));
    mnrm[3 * o + 1] = __dadd_rn(__dadd_rn(__dmul_rn(T[4], a), __dmul_rn(T[5], b)), __dmul_rn(T[6], c));
    mnrm[3 * o + 2] = __dadd_rn(__dadd_rn(__dmul_rn(T[8], a), __dmul_rn(T[9], b)), __dmul_rn(T[10], c));
  }
}

template <typename K>
__device__ __forceinline__ K morton3f(uint32_t x, uint32_t y, uint32_t z);
template <>
__device__ __forceinline__ uint32_t morton3f<uint32_t>(uint32_t x, uint32_t y, uint32_t z) {
  return morton_part10(x) | (morton_part10(y) << 1) | (morton_part10(z) << 2);
}
template <>
__device__ __forceinline__ uint64_t morton3f<uint64_t>(uint32_t x, uint32_t y, uint32_t z) {
  return morton_part21(x) | (morton_part21(y) << 1) | (morton_part21(z) << 2);
}

// key = Morton(floor(p * inv) - base) for points inside the cropper, sentinel otherwise.
// base: bounded cropper -> floor((centre - rmax) * inv) - 1 per axis; unbounded -> -2^20 (21-bit keys).
template <typename K>
__global__ void __launch_bounds__(FZ_THREADS) fuse_keys_kernel(const double* __restrict__ mxyz, const int32_t* __restrict__ d_tot,
                                                               CropDev crop, double inv, int bits, int bounded,
                                                               K* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* status) {
  const int n = *d_tot;
  const K invalid = (K)1 << (3 * bits);
  double bx, by, bz;
  if (bounded) {
    double cx = crop.cx, cy = crop.cy, cz = crop.cz;
    if (crop.pose_dev) { cx = crop.pose_dev[3]; cy = crop.pose_dev[7]; cz = crop.pose_dev[11]; }
    bx = floor((cx - crop.rmax) * inv) - 1.0; by = floor((cy - crop.rmax) * inv) - 1.0; bz = floor((cz - crop.rmax) * inv) - 1.0;
  } else { bx = by = bz = -1048576.0; }
  const double lim = (double)(1u << bits);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double x = mxyz[3 * i], y = mxyz[3 * i + 1], z = mxyz[3 * i + 2];
    K key = invalid;
    if (crop_within(crop, x, y, z)) {
      // getVoxelIdx(p, invVoxelSize): int(floor(p * invSize))   VoxelHashMap.hpp:47-50
      const double fx = floor(__dmul_rn(x, inv)) - bx, fy = floor(__dmul_rn(y, inv)) - by, fz = floor(__dmul_rn(z, inv)) - bz;
      if (fx >= 0.0 && fy >= 0.0 && fz >= 0.0 && fx < lim && fy < lim && fz < lim) key = morton3f<K>((uint32_t)fx, (uint32_t)fy, (uint32_t)fz);
      else atomicOr(status, ST_KEY_OVERFLOW);
    }
    keys[i] = key;
    vals[i] = (uint32_t)i;
  }
}

template <typename K>
__global__ void __launch_bounds__(FZ_THREADS) fuse_head_kernel(const K* __restrict__ keys, const int32_t* __restrict__ d_tot, int bits,
                                                               int32_t* __restrict__ head) {
  const int n = *d_tot;
  const K invalid = (K)1 << (3 * bits);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const K k = keys[j];
    head[j] = (k >= invalid) ? 1 : ((j == 0 || keys[j - 1] != k) ? 1 : 0);  // pass-through points are singleton segments
  }
}

template <typename K>
__global__ void __launch_bounds__(FZ_THREADS) fuse_mean_kernel(const K* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                               const int32_t* __restrict__ d_tot, int bits,
                                                               const int32_t* __restrict__ head, const int32_t* __restrict__ offs,
                                                               const double* __restrict__ mxyz, const double* __restrict__ mnrm,
                                                               double* __restrict__ oxyz, double* __restrict__ onrm, int32_t* d_out_n) {
  const int n = *d_tot;
  const K invalid = (K)1 << (3 * bits);
  if (blockIdx.x == 0 && threadIdx.x == 0) *d_out_n = n > 0 ? offs[n] : 0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    if (!head[j]) continue;
    const K k = keys[j];
    const int o = offs[j];
    if (k >= invalid) {  // outside the cropper: copied through unchanged (helpers.cpp:156-166)
      const uint32_t i = vals[j];
      oxyz[3 * o] = mxyz[3 * i]; oxyz[3 * o + 1] = mxyz[3 * i + 1]; oxyz[3 * o + 2] = mxyz[3 * i + 2];
      onrm[3 * o] = mnrm[3 * i]; onrm[3 * o + 1] = mnrm[3 * i + 1]; onrm[3 * o + 2] = mnrm[3 * i + 2];
      continue;
    }
    double sx = 0, sy = 0, sz = 0, nx = 0, ny = 0, nz = 0;
    int cnt = 0;
    for (int t = j; t < n && keys[t] == k; ++t) {
      const uint32_t i = vals[t];
      sx = __dadd_rn(sx, mxyz[3 * i]); sy = __dadd_rn(sy, mxyz[3 * i + 1]); sz = __dadd_rn(sz, mxyz[3 * i + 2]);
      const double a = mnrm[3 * i], b = mnrm[3 * i + 1], c = mnrm[3 * i + 2];
      if (a == a && b == b && c == c) { nx = __dadd_rn(nx, a); ny = __dadd_rn(ny, b); nz = __dadd_rn(nz, c); }
      cnt++;
    }
    const double c = (double)cnt;
    oxyz[3 * o] = __ddiv_rn(sx, c); oxyz[3 * o + 1] = __ddiv_rn(sy, c); oxyz[3 * o + 2] = __ddiv_rn(sz, c);
    double a0 = __ddiv_rn(nx, c), a1 = __ddiv_rn(ny, c), a2 = __ddiv_rn(nz, c);
    const double zz = __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(a2, a2));
    if (zz > 0.0) { const double sn = sqrt(zz); a0 = __ddiv_rn(a0, sn); a1 = __ddiv_rn(a1, sn); a2 = __ddiv_rn(a2, sn); }  // .normalized()
    onrm[3 * o] = a0; onrm[3 * o + 1] = a1; onrm[3 * o + 2] = a2;
  }
}

// mstate / last_pose: Submap::nScansInsertedMap_ and the pose mapBuilderCropper_ was last set to (Submap.cpp:71,73), which is
// also Mapper::mapToRangeSensorLastScanInsertion_ (Mapper.cpp:175) -- kept on the device for the chain's own gates
__global__ void __launch_bounds__(FZ_THREADS) fuse_commit_kernel(const double* __restrict__ oxyz, const double* __restrict__ onrm,
                                                                 const int32_t* __restrict__ d_out_n, const int32_t* __restrict__ d_tot,
                                                                 double* __restrict__ mxyz, double* __restrict__ mnrm, int32_t* d_nmap,
                                                                 int32_t* mstate, double* last_pose, const double* __restrict__ Tdev) {
  if (*d_tot <= 0) return;  // gate closed, empty scan or capacity error: the map stays as it was
  const int n = *d_out_n;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *d_nmap = n;
    if (mstate) mstate[MS_NINS] += 1;
    if (last_pose) for (int i = 0; i < 16; i++) last_pose[i] = Tdev[i];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n; i += gridDim.x * blockDim.x) { mxyz[i] = oxyz[i]; mnrm[i] = onrm[i]; }
}

static int bits_for_range(double cells) {
  int b = 1;
  while ((double)(1u << b) < cells && b < 22) b++;
  return b;
}

template <typename K>
static int32_t fuse_impl(b2s_handle* h, b2s_submap* sm, const b2s_cloud* scan, const double* T_dev, const int32_t* gate_dev,
                         const CropDev& crop, double inv, int bits, int bounded, size_t tot_max) {
  b2s_cloud* map = sm->cloud[0];
  b2s_cloud* tmp = sm->cloud[1];
  // scratch is sized once for the submap capacity: growing a device buffer costs a cudaMalloc/cudaFree, i.e. a
  // device-wide synchronisation that would stall every other chain sharing the GPU
  const size_t cap = sm->capacity;
  B2S_TRY(h->keys.ensure(cap * sizeof(K) * 2, h->stream));
  B2S_TRY(h->vals.ensure(cap * 4 * 2, h->stream));
  B2S_TRY(h->flags.ensure((cap + 1) * 4, h->stream));
  B2S_TRY(h->offs.ensure((cap + 2) * 4, h->stream));
  B2S_TRY(h->tmp_i32.ensure(64, h->stream));
  int32_t* d_tot = h->tmp_i32.as<int32_t>();
  int32_t* d_out_n = d_tot + 1;
  K* keys = h->keys.as<K>(); K* keys_alt = keys + cap;
  uint32_t* vals = h->vals.as<uint32_t>(); uint32_t* vals_alt = vals + cap;
  const int sblocks = grid_for(scan->n_max > 0 ? scan->n_max : 1, FZ_THREADS);
  const int blocks = grid_for(tot_max, FZ_THREADS);
  { ProfScope prof(h, PK_FUSE);
  fuse_append_kernel<<<sblocks, FZ_THREADS, 0, h->stream>>>(map->xyz.as<double>(), map->nrm.as<double>(), map->dn.as<int32_t>(),
                                                            scan->xyz.as<double>(), scan->nrm.as<double>(), scan->dn.as<int32_t>(), T_dev,
                                                            gate_dev, sm->capacity, d_tot, h->status.as<uint32_t>());
  fuse_keys_kernel<K><<<blocks, FZ_THREADS, 0, h->stream>>>(map->xyz.as<double>(), d_tot, crop, inv, bits, bounded, keys, vals,
                                                            h->status.as<uint32_t>());
  h->launches += 2; }
  if constexpr (sizeof(K) == 4) {
    B2S_TRY(radix_sort_pairs_u32(h, keys, vals, keys_alt, vals_alt, d_tot, tot_max, 3 * bits + 1));
  } else {
    B2S_TRY(radix_sort_pairs_u64(h, keys, vals, keys_alt, vals_alt, d_tot, tot_max, 3 * bits + 1));
  }
  ProfScope prof2(h, PK_FUSE);
  fuse_head_kernel<K><<<blocks, FZ_THREADS, 0, h->stream>>>(keys, d_tot, bits, h->flags.as<int32_t>());
  h->launches++;
  B2S_TRY(scan_exclusive_i32(h, h->flags.as<int32_t>(), h->offs.as<int32_t>(), d_tot, tot_max, nullptr));
  fuse_mean_kernel<K><<<blocks, FZ_THREADS, 0, h->stream>>>(keys, vals, d_tot, bits, h->flags.as<int32_t>(), h->offs.as<int32_t>(),
                                                            map->xyz.as<double>(), map->nrm.as<double>(), tmp->xyz.as<double>(),
                                                            tmp->nrm.as<double>(), d_out_n);
  fuse_commit_kernel<<<blocks, FZ_THREADS, 0, h->stream>>>(tmp->xyz.as<double>(), tmp->nrm.as<double>(), d_out_n, d_tot,
                                                           map->xyz.as<double>(), map->nrm.as<double>(), map->dn.as<int32_t>(),
                                                           sm->mstate.as<int32_t>(), sm->pose.as<double>() + 5 * 16, T_dev);
  h->launches += 2;
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

int32_t op_submap_insert(b2s_handle* h, b2s_submap* sm, const b2s_cloud* scan, const double* T_dev, const int32_t* gate_dev) {
  B2S_REQUIRE(scan->has_normals, B2S_E_NO_NORMALS, "Submap::insertScan: the pre-processed scan must carry normals");
  b2s_cloud* map = sm->cloud[0];
  const double v = h->cfg.map_voxel_size;
  B2S_REQUIRE(v > 0.0, B2S_E_UNSUPPORTED, "map_voxel_size <= 0 (no voxelisation) is not supported on the device path");
  // host-side upper bound of the map size.  The exact size is read back asynchronously after an insertion (pinned
  // host word + event); once that copy has landed the bound becomes exact-size + what was appended since.
  if (sm->cnt_pending && cudaEventQuery(sm->cnt_ev) == cudaSuccess) {
    map->n_max = (size_t)sm->pinned_cnt[0] + sm->adds_after_readback;
    sm->cnt_pending = false;
  }
  size_t tot_max = map->n_max + 2 * scan->n_max;
  if (sm->graph_mode) tot_max = sm->capacity;   // graph replay: constant launch dimensions, overflow is caught on the device
  if (tot_max > sm->capacity) {
    int32_t n = 0;
    B2S_CUDA(cudaMemcpyAsync(&n, map->dn.p, 4, cudaMemcpyDeviceToHost, h->stream));
    B2S_CUDA(cudaStreamSynchronize(h->stream));
    map->n_max = (size_t)n; map->n_known = n;
    tot_max = map->n_max + 2 * scan->n_max;
    B2S_REQUIRE(tot_max <= sm->capacity, B2S_E_CAPACITY, "submap capacity %zu too small for %zu points", sm->capacity, tot_max);
  }
  CropDev crop = make_crop(&h->cfg.scan.map_builder_cropper, T_dev);  // Submap.cpp:71 setPose(mapToRangeSensor)
  const double inv = 1.0 / v;
  const int bounded = (!crop.invert && (crop.kind == B2S_CROP_MAX_RADIUS || crop.kind == B2S_CROP_MINMAX_RADIUS)) ? 1 : 0;
  const int bits = bounded ? bits_for_range(floor(2.0 * crop.rmax * inv) + 4.0) : 21;
  B2S_REQUIRE(bits <= 21, B2S_E_INVALID, "map voxel size too small for the cropper radius");
  int32_t rc = (bits <= 10) ? fuse_impl<uint32_t>(h, sm, scan, T_dev, gate_dev, crop, inv, bits, bounded, tot_max)
                            : fuse_impl<uint64_t>(h, sm, scan, T_dev, gate_dev, crop, inv, bits, bounded, tot_max);
  map->n_max = tot_max;   // upper bound only; the exact count lives on the device
  map->n_known = -1;
  map->has_normals = true;
  if (rc == B2S_OK && !sm->graph_mode) {
    if (!sm->pinned_cnt) {
      B2S_CUDA(cudaMallocHost(&sm->pinned_cnt, 64));
      B2S_CUDA(cudaEventCreateWithFlags(&sm->cnt_ev, cudaEventDisableTiming));
    }
    if (!sm->cnt_pending) {
      B2S_CUDA(cudaMemcpyAsync(sm->pinned_cnt, map->dn.p, 4, cudaMemcpyDeviceToHost, h->stream));
      B2S_CUDA(cudaEventRecord(sm->cnt_ev, h->stream));
      sm->cnt_pending = true;
      sm->adds_after_readback = 0;
    } else {
      sm->adds_after_readback += 2 * scan->n_max;
    }
  }
  return rc;
}

// =====================================================================================================================
//  F3 dense map: open-addressing hash (64-bit packed key) of running position / normal sums and counts
// =====================================================================================================================
constexpr unsigned long long DENSE_EMPTY = ~0ull;

__device__ __forceinline__ unsigned long long dense_pack(int x, int y, int z) {
  return ((unsigned long long)(unsigned)(x + 1048576) << 42) | ((unsigned long long)(unsigned)(y + 1048576) << 21) |
         (unsigned long long)(unsigned)(z + 1048576);
}
__device__ __forceinline__ unsigned long long dense_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

__device__ __forceinline__ void dense_add_point(double x, double y, double z, double inv, unsigned long long* __restrict__ keys,
                                                double* __restrict__ sums, int32_t* __restrict__ cnts, size_t cap, int32_t* used, uint32_t* status) {
  const double fx = floor(__dmul_rn(x, inv)), fy = floor(__dmul_rn(y, inv)), fz = floor(__dmul_rn(z, inv));
  if (!(fabs(fx) < 1048575.0 && fabs(fy) < 1048575.0 && fabs(fz) < 1048575.0)) { atomicOr(status, ST_KEY_OVERFLOW); return; }
  const unsigned long long key = dense_pack((int)fx, (int)fy, (int)fz);
  size_t slot = (size_t)(dense_hash(key) % cap);
  for (size_t probe = 0; probe < cap; ++probe) {
    unsigned long long prev = atomicCAS(&keys[slot], DENSE_EMPTY, key);
    if (prev == DENSE_EMPTY) {
      if ((size_t)atomicAdd(used, 1) + 1 > cap - cap / 8) atomicOr(status, ST_HASH_FULL);
      prev = key;
    }
    if (prev == key) {
      atomicAdd(&sums[6 * slot], x); atomicAdd(&sums[6 * slot + 1], y); atomicAdd(&sums[6 * slot + 2], z);
      atomicAdd(&cnts[slot], 1);
      return;
    }
    slot = slot + 1 == cap ? 0 : slot + 1;
  }
}

// Submap::insertScanDenseMap goes through o3d_slam::transform (Submap.cpp:80), so its near-identity duplication quirk
// (helpers.cpp:275-292: |T - I|_max < 1e-4 -> the untransformed cloud is copied first and every transformed point is
// appended as well) applies: such a scan lands in the dense map twice.  Kept.
__global__ void __launch_bounds__(FZ_THREADS) dense_insert_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ d_n,
                                                                  const double* __restrict__ Tdev, CropDev crop, double inv,
                                                                  unsigned long long* __restrict__ keys, double* __restrict__ sums,
                                                                  int32_t* __restrict__ cnts, size_t cap, int32_t* used, uint32_t* status,
                                                                  const int32_t* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;
  const int n = *d_n;
  double T[16];
#pragma unroll
  for (int i = 0; i < 16; i++) T[i] = Tdev[i];
  double mx = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) mx = fmax(mx, fabs(T[i] - ((i % 5 == 0) ? 1.0 : 0.0)));
  const bool ident = mx < 1e-4;  // helpers.cpp:275
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    if (!(px == px && py == py && pz == pz)) continue;
    if (!crop_within(crop, px, py, pz)) continue;  // denseMapCropper_ at identity, applied in the sensor frame (Submap.cpp:78-79)
    if (ident) dense_add_point(px, py, pz, inv, keys, sums, cnts, cap, used, status);
    const double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], px), __dmul_rn(T[1], py)), __dmul_rn(T[2], pz)), T[3]);
    const double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], px), __dmul_rn(T[5], py)), __dmul_rn(T[6], pz)), T[7]);
    const double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], px), __dmul_rn(T[9], py)), __dmul_rn(T[10], pz)), T[11]);
    const double w = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[12], px), __dmul_rn(T[13], py)), __dmul_rn(T[14], pz)), T[15]);
    dense_add_point(__ddiv_rn(x, w), __ddiv_rn(y, w), __ddiv_rn(z, w), inv, keys, sums, cnts, cap, used, status);
  }
}

__global__ void dense_init_kernel(unsigned long long* keys, double* sums, int32_t* cnts, size_t cap, int32_t* used) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
    keys[i] = DENSE_EMPTY; cnts[i] = 0;
    for (int k = 0; k < 6; k++) sums[6 * i + k] = 0.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *used = 0;
}

int32_t dense_init(b2s_handle* h, b2s_submap* sm, size_t cap, double voxel) {
  B2S_TRY(sm->dense_keys.ensure(cap * 8, h->stream));
  B2S_TRY(sm->dense_sum.ensure(cap * 48, h->stream));
  B2S_TRY(sm->dense_cnt.ensure(cap * 4, h->stream));
  B2S_TRY(sm->dense_used.ensure(16, h->stream));
  sm->dense_cap = cap; sm->dense_voxel = voxel;
  dense_init_kernel<<<148 * 4, 256, 0, h->stream>>>(sm->dense_keys.as<unsigned long long>(), sm->dense_sum.as<double>(),
                                                    sm->dense_cnt.as<int32_t>(), cap, sm->dense_used.as<int32_t>());
  h->launches++;
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

int32_t op_dense_insert(b2s_handle* h, b2s_submap* sm, const b2s_cloud* raw, const double* T_host, const double* T_dev, const b2s_cropper* crop,
                        const int32_t* enable_dev) {
  B2S_REQUIRE(sm->dense_cap > 0, B2S_E_INVALID, "dense map not initialised");
  const double* Td = T_dev;
  if (!Td) {
    B2S_TRY(h->poses.ensure(64 * 16 * 8, h->stream, true));
    double* slot = h->poses.as<double>() + 16 * 62;
    B2S_TRY(pose_to_device(h, T_host, slot));
    Td = slot;
  }
  b2s_cropper c0;
  memset(&c0, 0, sizeof(c0));
  if (crop) c0 = *crop;
  c0.center[0] = c0.center[1] = c0.center[2] = 0.0;  // Submap.cpp:78 setPose(Identity)
  dense_insert_kernel<<<grid_for(raw->n_max > 0 ? raw->n_max : 1, FZ_THREADS), FZ_THREADS, 0, h->stream>>>(
      raw->xyz.as<double>(), raw->dn.as<int32_t>(), Td, make_crop(&c0), 1.0 / sm->dense_voxel, sm->dense_keys.as<unsigned long long>(),
      sm->dense_sum.as<double>(), sm->dense_cnt.as<int32_t>(), sm->dense_cap, sm->dense_used.as<int32_t>(), h->status.as<uint32_t>(), enable_dev);
  h->launches++;
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

// VoxelizedPointCloud::toPointCloud (Voxel.cpp:90-115): flags -> scan -> gather of sum / count
__global__ void dense_flags_kernel(const int32_t* __restrict__ cnts, size_t cap, int32_t* __restrict__ flags) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) flags[i] = cnts[i] > 0 ? 1 : 0;
}
__global__ void dense_gather_kernel(const unsigned long long* __restrict__ keys, const double* __restrict__ sums,
                                    const int32_t* __restrict__ cnts, size_t cap, const int32_t* __restrict__ flags,
                                    const int32_t* __restrict__ offs, double* __restrict__ oxyz, int32_t* __restrict__ okeys, int32_t* out_n) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *out_n = offs[cap];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
    if (!flags[i]) continue;
    const int o = offs[i];
    const double c = (double)cnts[i];
    oxyz[3 * o] = sums[6 * i] / c; oxyz[3 * o + 1] = sums[6 * i + 1] / c; oxyz[3 * o + 2] = sums[6 * i + 2] / c;
    const unsigned long long k = keys[i];
    okeys[3 * o] = (int)((k >> 42) & 0x1FFFFF) - 1048576; okeys[3 * o + 1] = (int)((k >> 21) & 0x1FFFFF) - 1048576;
    okeys[3 * o + 2] = (int)(k & 0x1FFFFF) - 1048576;
  }
}

int32_t dense_to_cloud(b2s_handle* h, b2s_submap* sm, double* d_xyz, int32_t* d_keys, int32_t* d_out_n) {
  const size_t cap = sm->dense_cap;
  B2S_TRY(h->flags.ensure((cap + 1) * 4, h->stream));
  B2S_TRY(h->offs.ensure((cap + 2) * 4, h->stream));
  dense_flags_kernel<<<148 * 4, 256, 0, h->stream>>>(sm->dense_cnt.as<int32_t>(), cap, h->flags.as<int32_t>());
  h->launches++;
  B2S_TRY(scan_exclusive_i32(h, h->flags.as<int32_t>(), h->offs.as<int32_t>(), nullptr, cap, nullptr));
  dense_gather_kernel<<<148 * 4, 256, 0, h->stream>>>(sm->dense_keys.as<unsigned long long>(), sm->dense_sum.as<double>(),
                                                      sm->dense_cnt.as<int32_t>(), cap, h->flags.as<int32_t>(), h->offs.as<int32_t>(), d_xyz,
                                                      d_keys, d_out_n);
  h->launches++;
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

// =====================================================================================================================
//  F2  VoxelHashMap query API on the dense map (core/include/open3d_slam/VoxelHashMap.hpp:104-158), batched:
//      hasVoxelContainingPoint / getVoxelContainingPointPtr (-> aggregated content), removeKey(getKey(p)), size, clear.
//  A removed voxel keeps its key in the table with count 0 (= absent for every reader); inserting into it again simply
//  re-populates the slot, so no tombstone handling is needed.
// =====================================================================================================================
__device__ __forceinline__ long long dense_find(const unsigned long long* __restrict__ keys, size_t cap, double x, double y, double z, double inv) {
  const double fx = floor(__dmul_rn(x, inv)), fy = floor(__dmul_rn(y, inv)), fz = floor(__dmul_rn(z, inv));   // getVoxelIdx(p, inverseVoxelSize_)
  if (!(fabs(fx) < 1048575.0 && fabs(fy) < 1048575.0 && fabs(fz) < 1048575.0)) return -1;
  const unsigned long long key = dense_pack((int)fx, (int)fy, (int)fz);
  size_t slot = (size_t)(dense_hash(key) % cap);
  for (size_t probe = 0; probe < cap; ++probe) {
    const unsigned long long k = keys[slot];
    if (k == DENSE_EMPTY) return -1;
    if (k == key) return (long long)slot;
    slot = slot + 1 == cap ? 0 : slot + 1;
  }
  return -1;
}

__global__ void __launch_bounds__(FZ_THREADS) dense_query_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ d_n, double inv,
                                                                 const unsigned long long* __restrict__ keys, const double* __restrict__ sums,
                                                                 const int32_t* __restrict__ cnts, size_t cap, int32_t* __restrict__ count_out,
                                                                 double* __restrict__ mean_out) {
  const int n = *d_n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const long long s = dense_find(keys, cap, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], inv);
    const int c = s >= 0 ? cnts[s] : 0;
    count_out[i] = c;
    if (mean_out) {
      const double cd = (double)c;
      for (int k = 0; k < 3; k++) mean_out[3 * i + k] = c > 0 ? sums[6 * s + k] / cd : 0.0;   // AggregatedVoxel::getAggregatedPosition
    }
  }
}

__global__ void __launch_bounds__(FZ_THREADS) dense_remove_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ d_n, double inv,
                                                                  const unsigned long long* __restrict__ keys, double* __restrict__ sums,
                                                                  int32_t* __restrict__ cnts, size_t cap) {
  const int n = *d_n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const long long s = dense_find(keys, cap, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], inv);
    if (s < 0) continue;
    cnts[s] = 0;   // several points of the same voxel write the same zeros
    for (int k = 0; k < 6; k++) sums[6 * s + k] = 0.0;
  }
}

__global__ void dense_count_kernel(const int32_t* __restrict__ cnts, size_t cap, int32_t* out) {
  int c = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) c += cnts[i] > 0;
  c = warp_sum_i(c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

int32_t op_dense_query(b2s_handle* h, const b2s_submap* sm, const b2s_cloud* pts, int32_t* count_dev, double* mean_dev) {
  B2S_REQUIRE(sm->dense_cap > 0, B2S_E_INVALID, "dense map not initialised");
  dense_query_kernel<<<grid_for(pts->n_max > 0 ? pts->n_max : 1, FZ_THREADS), FZ_THREADS, 0, h->stream>>>(
      pts->xyz.as<double>(), pts->dn.as<int32_t>(), 1.0 / sm->dense_voxel, sm->dense_keys.as<unsigned long long>(), sm->dense_sum.as<double>(),
      sm->dense_cnt.as<int32_t>(), sm->dense_cap, count_dev, mean_dev);
  h->launches++;
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

int32_t op_dense_remove(b2s_handle* h, b2s_submap* sm, const b2s_cloud* pts) {
  B2S_REQUIRE(sm->dense_cap > 0, B2S_E_INVALID, "dense map not initialised");
  dense_remove_kernel<<<grid_for(pts->n_max > 0 ? pts->n_max : 1, FZ_THREADS), FZ_THREADS, 0, h->stream>>>(
      pts->xyz.as<double>(), pts->dn.as<int32_t>(), 1.0 / sm->dense_voxel, sm->dense_keys.as<unsigned long long>(), sm->dense_sum.as<double>(),
      sm->dense_cnt.as<int32_t>(), sm->dense_cap);
  h->launches++;
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

int32_t op_dense_count(b2s_handle* h, const b2s_submap* sm, int32_t* out_dev) {
  B2S_CUDA(cudaMemsetAsync(out_dev, 0, 4, h->stream));
  if (sm->dense_cap == 0) return B2S_OK;
  dense_count_kernel<<<148 * 4, 256, 0, h->stream>>>(sm->dense_cnt.as<int32_t>(), sm->dense_cap, out_dev);
  h->launches++;
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

// =====================================================================================================================
//  C2  space carving of the dense map: Submap::carve(scan, sensorPosition, param, VoxelizedPointCloud*)
//      core/src/Submap.cpp:125-136 -> removeDuplicatePointsWithinSameVoxels (core/src/Voxel.cpp:162-192),
//      getKeysOfCarvedPoints (core/src/helpers.cpp:347-377), getVoxelsWithinPointNeighborhood (core/src/VoxelHashMap.cpp:13-45)
//  (1) first point of every voxel of the scan = the ray set (atomicMin of the index per voxel of a scratch hash);
//  (2) one thread per ray: steps of 2*radius, at every step the reference's dx/dy/dz loops (floating accumulation kept as
//      written) enumerate test points; a test point within `radius` of its voxel centre nominates that voxel; nominated
//      voxels that exist in the dense map are flagged; (3) flagged voxels are emptied (removeKey).
// =====================================================================================================================
__device__ __forceinline__ long long dense_find_key(const unsigned long long* __restrict__ keys, size_t cap, int kx, int ky, int kz) {
  if (!(abs(kx) < 1048575 && abs(ky) < 1048575 && abs(kz) < 1048575)) return -1;
  const unsigned long long key = dense_pack(kx, ky, kz);
  size_t slot = (size_t)(dense_hash(key) % cap);
  for (size_t probe = 0; probe < cap; ++probe) {
    const unsigned long long k = keys[slot];
    if (k == DENSE_EMPTY) return -1;
    if (k == key) return (long long)slot;
    slot = slot + 1 == cap ? 0 : slot + 1;
  }
  return -1;
}

__global__ void __launch_bounds__(FZ_THREADS) dcarve_first_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ d_n, double inv,
                                                                  unsigned long long* keys, int32_t* first, size_t mask,
                                                                  int32_t* __restrict__ slot_of, const int32_t* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;
  const int n = *d_n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double fx = floor(__dmul_rn(xyz[3 * i], inv)), fy = floor(__dmul_rn(xyz[3 * i + 1], inv)), fz = floor(__dmul_rn(xyz[3 * i + 2], inv));
    slot_of[i] = -1;
    if (!(fabs(fx) < 1048575.0 && fabs(fy) < 1048575.0 && fabs(fz) < 1048575.0)) continue;   // NaN / far away: never a ray
    const unsigned long long key = dense_pack((int)fx, (int)fy, (int)fz);
    size_t s = (size_t)dense_hash(key) & mask;
    for (size_t probe = 0; probe <= mask; ++probe, s = (s + 1) & mask) {
      const unsigned long long old = atomicCAS(&keys[s], DENSE_EMPTY, key);
      if (old == DENSE_EMPTY || old == key) { atomicMin(&first[s], i); slot_of[i] = (int32_t)s; break; }
    }
  }
}

__global__ void dcarve_init_kernel(unsigned long long* keys, int32_t* first, size_t cap, int32_t* rm, size_t dense_cap,
                                   const int32_t* __restrict__ enable, int32_t* removed) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && removed) *removed = 0;
  if (enable != nullptr && *enable == 0) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) { keys[i] = DENSE_EMPTY; first[i] = 0x7fffffff; }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < dense_cap; i += (size_t)gridDim.x * blockDim.x) rm[i] = 0;
}

__global__ void __launch_bounds__(FZ_THREADS) dcarve_march_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ d_n,
                                                                  const int32_t* __restrict__ slot_of, const int32_t* __restrict__ first,
                                                                  double sx, double sy, double sz, const double* __restrict__ sensor_dev,
                                                                  double voxel, double radius, double trunc,
                                                                  double max_len, const unsigned long long* __restrict__ dkeys,
                                                                  const int32_t* __restrict__ dcnt, size_t dcap, int32_t* __restrict__ rm,
                                                                  const int32_t* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;
  if (sensor_dev) { sx = sensor_dev[3]; sy = sensor_dev[7]; sz = sensor_dev[11]; }   // mapToRangeSensor.translation()
  const int n = *d_n;
  const double step = 2.0 * radius;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int so = slot_of[i];
    if (so < 0 || first[so] != i) continue;       // removeDuplicatePointsWithinSameVoxels keeps the first point of a voxel
    const double dx = xyz[3 * i] - sx, dy = xyz[3 * i + 1] - sy, dz = xyz[3 * i + 2] - sz;
    const double length = sqrt(dx * dx + dy * dy + dz * dz);
    const double ux = dx / length, uy = dy / length, uz = dz / length;
    double mp = length - trunc;
    if (max_len < mp) mp = max_len;
    if (step > mp) mp = step;
    if (!(mp == mp)) continue;
    double distance = 0.0;
    while (distance < mp) {
      const double cx = distance * ux + sx, cy = distance * uy + sy, cz = distance * uz + sz;
      const int ckx = (int)floor(cx / voxel), cky = (int)floor(cy / voxel), ckz = (int)floor(cz / voxel);
      bool center_added = false;
      if (radius > 0.0) {
        for (double ox = -radius; ox <= radius; ox += voxel)
          for (double oy = -radius; oy <= radius; oy += voxel)
            for (double oz = -radius; oz <= radius; oz += voxel) {
              const double tx = cx + ox, ty = cy + oy, tz = cz + oz;
              const int kx = (int)floor(tx / voxel), ky = (int)floor(ty / voxel), kz = (int)floor(tz / voxel);
              const double ex = tx - ((double)kx * voxel + voxel * 0.5), ey = ty - ((double)ky * voxel + voxel * 0.5),
                           ez = tz - ((double)kz * voxel + voxel * 0.5);
              if (sqrt(ex * ex + ey * ey + ez * ez) <= radius) {
                const long long s = dense_find_key(dkeys, dcap, kx, ky, kz);
                if (s >= 0 && dcnt[s] > 0) rm[s] = 1;
                if (kx == ckx && ky == cky && kz == ckz) center_added = true;
              }
            }
      }
      if (!center_added) {
        const long long s = dense_find_key(dkeys, dcap, ckx, cky, ckz);
        if (s >= 0 && dcnt[s] > 0) rm[s] = 1;
      }
      distance += step;
    }
  }
}

__global__ void dcarve_apply_kernel(const int32_t* __restrict__ rm, size_t cap, double* __restrict__ sums, int32_t* __restrict__ cnts, int32_t* removed,
                                    const int32_t* __restrict__ enable, int32_t* mstate) {
  if (enable != nullptr && *enable == 0) return;
  if (mstate && blockIdx.x == 0 && threadIdx.x == 0) mstate[MS_NDCARVE] += 1;
  int c = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
    if (!rm[i]) continue;
    cnts[i] = 0;
    for (int k = 0; k < 6; k++) sums[6 * i + k] = 0.0;
    c++;
  }
  c = warp_sum_i(c);
  if ((threadIdx.x & 31) == 0 && c) { atomicAdd(removed, c); if (mstate) atomicAdd(&mstate[MS_DCARVED], c); }
}

int32_t op_dense_carve(b2s_handle* h, b2s_submap* sm, const b2s_cloud* scan, const double* sensor, const double* sensor_dev, double radius,
                       double trunc, double max_len, int32_t* removed_dev, const int32_t* enable_dev) {
  B2S_REQUIRE(sm->dense_cap > 0, B2S_E_INVALID, "dense map not initialised");
  const size_t n_max = scan->n_max > 0 ? scan->n_max : 1;
  size_t cap = 1024;
  while (cap < 2 * n_max) cap <<= 1;
  B2S_TRY(h->keys.ensure(cap * 8, h->stream));
  B2S_TRY(h->vals.ensure(cap * 4, h->stream));
  B2S_TRY(h->tmp_i32.ensure((n_max + 64) * 4, h->stream));
  B2S_TRY(h->offs.ensure((sm->dense_cap + 2) * 4, h->stream));   // removal flags per dense slot
  unsigned long long* keys = h->keys.as<unsigned long long>();
  int32_t* first = h->vals.as<int32_t>();
  int32_t* slot_of = h->tmp_i32.as<int32_t>();
  int32_t* rm = h->offs.as<int32_t>();
  const double voxel = sm->dense_voxel;
  const double s0 = sensor ? sensor[0] : 0.0, s1 = sensor ? sensor[1] : 0.0, s2 = sensor ? sensor[2] : 0.0;
  ProfScope prof(h, PK_FUSE);
  dcarve_init_kernel<<<148 * 8, 256, 0, h->stream>>>(keys, first, cap, rm, sm->dense_cap, enable_dev, removed_dev);
  dcarve_first_kernel<<<grid_for(n_max, FZ_THREADS), FZ_THREADS, 0, h->stream>>>(scan->xyz.as<double>(), scan->dn.as<int32_t>(), 1.0 / voxel, keys, first,
                                                                                cap - 1, slot_of, enable_dev);
  dcarve_march_kernel<<<grid_for(n_max, FZ_THREADS), FZ_THREADS, 0, h->stream>>>(scan->xyz.as<double>(), scan->dn.as<int32_t>(), slot_of, first, s0, s1, s2,
                                                                                sensor_dev, voxel, radius, trunc, max_len,
                                                                                sm->dense_keys.as<unsigned long long>(), sm->dense_cnt.as<int32_t>(),
                                                                                sm->dense_cap, rm, enable_dev);
  dcarve_apply_kernel<<<148 * 8, 256, 0, h->stream>>>(rm, sm->dense_cap, sm->dense_sum.as<double>(), sm->dense_cnt.as<int32_t>(), removed_dev, enable_dev,
                                                      sm->mstate.as<int32_t>());
  h->launches += 4;
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

}  // namespace b2s
