// overlap.cu -- K-overlap: the overlap selection in front of the loop-closure ICP.
//
// Reference: computeIndicesOfOverlappingPoints (core/src/helpers.cpp:307-332), called from
// PlaceRecognition::buildLoopClosureConstraints (core/src/PlaceRecognition.cpp:103-106): both clouds are binned into a
// VoxelMap (key = floor(p * (1/voxel)), the source after sourceToTarget), and a point survives when its voxel holds at least
// minNumPointsPerVoxel points of BOTH clouds.  The reference walks an unordered_map with string-keyed layers and returns
// index lists in hash order; the callers only feed them to SelectByIndex, so the selected SETS are the contract.
//
// Device: one open-addressing hash of packed voxel keys with a (source, target) counter pair per slot; every point
// remembers its slot, a second pass turns the counters into keep-flags, and the two clouds are compacted in their
// original order.  Three kernels + two compactions, all bandwidth-trivial (the clouds are a few MB).
#include "common.cuh"

namespace b2s {

constexpr unsigned long long OV_EMPTY = ~0ull;
constexpr int OV_THREADS = 256;

__device__ __forceinline__ unsigned long long ov_pack(int x, int y, int z) {
  return ((unsigned long long)(unsigned)(x + 1048576) << 42) | ((unsigned long long)(unsigned)(y + 1048576) << 21) |
         (unsigned long long)(unsigned)(z + 1048576);
}
__device__ __forceinline__ unsigned long long ov_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

__global__ void __launch_bounds__(OV_THREADS) ov_init_kernel(unsigned long long* __restrict__ keys, int32_t* __restrict__ cnt, size_t cap) {
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
    keys[i] = OV_EMPTY; cnt[2 * i] = 0; cnt[2 * i + 1] = 0;
  }
}

// which = 0: source (transformed by T like [O3D] PointCloud::Transform), which = 1: target
__global__ void __launch_bounds__(OV_THREADS) ov_insert_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ d_n,
                                                               const double* __restrict__ Tdev, int which, double inv, unsigned long long* keys,
                                                               int32_t* cnt, size_t mask, int32_t* __restrict__ slot_of, uint32_t* status) {
  pdl_wait();
  const int n = *d_n;
  double T[16];
#pragma unroll
  for (int i = 0; i < 16; i++) T[i] = Tdev ? Tdev[i] : ((i % 5 == 0) ? 1.0 : 0.0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    if (which == 0) {
      const double a = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], x), __dmul_rn(T[1], y)), __dmul_rn(T[2], z)), T[3]);
      const double b = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], x), __dmul_rn(T[5], y)), __dmul_rn(T[6], z)), T[7]);
      const double c = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], x), __dmul_rn(T[9], y)), __dmul_rn(T[10], z)), T[11]);
      const double w = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[12], x), __dmul_rn(T[13], y)), __dmul_rn(T[14], z)), T[15]);
      x = __ddiv_rn(a, w); y = __ddiv_rn(b, w); z = __ddiv_rn(c, w);
    }
    const double fx = floor(x * inv), fy = floor(y * inv), fz = floor(z * inv);
    slot_of[i] = -1;
    if (!(fabs(fx) < 1048575.0 && fabs(fy) < 1048575.0 && fabs(fz) < 1048575.0)) { atomicOr(status, ST_KEY_OVERFLOW); continue; }
    const unsigned long long key = ov_pack((int)fx, (int)fy, (int)fz);
    size_t s = (size_t)ov_hash(key) & mask;
    for (size_t probe = 0; probe <= mask; ++probe, s = (s + 1) & mask) {
      const unsigned long long old = atomicCAS(&keys[s], OV_EMPTY, key);
      if (old == OV_EMPTY || old == key) { atomicAdd(&cnt[2 * s + which], 1); slot_of[i] = (int32_t)s; break; }
    }
  }
}

__global__ void __launch_bounds__(OV_THREADS) ov_flags_kernel(const int32_t* __restrict__ d_n, const int32_t* __restrict__ slot_of,
                                                              const int32_t* __restrict__ cnt, int min_pts, int32_t* __restrict__ keep) {
  pdl_wait();
  const int n = *d_n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = slot_of[i];
    keep[i] = (s >= 0 && cnt[2 * s] >= min_pts && cnt[2 * s + 1] >= min_pts) ? 1 : 0;
  }
}

int32_t compact_cloud(b2s_handle* h, const b2s_cloud* in, const int32_t* flags, b2s_cloud* out, const int32_t* d_n_override = nullptr);   // voxel.cu

int32_t op_overlap(b2s_handle* h, const b2s_cloud* source, const b2s_cloud* target, const double* T_dev, double voxel, int min_pts,
                   b2s_cloud* source_overlap, b2s_cloud* target_overlap) {
  const size_t ns = source->n_max > 0 ? source->n_max : 1, nt = target->n_max > 0 ? target->n_max : 1;
  size_t cap = 1024;
  while (cap < 2 * (ns + nt)) cap <<= 1;
  B2S_TRY(h->keys.ensure(cap * 8, h->stream));
  B2S_TRY(h->vals.ensure(cap * 8, h->stream));                       // (source, target) counters
  B2S_TRY(h->tmp_i32.ensure((ns + nt + 64) * 4, h->stream));         // slot of every point
  B2S_TRY(h->flags.ensure((ns + nt + 2) * 4, h->stream));
  unsigned long long* keys = h->keys.as<unsigned long long>();
  int32_t* cnt = h->vals.as<int32_t>();
  int32_t* slot_s = h->tmp_i32.as<int32_t>();
  int32_t* slot_t = slot_s + ns;
  int32_t* keep_s = h->flags.as<int32_t>();
  int32_t* keep_t = keep_s + ns + 1;
  const double inv = 1.0 / voxel;   // VoxelMap(Eigen::Vector3d::Constant(voxelSize)) -> fromVoxelSize
  ProfScope prof(h, PK_FUSE);
  launch_pdl(ov_init_kernel, grid_for(cap, OV_THREADS), OV_THREADS, 0, h->stream, keys, cnt, cap);
  launch_pdl(ov_insert_kernel, grid_for(nt, OV_THREADS), OV_THREADS, 0, h->stream, target->xyz.as<double>(), target->dn.as<int32_t>(), nullptr, 1, inv, keys,
                                                                          cnt, cap - 1, slot_t, h->status.as<uint32_t>());
  launch_pdl(ov_insert_kernel, grid_for(ns, OV_THREADS), OV_THREADS, 0, h->stream, source->xyz.as<double>(), source->dn.as<int32_t>(), T_dev, 0, inv, keys,
                                                                          cnt, cap - 1, slot_s, h->status.as<uint32_t>());
  launch_pdl(ov_flags_kernel, grid_for(ns, OV_THREADS), OV_THREADS, 0, h->stream, source->dn.as<int32_t>(), slot_s, cnt, min_pts, keep_s);
  launch_pdl(ov_flags_kernel, grid_for(nt, OV_THREADS), OV_THREADS, 0, h->stream, target->dn.as<int32_t>(), slot_t, cnt, min_pts, keep_t);
  h->launches += 5;
  B2S_TRY(compact_cloud(h, source, keep_s, source_overlap));   // SelectByIndex on the ORIGINAL (untransformed) source
  B2S_TRY(compact_cloud(h, target, keep_t, target_overlap));
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

}  // namespace b2s
