// normals.cu -- K-normals (P3): RegistrationIcpPointToPlane::estimateNormalsOrCovariancesIfNeeded
// (core/src/CloudRegistration.cpp:49-56) = [O3D] EstimateNormals(KDTreeSearchParamHybrid(radius, knn)) +
// NormalizeNormals + OrientNormalsTowardsCameraLocation(0,0,0).
//
// One WARP per query point, three kernels (default path):
//   normals_select2_kernel  gathers the (2R+1)^3 block of grid cells (grid_index.cu) around the query into a per-warp shared-memory
//                           buffer, R grown until the k-th neighbour provably lies inside the block (ball-within-bounds, like a
//                           KD-tree), selects the k nearest by counting (32-bin histogram of d2 + exact ranking of the boundary
//                           bin; ties -> lower index) and sums the nine cumulants with a transposed warp butterfly;
//   normals_finish_kernel   one thread per query: analytic 3x3 eigen-solver ([O3D] FastEigen3x3, geometrictools
//                           RobustEigenSymmetric3x3), normalise, orient;
//   normals_phase2_kernel   the few queries the block gather cannot certify (more than NS2_CAP candidates, or a search radius the
//                           row table cannot cover): ring walk with a sorted k-best list, one entry per lane, shuffle insert.
// The neighbour SET is the oracle's exactly; the cumulants are summed in butterfly order instead of ascending-distance order, which
// moves the normal by < 1e-9.  A thread-per-query variant (normals_kernel<K>) is kept behind B2S_NORMALS_RING_LIMIT for comparison.
#include "common.cuh"

namespace b2s {

constexpr int NK_THREADS = 128;

__device__ __forceinline__ bool lex_less(double da, int ia, double db, int ib) { return da < db || (da == db && ia < ib); }

__device__ __forceinline__ double slab_gap_n(double q, double o, double cell, int i, int n, double eps) {
  double g = 0.0;
  if (i > 0) { double lo = o + (double)i * cell; if (q < lo) g = lo - q; }
  if (i < n - 1) { double hi = o + (double)(i + 1) * cell; if (q > hi) g = q - hi; }
  g -= eps;
  return g > 0.0 ? g : 0.0;
}

__device__ __forceinline__ void cross3d(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3d(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ void eigvec0_dev(const double* A, double eval0, double* out) {
  double row0[3] = {A[0] - eval0, A[1], A[2]}, row1[3] = {A[1], A[4] - eval0, A[5]}, row2[3] = {A[2], A[5], A[8] - eval0};
  double r01[3], r02[3], r12[3];
  cross3d(row0, row1, r01); cross3d(row0, row2, r02); cross3d(row1, row2, r12);
  const double d0 = dot3d(r01, r01), d1 = dot3d(r02, r02), d2 = dot3d(r12, r12);
  double dmax = d0; int imax = 0;
  if (d1 > dmax) { dmax = d1; imax = 1; }
  if (d2 > dmax) { imax = 2; }
  const double* v = imax == 0 ? r01 : (imax == 1 ? r02 : r12);
  const double s = sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
  out[0] = v[0] / s; out[1] = v[1] / s; out[2] = v[2] / s;
}

__device__ void eigvec1_dev(const double* A, const double* e0, double eval1, double* out) {
  double U[3], V[3];
  if (fabs(e0[0]) > fabs(e0[1])) {
    const double inv = 1 / sqrt(e0[0] * e0[0] + e0[2] * e0[2]);
    U[0] = -e0[2] * inv; U[1] = 0; U[2] = e0[0] * inv;
  } else {
    const double inv = 1 / sqrt(e0[1] * e0[1] + e0[2] * e0[2]);
    U[0] = 0; U[1] = e0[2] * inv; U[2] = -e0[1] * inv;
  }
  cross3d(e0, U, V);
  const double AU[3] = {A[0] * U[0] + A[1] * U[1] + A[2] * U[2], A[1] * U[0] + A[4] * U[1] + A[5] * U[2], A[2] * U[0] + A[5] * U[1] + A[8] * U[2]};
  const double AV[3] = {A[0] * V[0] + A[1] * V[1] + A[2] * V[2], A[1] * V[0] + A[4] * V[1] + A[5] * V[2], A[2] * V[0] + A[5] * V[1] + A[8] * V[2]};
  double m00 = dot3d(U, AU) - eval1, m01 = dot3d(U, AV), m11 = dot3d(V, AV) - eval1;
  const double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
  if (a00 >= a11) {
    const double mx = a00 > a01 ? a00 : a01;
    if (mx > 0) {
      if (a00 >= a01) { m01 /= m00; m00 = 1 / sqrt(1 + m01 * m01); m01 *= m00; }
      else { m00 /= m01; m01 = 1 / sqrt(1 + m00 * m00); m00 *= m01; }
      for (int d = 0; d < 3; d++) out[d] = m01 * U[d] - m00 * V[d];
    } else { out[0] = U[0]; out[1] = U[1]; out[2] = U[2]; }
  } else {
    const double mx = a11 > a01 ? a11 : a01;
    if (mx > 0) {
      if (a11 >= a01) { m01 /= m11; m11 = 1 / sqrt(1 + m01 * m01); m01 *= m11; }
      else { m11 /= m01; m01 = 1 / sqrt(1 + m11 * m11); m11 *= m01; }
      for (int d = 0; d < 3; d++) out[d] = m11 * U[d] - m01 * V[d];
    } else { out[0] = U[0]; out[1] = U[1]; out[2] = U[2]; }
  }
}

// eigenvector of the smallest eigenvalue of a symmetric 3x3 (row-major, full)
__device__ void fast_eigen3x3_dev(const double* cov, double* out) {
  double A[9];
  for (int i = 0; i < 9; i++) A[i] = cov[i];
  double mc = A[0];
  for (int i = 1; i < 9; i++) if (A[i] > mc) mc = A[i];
  if (mc == 0) { out[0] = out[1] = out[2] = 0; return; }
  for (int i = 0; i < 9; i++) A[i] /= mc;
  const double norm = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
  if (norm > 0) {
    double eval[3], e0[3], e1[3], e2[3];
    const double q = (A[0] + A[4] + A[8]) / 3;
    const double b00 = A[0] - q, b11 = A[4] - q, b22 = A[8] - q;
    const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
    const double c00 = b11 * b22 - A[5] * A[5];
    const double c01 = A[1] * b22 - A[5] * A[2];
    const double c02 = A[1] * A[5] - b11 * A[2];
    const double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
    double half_det = det * 0.5;
    half_det = fmin(fmax(half_det, -1.0), 1.0);
    const double angle = acos(half_det) / 3.0;
    const double two_thirds_pi = 2.09439510239319549;
    const double beta2 = cos(angle) * 2;
    const double beta0 = cos(angle + two_thirds_pi) * 2;
    const double beta1 = -(beta0 + beta2);
    eval[0] = q + p * beta0; eval[1] = q + p * beta1; eval[2] = q + p * beta2;
    if (half_det >= 0) {
      eigvec0_dev(A, eval[2], e2);
      if (eval[2] < eval[0] && eval[2] < eval[1]) { out[0] = e2[0]; out[1] = e2[1]; out[2] = e2[2]; return; }
      eigvec1_dev(A, e2, eval[1], e1);
      if (eval[1] < eval[0] && eval[1] < eval[2]) { out[0] = e1[0]; out[1] = e1[1]; out[2] = e1[2]; return; }
      cross3d(e1, e2, e0);
      out[0] = e0[0]; out[1] = e0[1]; out[2] = e0[2];
    } else {
      eigvec0_dev(A, eval[0], e0);
      if (eval[0] < eval[1] && eval[0] < eval[2]) { out[0] = e0[0]; out[1] = e0[1]; out[2] = e0[2]; return; }
      eigvec1_dev(A, e0, eval[1], e1);
      if (eval[1] < eval[0] && eval[1] < eval[2]) { out[0] = e1[0]; out[1] = e1[1]; out[2] = e1[2]; return; }
      cross3d(e0, e1, e2);
      out[0] = e2[0]; out[1] = e2[1]; out[2] = e2[2];
    }
  } else {
    const double a0 = A[0] * mc, a1 = A[4] * mc, a2 = A[8] * mc;
    if (a0 < a1 && a0 < a2) { out[0] = 1; out[1] = 0; out[2] = 0; }
    else if (a1 < a0 && a1 < a2) { out[0] = 0; out[1] = 1; out[2] = 0; }
    else { out[0] = 0; out[1] = 0; out[2] = 1; }
  }
}


// Post-search part shared by all KMAX: covariance in the reference's neighbour order (ascending (d2, index)) with
// the reference's single-pass cumulant formula in explicitly rounded fp64, analytic eigen-solver, normalise, orient.
__device__ __forceinline__ void finish_normal(const double c_in[9], int kk, double qx, double qy, double qz, double* nr) {
  double cov[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // [O3D] fewer than 3 neighbours -> identity covariance
  if (kk >= 3) {
    double c[9];
    const double kf = (double)kk;
#pragma unroll
    for (int t = 0; t < 9; t++) c[t] = __ddiv_rn(c_in[t], kf);
    cov[0] = __dsub_rn(c[3], __dmul_rn(c[0], c[0]));
    cov[4] = __dsub_rn(c[6], __dmul_rn(c[1], c[1]));
    cov[8] = __dsub_rn(c[8], __dmul_rn(c[2], c[2]));
    cov[1] = cov[3] = __dsub_rn(c[4], __dmul_rn(c[0], c[1]));
    cov[2] = cov[6] = __dsub_rn(c[5], __dmul_rn(c[0], c[2]));
    cov[5] = cov[7] = __dsub_rn(c[7], __dmul_rn(c[1], c[2]));
  }
  fast_eigen3x3_dev(cov, nr);
  if (sqrt(dot3d(nr, nr)) == 0.0) { nr[0] = 0; nr[1] = 0; nr[2] = 1; }
  const double zz = dot3d(nr, nr);  // NormalizeNormals
  if (zz > 0) { const double sn = sqrt(zz); nr[0] /= sn; nr[1] /= sn; nr[2] /= sn; }
  if (nr[0] != nr[0]) { nr[0] = 0; nr[1] = 0; nr[2] = 1; }
  const double ref[3] = {-qx, -qy, -qz};  // OrientNormalsTowardsCameraLocation(0,0,0)
  if (sqrt(dot3d(nr, nr)) == 0.0) {
    const double rn = sqrt(dot3d(ref, ref));
    if (rn == 0.0) { nr[0] = 0; nr[1] = 0; nr[2] = 1; }
    else { nr[0] = ref[0] / rn; nr[1] = ref[1] / rn; nr[2] = ref[2] / rn; }
  } else if (dot3d(nr, ref) < 0.0) { nr[0] *= -1.0; nr[1] *= -1.0; nr[2] *= -1.0; }
}

// One THREAD per query point, queries taken in grid-slot order so that the 32 lanes of a warp sit in the same or
// in adjacent cells and walk (nearly) the same candidate ranges: the candidate loads are warp-broadcasts out of L1.
// The k best are a sorted list held entirely in REGISTERS (KMAX compile-time, insertion fully unrolled into
// predicated moves; no local memory).  Exact k-NN: ring expansion stops when the k-th distance is below the distance
// to the unvisited shell or the shell is beyond the radius; ties are broken towards the lower original index.
// Queries that still need rings beyond `ring_limit` (sparse far-range areas, isolated points) are NOT finished here:
// their slot is pushed to `queue` and normals_phase2_kernel finishes them with one warp each.  Without that split a
// few threads walking hundreds of empty cells serially set the duration of the whole kernel.
template <int KMAX, bool EXACT>
__global__ void __launch_bounds__(NK_THREADS) normals_kernel(const GridHeader* __restrict__ hdr, const int32_t* __restrict__ cs,
                                                             const double4* __restrict__ pts, int knn, double radius, int ring_limit,
                                                             int32_t* __restrict__ queue, int32_t* queue_n,
                                                             const int32_t* __restrict__ qlist, const int32_t* __restrict__ qcount,
                                                             double* __restrict__ out_nrm) {
  pdl_wait();
  __shared__ GridHeader g;
  if (threadIdx.x == 0) g = *hdr;
  __syncthreads();
  const int n = g.n;
  const double r2 = radius * radius;
  const double eps = 1e-9 * g.cell;
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  const int nq = qlist ? *qcount : n;   // optional query list: only these slots get a normal (fused down-sample)
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nq; t += gridDim.x * blockDim.x) {
    const int s = qlist ? qlist[t] : t;
    if (ring_limit < 0) { queue[atomicAdd(queue_n, 1)] = s; continue; }   // everything to the warp-cooperative kernel
    const double4 qp = pts[s];
    const double qx = qp.x, qy = qp.y, qz = qp.z;
    const int qi = (int)__double_as_longlong(qp.w);
    const int cx = (int)fmin(fmax(floor((qx - g.origin[0]) * g.inv_cell), 0.0), (double)(nx - 1));
    const int cy = (int)fmin(fmax(floor((qy - g.origin[1]) * g.inv_cell), 0.0), (double)(ny - 1));
    const int cz = (int)fmin(fmax(floor((qz - g.origin[2]) * g.inv_cell), 0.0), (double)(nz - 1));
    double bd[KMAX];
    int bs[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; j++) { bd[j] = INFINITY; bs[j] = -1; }
    double kd = INFINITY;   // current k-th best distance (entry knn-1); +inf until k neighbours are known
    int kslot = -1;
    bool unresolved = false;
    for (int R = 0;; ++R) {
      const int z0 = max(cz - R, 0), z1 = min(cz + R, nz - 1);
      const int y0 = max(cy - R, 0), y1 = min(cy + R, ny - 1);
      const int x0 = max(cx - R, 0), x1 = min(cx + R, nx - 1);
      for (int z = z0; z <= z1; ++z) {
        const double gz = slab_gap_n(qz, g.origin[2], g.cell, z, nz, eps);
        const double gz2 = gz * gz;
        if (gz2 > fmin(kd, r2)) continue;
        const bool zface = (z == cz - R) || (z == cz + R);
        for (int y = y0; y <= y1; ++y) {
          const double gy = slab_gap_n(qy, g.origin[1], g.cell, y, ny, eps);
          if (gz2 + gy * gy > fmin(kd, r2)) continue;
          const int row = (z * ny + y) * nx;
          const bool shell = zface || y == cy - R || y == cy + R;
          for (int part = 0; part < 2; ++part) {
            int a, b;
            if (shell) { if (part == 1) break; a = cs[row + x0]; b = cs[row + x1 + 1]; }
            else if (part == 0) { if (cx - R < 0) continue; a = cs[row + cx - R]; b = cs[row + cx - R + 1]; }
            else { if (cx + R > nx - 1) continue; a = cs[row + cx + R]; b = cs[row + cx + R + 1]; }
            for (int j = a; j < b; ++j) {
              const double4 p = pts[j];
              const double d = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
              if (!(d < r2)) continue;
              if (d > kd) continue;
              if (d == kd) {  // exact tie with the current k-th: lower original index wins (rare path)
                const int ik = (int)__double_as_longlong(pts[kslot].w), ic = (int)__double_as_longlong(p.w);
                if (!(ic < ik)) continue;
              }
              // sorted insertion, fully unrolled (registers only).  Equal distances: order by original index.
              const int ic = (int)__double_as_longlong(p.w);
#pragma unroll
              for (int t = KMAX - 1; t >= 1; --t) {
                bool before_prev = d < bd[t - 1];
                if (d == bd[t - 1]) before_prev = ic < (int)__double_as_longlong(pts[bs[t - 1]].w);
                bool before_cur = d < bd[t];
                if (d == bd[t] && bs[t] >= 0) before_cur = ic < (int)__double_as_longlong(pts[bs[t]].w);
                if (before_prev) { bd[t] = bd[t - 1]; bs[t] = bs[t - 1]; }
                else if (before_cur) { bd[t] = d; bs[t] = j; }
              }
              {
                bool before0 = d < bd[0];
                if (d == bd[0] && bs[0] >= 0) before0 = ic < (int)__double_as_longlong(pts[bs[0]].w);
                if (before0) { bd[0] = d; bs[0] = j; }
              }
              if (EXACT) { kd = bd[KMAX - 1]; kslot = bs[KMAX - 1]; }   // knn == KMAX: static index, the list stays in registers
              else { kd = bd[knn - 1]; kslot = bs[knn - 1]; }          // generic knn: dynamic index (list lives in local memory)
            }
          }
        }
      }
      double bound = INFINITY;
      if (cx - R > 0) bound = fmin(bound, qx - (g.origin[0] + (double)(cx - R) * g.cell));
      if (cx + R < nx - 1) bound = fmin(bound, (g.origin[0] + (double)(cx + R + 1) * g.cell) - qx);
      if (cy - R > 0) bound = fmin(bound, qy - (g.origin[1] + (double)(cy - R) * g.cell));
      if (cy + R < ny - 1) bound = fmin(bound, (g.origin[1] + (double)(cy + R + 1) * g.cell) - qy);
      if (cz - R > 0) bound = fmin(bound, qz - (g.origin[2] + (double)(cz - R) * g.cell));
      if (cz + R < nz - 1) bound = fmin(bound, (g.origin[2] + (double)(cz + R + 1) * g.cell) - qz);
      bound -= eps;
      if (bound < 0.0) bound = 0.0;
      if (bound == INFINITY || bound * bound > fmin(kd, r2)) break;
      if (R >= ring_limit) { unresolved = true; break; }
    }
    if (unresolved) { queue[atomicAdd(queue_n, 1)] = s; continue; }
    // cumulants over the kk <= knn neighbours in ascending (d2, index) order
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int kk = 0;
#pragma unroll
    for (int t = 0; t < KMAX; ++t) {
      if (t < knn && bs[t] >= 0) {
        const double4 p = pts[bs[t]];
        c[0] = __dadd_rn(c[0], p.x); c[1] = __dadd_rn(c[1], p.y); c[2] = __dadd_rn(c[2], p.z);
        c[3] = __dadd_rn(c[3], __dmul_rn(p.x, p.x)); c[4] = __dadd_rn(c[4], __dmul_rn(p.x, p.y)); c[5] = __dadd_rn(c[5], __dmul_rn(p.x, p.z));
        c[6] = __dadd_rn(c[6], __dmul_rn(p.y, p.y)); c[7] = __dadd_rn(c[7], __dmul_rn(p.y, p.z)); c[8] = __dadd_rn(c[8], __dmul_rn(p.z, p.z));
        kk++;
      }
    }
    double nr[3];
    finish_normal(c, kk, qx, qy, qz, nr);
    out_nrm[3 * (size_t)qi] = nr[0]; out_nrm[3 * (size_t)qi + 1] = nr[1]; out_nrm[3 * (size_t)qi + 2] = nr[2];
  }
}

// Phase 2: one WARP per queued query, restarted from ring 0.  Per ring, each lane first resolves ONE (y, z) row --
// pruning test and the two dependent cell_start loads, the latency that dominates in empty space -- then the warp
// walks the non-empty rows together: 32 candidates per step, the k best kept as a sorted list with one entry per lane
// (k <= 32), a qualifying candidate inserted with a single shuffle-up step.
__global__ void __launch_bounds__(NK_THREADS) normals_phase2_kernel(const GridHeader* __restrict__ hdr, const int32_t* __restrict__ cs,
                                                                    const double4* __restrict__ pts, int knn, double radius,
                                                                    const int32_t* __restrict__ queue, const int32_t* __restrict__ queue_n,
                                                                    double* __restrict__ out_nrm) {
  pdl_wait();
  __shared__ GridHeader g;
  if (threadIdx.x == 0) g = *hdr;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (NK_THREADS / 32);
  const int nq = *queue_n;
  const double r2 = radius * radius;
  const double eps = 1e-9 * g.cell;
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  for (int w = blockIdx.x * (NK_THREADS / 32) + (threadIdx.x >> 5); w < nq; w += warps_total) {
    const int s = queue[w];
    const double4 qp = pts[s];
    const double qx = qp.x, qy = qp.y, qz = qp.z;
    const int qi = (int)__double_as_longlong(qp.w);
    const int cx = (int)fmin(fmax(floor((qx - g.origin[0]) * g.inv_cell), 0.0), (double)(nx - 1));
    const int cy = (int)fmin(fmax(floor((qy - g.origin[1]) * g.inv_cell), 0.0), (double)(ny - 1));
    const int cz = (int)fmin(fmax(floor((qz - g.origin[2]) * g.inv_cell), 0.0), (double)(nz - 1));
    double ed = INFINITY; int ei = 0x7fffffff; int es = -1;  // this lane's entry of the sorted k-best list
    double kd = INFINITY; int ki = 0x7fffffff;               // current k-th best (lane knn-1)
    for (int R = 0;; ++R) {
      const int side = 2 * R + 1;
      const int x0 = max(cx - R, 0), x1 = min(cx + R, nx - 1);
      for (int t0 = 0; t0 < side * side; t0 += 32) {
        // each lane resolves one row: up to two candidate ranges [a0,b0) and [a1,b1)
        int a0 = 0, b0 = 0, a1 = 0, b1 = 0;
        const int t = t0 + lane;
        if (t < side * side) {
          const int z = cz - R + t / side, y = cy - R + t % side;
          if (z >= 0 && z < nz && y >= 0 && y < ny) {
            const double gz = slab_gap_n(qz, g.origin[2], g.cell, z, nz, eps);
            const double gy = slab_gap_n(qy, g.origin[1], g.cell, y, ny, eps);
            if (gz * gz + gy * gy <= fmin(kd, r2)) {
              const int row = (z * ny + y) * nx;
              if (z == cz - R || z == cz + R || y == cy - R || y == cy + R) { a0 = cs[row + x0]; b0 = cs[row + x1 + 1]; }
              else {
                if (cx - R >= 0) { a0 = cs[row + cx - R]; b0 = cs[row + cx - R + 1]; }
                if (cx + R <= nx - 1) { a1 = cs[row + cx + R]; b1 = cs[row + cx + R + 1]; }
              }
            }
          }
        }
        for (int part = 0; part < 2; ++part) {
          unsigned rows = __ballot_sync(0xffffffffu, part == 0 ? (b0 > a0) : (b1 > a1));
          while (rows) {
            const int src_lane = __ffs(rows) - 1;
            rows &= rows - 1;
            const int a = __shfl_sync(0xffffffffu, part == 0 ? a0 : a1, src_lane);
            const int b = __shfl_sync(0xffffffffu, part == 0 ? b0 : b1, src_lane);
            for (int j0 = a; j0 < b; j0 += 32) {
              const int j = j0 + lane;
              double d = INFINITY; int idx = 0x7fffffff;
              if (j < b) {
                const double4 p = pts[j];
                d = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
                idx = (int)__double_as_longlong(p.w);
              }
              unsigned mask = __ballot_sync(0xffffffffu, j < b && d < r2 && lex_less(d, idx, kd, ki));
              while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const double cd = __shfl_sync(0xffffffffu, d, src);
                const int ci = __shfl_sync(0xffffffffu, idx, src);
                const int cslot = j0 + src;
                const double pd = __shfl_up_sync(0xffffffffu, ed, 1);
                const int pi = __shfl_up_sync(0xffffffffu, ei, 1);
                const int ps = __shfl_up_sync(0xffffffffu, es, 1);
                if (lex_less(cd, ci, ed, ei)) {
                  if (lane > 0 && lex_less(cd, ci, pd, pi)) { ed = pd; ei = pi; es = ps; }
                  else { ed = cd; ei = ci; es = cslot; }
                }
                kd = __shfl_sync(0xffffffffu, ed, knn - 1);
                ki = __shfl_sync(0xffffffffu, ei, knn - 1);
              }
            }
          }
        }
      }
      double bound = INFINITY;
      if (cx - R > 0) bound = fmin(bound, qx - (g.origin[0] + (double)(cx - R) * g.cell));
      if (cx + R < nx - 1) bound = fmin(bound, (g.origin[0] + (double)(cx + R + 1) * g.cell) - qx);
      if (cy - R > 0) bound = fmin(bound, qy - (g.origin[1] + (double)(cy - R) * g.cell));
      if (cy + R < ny - 1) bound = fmin(bound, (g.origin[1] + (double)(cy + R + 1) * g.cell) - qy);
      if (cz - R > 0) bound = fmin(bound, qz - (g.origin[2] + (double)(cz - R) * g.cell));
      if (cz + R < nz - 1) bound = fmin(bound, (g.origin[2] + (double)(cz + R + 1) * g.cell) - qz);
      bound -= eps;
      if (bound < 0.0) bound = 0.0;
      if (bound == INFINITY || bound * bound > fmin(kd, r2)) break;
    }
    // cumulants in ascending (d2, index) order: lane t holds the t-th neighbour, broadcast by shuffles
    const int kk = __popc(__ballot_sync(0xffffffffu, lane < knn && es >= 0));
    double4 np = make_double4(0, 0, 0, 0);
    if (lane < kk) np = pts[es];
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < kk; ++t) {
      const double x = __shfl_sync(0xffffffffu, np.x, t), y = __shfl_sync(0xffffffffu, np.y, t), z = __shfl_sync(0xffffffffu, np.z, t);
      c[0] = __dadd_rn(c[0], x); c[1] = __dadd_rn(c[1], y); c[2] = __dadd_rn(c[2], z);
      c[3] = __dadd_rn(c[3], __dmul_rn(x, x)); c[4] = __dadd_rn(c[4], __dmul_rn(x, y)); c[5] = __dadd_rn(c[5], __dmul_rn(x, z));
      c[6] = __dadd_rn(c[6], __dmul_rn(y, y)); c[7] = __dadd_rn(c[7], __dmul_rn(y, z)); c[8] = __dadd_rn(c[8], __dmul_rn(z, z));
    }
    if (lane == 0) {
      double nr[3];
      finish_normal(c, kk, qx, qy, qz, nr);
      out_nrm[3 * (size_t)qi] = nr[0]; out_nrm[3 * (size_t)qi + 1] = nr[1]; out_nrm[3 * (size_t)qi + 2] = nr[2];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fast path: one warp per query, GATHER the candidates of the 3x3x3 cell block once, SELECT the k nearest by counting,
// reduce the covariance with warp shuffles.
//   1. lanes 0..8 resolve the nine (y, z) rows of the block (one contiguous slot range each);
//   2. the candidates (at most NS_CHUNKS*32) are dealt round-robin to the lanes: distance + index live in registers;
//   3. the k-th smallest (d2, index) is found WITHOUT sorting: a 32-bin histogram over d2 (neighbours on a surface are
//      ~uniform in d2) built with shared-memory atomics, a warp prefix sum to locate the bin that holds the k-th, and a
//      few warp arg-min rounds inside that bin;
//   4. every lane accumulates the cumulants of its own selected candidates, a butterfly of shuffles sums them
//      (the covariance is therefore summed in a different order than the reference's ascending-distance order: the
//      difference is O(1e-16) relative);
//   5. the query is exact iff the k-th distance lies inside the scanned block (or the block already covers the radius);
//      anything else -- sparse neighbourhoods, more than NS_CHUNKS*32 candidates -- is queued for normals_phase2_kernel.
// A warp handles 32 consecutive queries and only then runs the eigen-solver, one query per lane.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int NS_CHUNKS = 8;

__global__ void __launch_bounds__(NK_THREADS) normals_select_kernel(const GridHeader* __restrict__ hdr, const int32_t* __restrict__ cs,
                                                                    const double4* __restrict__ pts, int knn, double radius,
                                                                    const int32_t* __restrict__ qlist, const int32_t* __restrict__ qcount,
                                                                    int32_t* __restrict__ queue, int32_t* queue_n,
                                                                    double* __restrict__ cum) {
  pdl_wait();
  __shared__ GridHeader g;
  __shared__ int s_hist[NK_THREADS / 32][32];
  if (threadIdx.x == 0) g = *hdr;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n = g.n;
  const int nq = qlist ? *qcount : n;
  const double r2 = radius * radius;
  const double eps = 1e-9 * g.cell;
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  const int warps_total = gridDim.x * (NK_THREADS / 32);
  {
    for (int tq = blockIdx.x * (NK_THREADS / 32) + wib; tq < nq; tq += warps_total) {
      const int s = qlist ? qlist[tq] : tq;
      const double4 qp = pts[s];
      const double qx = qp.x, qy = qp.y, qz = qp.z;
      const int cx = (int)fmin(fmax(floor((qx - g.origin[0]) * g.inv_cell), 0.0), (double)(nx - 1));
      const int cy = (int)fmin(fmax(floor((qy - g.origin[1]) * g.inv_cell), 0.0), (double)(ny - 1));
      const int cz = (int)fmin(fmax(floor((qz - g.origin[2]) * g.inv_cell), 0.0), (double)(nz - 1));
      // 1. rows of the 3x3x3 block
      int a = 0, cnt = 0;
      if (lane < 9) {
        const int y = cy - 1 + lane % 3, z = cz - 1 + lane / 3;
        if (y >= 0 && y < ny && z >= 0 && z < nz) {
          const int row = (z * ny + y) * nx;
          a = cs[row + max(cx - 1, 0)];
          cnt = cs[row + min(cx + 1, nx - 1) + 1] - a;
        }
      }
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      const int ntot = __shfl_sync(0xffffffffu, inc, 8);
      const int delta_l = a - (inc - cnt);   // slot = t + delta for candidate number t of this row
      int offs[9], delta[9];
#pragma unroll
      for (int r = 0; r < 9; r++) { offs[r] = __shfl_sync(0xffffffffu, inc - cnt, r); delta[r] = __shfl_sync(0xffffffffu, delta_l, r); }
      bool fallback = ntot > NS_CHUNKS * 32;
      if (fallback && lane == 0) atomicAdd(queue_n + 2, 1);   // debug counter: block holds too many candidates
      // 2. candidates -> registers
      double d[NS_CHUNKS]; int idx[NS_CHUNKS], sl[NS_CHUNKS];
      int nvalid = 0;
      double dmax = 0.0;
      if (!fallback) {
#pragma unroll
        for (int c = 0; c < NS_CHUNKS; c++) {
          const int t = c * 32 + lane;
          int dl = delta[0];
#pragma unroll
          for (int r = 1; r < 9; r++) if (t >= offs[r]) dl = delta[r];
          d[c] = INFINITY; idx[c] = 0x7fffffff; sl[c] = -1;
          if (t < ntot) {
            const int slot = t + dl;
            const double4 p = pts[slot];
            const double dd = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
            if (dd < r2) { d[c] = dd; idx[c] = (int)__double_as_longlong(p.w); sl[c] = slot; dmax = fmax(dmax, dd); }
          }
          nvalid += __popc(__ballot_sync(0xffffffffu, sl[c] >= 0));
        }
      }
      // 3. threshold (td, ti): the `need`-th smallest (d2, index)
      const int need = min(knn, nvalid);
      double td = INFINITY; int ti = 0x7fffffff;   // nvalid <= knn: everything valid is selected
      if (!fallback && nvalid > knn) {
        dmax = warp_max(dmax);
        const double scale = dmax > 0.0 ? 32.0 / dmax : 0.0;
        s_hist[wib][lane] = 0;
        __syncwarp();
        int bin[NS_CHUNKS];
#pragma unroll
        for (int c = 0; c < NS_CHUNKS; c++) {
          bin[c] = 32;
          if (sl[c] >= 0) { bin[c] = min(31, (int)(d[c] * scale)); atomicAdd(&s_hist[wib][bin[c]], 1); }
        }
        __syncwarp();
        int cum = s_hist[wib][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += t; }
        const int B = __ffs(__ballot_sync(0xffffffffu, cum >= need)) - 1;
        const int below = B > 0 ? __shfl_sync(0xffffffffu, cum, B - 1) : 0;
        const int m = need - below;    // how many of bin B belong to the k nearest (>= 1)
        double ld = -1.0; int li = -1; // last extracted (d2, index), lexicographic lower bound
        for (int round = 0; round < m; ++round) {
          double bd = INFINITY; int bi = 0x7fffffff;
#pragma unroll
          for (int c = 0; c < NS_CHUNKS; c++) {
            const bool in_bin = bin[c] == B;
            const bool after = d[c] > ld || (d[c] == ld && idx[c] > li);
            const bool better = d[c] < bd || (d[c] == bd && idx[c] < bi);
            if (in_bin && after && better) { bd = d[c]; bi = idx[c]; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
          }
          ld = bd; li = bi;
        }
        td = ld; ti = li;
        // everything in a lower bin is selected, bin B up to (td, ti): express both with one lexicographic threshold
#pragma unroll
        for (int c = 0; c < NS_CHUNKS; c++) if (bin[c] > B) sl[c] = -1;   // beyond the k-th
#pragma unroll
        for (int c = 0; c < NS_CHUNKS; c++) if (bin[c] == B && (d[c] > td || (d[c] == td && idx[c] > ti))) sl[c] = -1;
      }
      // 5. exact?  the k-th distance must lie inside the scanned block, or the block must cover the whole radius
      double bound = INFINITY;
      if (cx - 1 > 0) bound = fmin(bound, qx - (g.origin[0] + (double)(cx - 1) * g.cell));
      if (cx + 1 < nx - 1) bound = fmin(bound, (g.origin[0] + (double)(cx + 2) * g.cell) - qx);
      if (cy - 1 > 0) bound = fmin(bound, qy - (g.origin[1] + (double)(cy - 1) * g.cell));
      if (cy + 1 < ny - 1) bound = fmin(bound, (g.origin[1] + (double)(cy + 2) * g.cell) - qy);
      if (cz - 1 > 0) bound = fmin(bound, qz - (g.origin[2] + (double)(cz - 1) * g.cell));
      if (cz + 1 < nz - 1) bound = fmin(bound, (g.origin[2] + (double)(cz + 2) * g.cell) - qz);
      if (bound != INFINITY) { bound -= eps; if (bound < 0.0) bound = 0.0; }
      const double b2 = bound == INFINITY ? INFINITY : bound * bound;
      const double kth = nvalid >= knn ? td : INFINITY;   // fewer than k found: only exact if the block covers the radius
      if (!fallback && !(b2 > fmin(kth, r2))) { fallback = true; if (lane == 0) atomicAdd(queue_n + (nvalid >= knn ? 3 : 4), 1); }  // debug counters: k-th outside the block / fewer than k in the block
      if (fallback) {
        if (lane == 0) { queue[atomicAdd(queue_n, 1)] = s; cum[10 * (size_t)tq + 9] = -1.0; }
        continue;
      }
      // 4. cumulants of the selected candidates, butterfly sum
      double c9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int c = 0; c < NS_CHUNKS; c++) {
        if (sl[c] >= 0) {
          const double4 p = pts[sl[c]];
          c9[0] += p.x; c9[1] += p.y; c9[2] += p.z;
          c9[3] += p.x * p.x; c9[4] += p.x * p.y; c9[5] += p.x * p.z;
          c9[6] += p.y * p.y; c9[7] += p.y * p.z; c9[8] += p.z * p.z;
        }
      }
#pragma unroll
      for (int t = 0; t < 9; t++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c9[t] += __shfl_xor_sync(0xffffffffu, c9[t], o);
      }
      {
        double v = (double)need;   // lane 9 writes the neighbour count, lanes 0..8 one cumulant each (all lanes hold the sums)
#pragma unroll
        for (int t = 0; t < 9; t++) if (lane == t) v = c9[t];
        if (lane < 10) cum[10 * (size_t)tq + lane] = v;
      }
    }
  }
}

// Butterfly sum of 9 values over the warp with the exchanges TRANSPOSED: at distance 16 the two halves of the warp split the values
// between them (each lane keeps the half it will finish and receives the partner's copy of it), at distance 8 the quarters do, and so
// on -- 8 + 4 + 2 + 1 + 1 = 16 exchanges instead of 9 x 5.  Every partial sum is the same pair of operands the plain xor butterfly
// adds at that distance (fp addition commutes), so the totals are bit-identical to it.  On return lane l holds the total of value
// l >> 1 in v[0] (values 9..15 are padding).
__device__ __forceinline__ void warp_sum9_transposed(double (&v)[9], int lane) {
  const unsigned FULL = 0xffffffffu;
  double a[8];
  {  // distance 16: lower half keeps 0..7, upper half keeps 8..15 (only 8 is real)
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double hi = k == 0 ? v[8] : 0.0;                 // value 8 + k
      const double send = up ? v[k] : hi;
      const double recv = __shfl_xor_sync(FULL, send, 16);
      a[k] = (up ? hi : v[k]) + recv;
    }
  }
  double b[4];
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double send = up ? a[k] : a[k + 4];
      const double recv = __shfl_xor_sync(FULL, send, 8);
      b[k] = (up ? a[k + 4] : a[k]) + recv;
    }
  }
  double c[2];
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const double send = up ? b[k] : b[k + 2];
      const double recv = __shfl_xor_sync(FULL, send, 4);
      c[k] = (up ? b[k + 2] : b[k]) + recv;
    }
  }
  double d;
  {
    const bool up = (lane & 2) != 0;
    const double send = up ? c[0] : c[1];
    const double recv = __shfl_xor_sync(FULL, send, 2);
    d = (up ? c[1] : c[0]) + recv;
  }
  d += __shfl_xor_sync(FULL, d, 1);
  v[0] = d;
}

// Second-generation fast path: same GATHER -> SELECT-BY-COUNTING -> BUTTERFLY pipeline as normals_select_kernel, but the
// gathered candidates go through a per-warp shared-memory buffer, which lets the block radius R grow (1, 2, 3 cells)
// until the k-th neighbour provably lies inside the block: dense areas finish at R = 1, sparse far-range areas at
// R = 2 or 3, and only what is still unresolved (or holds more than NS2_CAP candidates) goes to normals_phase2_kernel.
#ifndef B2S_NS2_MINBLOCKS
#define B2S_NS2_MINBLOCKS 8   // resident CTAs per SM the select kernel is compiled for (8 -> 64 registers; A/B knob of the build)
#endif
constexpr int NS2_CAP = 256;
constexpr int NS2_CHUNKS = NS2_CAP / 32;
constexpr int NS2_RMAX = 3;
constexpr int NS2_ROWS = 320;   // row-table entries per warp: (2 R + 1)^2 rows of the largest block, rounded up to 32 (R = 8 -> 289)

__global__ void __launch_bounds__(NK_THREADS, B2S_NS2_MINBLOCKS) normals_select2_kernel(const GridHeader* __restrict__ hdr, const int32_t* __restrict__ cs,
                                                                     const double4* __restrict__ pts, int knn, double radius,
                                                                     const int32_t* __restrict__ qlist, const int32_t* __restrict__ qcount,
                                                                     int32_t* __restrict__ queue, int32_t* queue_n,
                                                                     double* __restrict__ cum) {
  pdl_wait();
  __shared__ GridHeader g;
  __shared__ double s_d[NK_THREADS / 32][NS2_CAP];
  __shared__ int s_i[NK_THREADS / 32][NS2_CAP];
  __shared__ int s_s[NK_THREADS / 32][NS2_CAP];
  __shared__ int s_hist[NK_THREADS / 32][32];
  __shared__ int s_ra[NK_THREADS / 32][NS2_ROWS];       // first slot of every (y, z) row of the current block
  __shared__ int s_rp[NK_THREADS / 32][NS2_ROWS + 1];   // exclusive prefix of the row sizes
  if (threadIdx.x == 0) g = *hdr;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int n = g.n;
  const int nq = qlist ? *qcount : n;
  const double r2 = radius * radius;
  const double eps = 1e-9 * g.cell;
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  const int warps_total = gridDim.x * (NK_THREADS / 32);
  for (int tq = blockIdx.x * (NK_THREADS / 32) + wib; tq < nq; tq += warps_total) {
    const int s = qlist ? qlist[tq] : tq;
    const double4 qp = pts[s];
    const double qx = qp.x, qy = qp.y, qz = qp.z;
    const int cx = (int)fmin(fmax(floor((qx - g.origin[0]) * g.inv_cell), 0.0), (double)(nx - 1));
    const int cy = (int)fmin(fmax(floor((qy - g.origin[1]) * g.inv_cell), 0.0), (double)(ny - 1));
    const int cz = (int)fmin(fmax(floor((qz - g.origin[2]) * g.inv_cell), 0.0), (double)(nz - 1));
    bool resolved = false;
    int need = 0;
    double c9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    // starting radius from the local density: 25 lanes read the run boundaries of the 5x5 rows once; a 3x3x3 block with
    // fewer than ~3k points will almost never contain the k-th neighbour provably, so sparse queries start at R = 2
    int Rstart = 1;
    {
      int c1 = 0, c2 = 0;
      if (lane < 25) {
        const int dy = lane % 5 - 2, dz = lane / 5 - 2;
        const int y = cy + dy, z = cz + dz;
        if (y >= 0 && y < ny && z >= 0 && z < nz) {
          const int row = (z * ny + y) * nx;
          c2 = cs[row + min(cx + 2, nx - 1) + 1] - cs[row + max(cx - 2, 0)];
          if (dy >= -1 && dy <= 1 && dz >= -1 && dz <= 1) c1 = cs[row + min(cx + 1, nx - 1) + 1] - cs[row + max(cx - 1, 0)];
        }
      }
      const int n1 = warp_sum_i(c1), n2 = warp_sum_i(c2);
      const int n1b = __shfl_sync(0xffffffffu, n1, 0), n2b = __shfl_sync(0xffffffffu, n2, 0);
      if (n1b < 3 * knn && n2b <= NS2_CAP) Rstart = 2;
    }
    // Block radii tried: Rstart .. NS2_RMAX, then -- for the sparse queries that are still open (isolated far-range points, whose k
    // neighbours lie further apart than any certified ball of a small block) -- ONE block that covers the whole search radius, if its
    // rows fit the table: everything within the radius is then in the buffer and the answer is final even with fewer than k members.
    const int Rfull = (int)ceil(radius * g.inv_cell);
    const bool full_fits = Rfull > NS2_RMAX && (2 * Rfull + 1) * (2 * Rfull + 1) <= NS2_ROWS;
    constexpr int R_NONE = 1 << 20;
    for (int R = Rstart; !resolved && R != R_NONE; R = (R < NS2_RMAX ? R + 1 : (full_fits && R == NS2_RMAX ? Rfull : R_NONE))) {
      const bool last_try = R > NS2_RMAX || (R == NS2_RMAX && !full_fits);
      // ---- gather the (2R+1)^3 block into shared memory (valid = inside the radius), rows resolved by the lanes ----
      const int side = 2 * R + 1;
      const int x0 = max(cx - R, 0), x1 = min(cx + R, nx - 1);
      double bound = INFINITY;   // distance from the query to the nearest face of the block that has cells beyond it
      if (cx - R > 0) bound = fmin(bound, qx - (g.origin[0] + (double)(cx - R) * g.cell));
      if (cx + R < nx - 1) bound = fmin(bound, (g.origin[0] + (double)(cx + R + 1) * g.cell) - qx);
      if (cy - R > 0) bound = fmin(bound, qy - (g.origin[1] + (double)(cy - R) * g.cell));
      if (cy + R < ny - 1) bound = fmin(bound, (g.origin[1] + (double)(cy + R + 1) * g.cell) - qy);
      if (cz - R > 0) bound = fmin(bound, qz - (g.origin[2] + (double)(cz - R) * g.cell));
      if (cz + R < nz - 1) bound = fmin(bound, (g.origin[2] + (double)(cz + R + 1) * g.cell) - qz);
      if (bound != INFINITY) { bound -= eps; if (bound < 0.0) bound = 0.0; }
      const double b2 = bound == INFINITY ? INFINITY : bound * bound;
      const double lim2 = fmin(r2, b2);   // candidates beyond the guaranteed ball cannot be certified at this R: drop them
      int nc = 0;
      // rows of the block -> (first slot, exclusive prefix of the row sizes) in shared memory; the candidates are then
      // walked as ONE flat sequence, 32 per step, so that short rows (a handful of points each) do not leave lanes idle
      int total = 0;
      const int nrows = (side * side + 31) & ~31;
      for (int t0 = 0; t0 < nrows; t0 += 32) {
        int a = 0, b = 0;
        const int t = t0 + lane;
        if (t < side * side) {
          const int z = cz - R + t / side, y = cy - R + t % side;
          if (z >= 0 && z < nz && y >= 0 && y < ny) { const int row = (z * ny + y) * nx; a = cs[row + x0]; b = cs[row + x1 + 1]; }
        }
        const int cnt = b - a;
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        s_ra[wib][t] = a;
        s_rp[wib][t] = total + inc - cnt;
        total += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) s_rp[wib][nrows] = total;
      __syncwarp();
      for (int t0 = 0; t0 < total; t0 += 32) {
        const int t = t0 + lane;
        double dd = INFINITY; int ii = 0x7fffffff, j = -1;
        if (t < total) {
          int lo = 0, hi = side * side;   // first index whose prefix exceeds t; s_rp[side * side] = total > t (the padding rows are empty)
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_rp[wib][mid] > t) hi = mid; else lo = mid + 1; }
          const int r = lo - 1;
          j = s_ra[wib][r] + (t - s_rp[wib][r]);
          const double4 p = pts[j];
          dd = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
          ii = (int)__double_as_longlong(p.w);
        }
        const bool ok = j >= 0 && dd < lim2;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        const int pos = nc + __popc(m & lt_mask);
        if (ok && pos < NS2_CAP) { s_d[wib][pos] = dd; s_i[wib][pos] = ii; s_s[wib][pos] = j; }
        nc += __popc(m);
      }
      __syncwarp();
      if (nc > NS2_CAP) { if (lane == 0) atomicAdd(queue_n + 2, 1); break; }   // too dense for the buffer: general kernel
      // ---- candidates -> registers (round-robin), statistics ----
      // nc is uniform over the warp, so is the number of 32-candidate chunks in use: every chunk loop below stops there
      // (the loops stay fully unrolled -- the register arrays need static indices -- but the unused tail is branched over)
      const int nch = (nc + 31) >> 5;
      // only the keys live in registers; a candidate's slot stays in the shared buffer until the cumulants need it, and its
      // histogram bin is recomputed where it is used (one multiply) -- registers, i.e. resident warps, are what this kernel is short of
      double d[NS2_CHUNKS]; int idx[NS2_CHUNKS];
#pragma unroll
      for (int c = 0; c < NS2_CHUNKS; c++) {
        if (c >= nch) break;
        const int t = c * 32 + lane;
        d[c] = INFINITY; idx[c] = 0x7fffffff;
        if (t < nc) { d[c] = s_d[wib][t]; idx[c] = s_i[wib][t]; }
      }
      __syncwarp();
      need = min(knn, nc);
      double td = INFINITY; int ti = 0x7fffffff;
      if (nc > knn) {   // k-th smallest (d2, index) by counting: 32-bin histogram over d2, then arg-min rounds in one bin
        // every candidate kept lies below lim2 (finite: lim2 <= radius^2), so that is the histogram's range -- no maximum to reduce
        const double scale = lim2 > 0.0 ? 32.0 / lim2 : 0.0;
        s_hist[wib][lane] = 0;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < NS2_CHUNKS; c++) {
          if (c >= nch) break;
          if (c * 32 + lane < nc) atomicAdd(&s_hist[wib][min(31, (int)(d[c] * scale))], 1);
        }
        __syncwarp();
        int cumh = s_hist[wib][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, cumh, o); if (lane >= o) cumh += t; }
        const int B = __ffs(__ballot_sync(0xffffffffu, cumh >= need)) - 1;
        const int below = B > 0 ? __shfl_sync(0xffffffffu, cumh, B - 1) : 0;
        const int m = need - below;
        // the m-th smallest (d2, index) of bin B by counting: its members go back to the (now free) candidate buffer, every lane ranks
        // one of them against all the others -- the keys are distinct, so exactly one member has rank m - 1
        int nb = 0;
#pragma unroll
        for (int c = 0; c < NS2_CHUNKS; c++) {
          if (c >= nch) break;
          const bool in_bin = c * 32 + lane < nc && min(31, (int)(d[c] * scale)) == B;
          const unsigned bm = __ballot_sync(0xffffffffu, in_bin);
          if (in_bin) { const int pos = nb + __popc(bm & lt_mask); s_d[wib][pos] = d[c]; s_i[wib][pos] = idx[c]; }
          nb += __popc(bm);
        }
        __syncwarp();
        for (int base = 0; base < nb; base += 32) {
          const int t = base + lane;
          double md = INFINITY; int mi = 0x7fffffff, rank = -1;
          if (t < nb) {
            md = s_d[wib][t]; mi = s_i[wib][t]; rank = 0;
            for (int u = 0; u < nb; u++) { const double od = s_d[wib][u]; const int oi = s_i[wib][u]; rank += (od < md || (od == md && oi < mi)) ? 1 : 0; }
          }
          const unsigned hit = __ballot_sync(0xffffffffu, rank == m - 1);
          if (hit) { const int src = __ffs(hit) - 1; td = __shfl_sync(0xffffffffu, md, src); ti = __shfl_sync(0xffffffffu, mi, src); break; }
        }
        __syncwarp();   // the buffer is written again by the next block radius
      }
      // (td, ti) = the k-th smallest key (inf when there are at most k candidates): the selected set is every key up to it
      // ---- exact?  every candidate kept lies strictly inside the guaranteed ball (radius sqrt(lim2) <= distance to the
      // nearest block face), so k kept candidates contain the true k nearest; fewer than k is final only when the
      // block covers the whole search radius ----
      if (!(nc >= knn || b2 > r2)) { if (last_try) break; continue; }   // grow the block
      // ---- cumulants of the selected candidates, butterfly sum over the warp ----
#pragma unroll
      for (int c = 0; c < NS2_CHUNKS; c++) {
        if (c >= nch) break;
        const int t = c * 32 + lane;
        if (t < nc && (d[c] < td || (d[c] == td && idx[c] <= ti))) {
          const double4 p = pts[s_s[wib][t]];
          c9[0] += p.x; c9[1] += p.y; c9[2] += p.z;
          c9[3] += p.x * p.x; c9[4] += p.x * p.y; c9[5] += p.x * p.z;
          c9[6] += p.y * p.y; c9[7] += p.y * p.z; c9[8] += p.z * p.z;
        }
      }
      warp_sum9_transposed(c9, lane);   // lane l now holds the warp total of cumulant l >> 1 in c9[0] (l < 18)
      resolved = true;
      if (lane == 0 && R > 1) atomicAdd(queue_n + 2 + min(R, 3), 1);   // statistics (B2S_DEBUG_NORMALS): resolved at R = 2 / at R >= 3
    }
    if (!resolved) {
      if (lane == 0) { queue[atomicAdd(queue_n, 1)] = s; cum[10 * (size_t)tq + 9] = -1.0; }
      continue;
    }
    {
      // lanes 0, 2, .., 16 hold one cumulant each (see warp_sum9_transposed), lane 18 writes the neighbour count
      const int slot = lane >> 1;
      if (!(lane & 1) && slot < 10) cum[10 * (size_t)tq + slot] = slot < 9 ? c9[0] : (double)need;
    }
  }
}

// eigen-solver + normalise + orient for the queries the select kernel resolved: one THREAD per query
__global__ void __launch_bounds__(NK_THREADS) normals_finish_kernel(const GridHeader* __restrict__ hdr, const double4* __restrict__ pts,
                                                                    const int32_t* __restrict__ qlist, const int32_t* __restrict__ qcount,
                                                                    const double* __restrict__ cum, double* __restrict__ out_nrm) {
  pdl_wait();
  const int nq = qlist ? *qcount : hdr->n;
  for (int tq = blockIdx.x * blockDim.x + threadIdx.x; tq < nq; tq += gridDim.x * blockDim.x) {
    const double kkd = cum[10 * (size_t)tq + 9];
    if (kkd < 0.0) continue;   // left to normals_phase2_kernel
    double c9[9];
#pragma unroll
    for (int t = 0; t < 9; t++) c9[t] = cum[10 * (size_t)tq + t];
    const double4 qp = pts[qlist ? qlist[tq] : tq];
    double nr[3];
    finish_normal(c9, (int)kkd, qp.x, qp.y, qp.z, nr);
    const size_t qi = (size_t)(int)__double_as_longlong(qp.w);
    out_nrm[3 * qi] = nr[0]; out_nrm[3 * qi + 1] = nr[1]; out_nrm[3 * qi + 2] = nr[2];
  }
}


__global__ void zero_i32_kernel(int32_t* p) {
  pdl_wait(); for (int i = 0; i < 6; i++) p[i] = 0; }

// query list = grid slots whose original point is flagged; warp-aggregated append keeps neighbouring slots together
__global__ void __launch_bounds__(NK_THREADS) normals_qlist_kernel(const GridHeader* __restrict__ hdr, const double4* __restrict__ pts,
                                                                   const int32_t* __restrict__ flags, int32_t* __restrict__ qlist,
                                                                   int32_t* qcount) {
  pdl_wait();
  const int n = hdr->n;
  const int lane = threadIdx.x & 31;
  const int n_round = (n + 31) & ~31;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_round; s += gridDim.x * blockDim.x) {
    bool take = false;
    if (s < n) take = flags[(int)__double_as_longlong(pts[s].w)] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, take);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(qcount, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) qlist[base + __popc(m & ((1u << lane) - 1u))] = s;
  }
}

int32_t op_estimate_normals(b2s_handle* h, b2s_cloud* c, int knn, double radius, double cell_hint, const int32_t* flags) {
  B2S_REQUIRE(radius > 0.0, B2S_E_INVALID, "maxRadiusNormalEstimation_ must be > 0");  // CloudRegistration.cpp:50
  B2S_REQUIRE(knn > 0, B2S_E_INVALID, "knnNormalEstimation_ must be > 0");            // CloudRegistration.cpp:51
  B2S_REQUIRE(knn <= 32, B2S_E_UNSUPPORTED, "knn > 32 is not supported by the register-resident k-best list yet");
  double cell = cell_hint > 0.0 ? cell_hint : radius / 4.0;
  if (cell < radius / 16.0) cell = radius / 16.0;  // bound the ring count of the worst case
  B2S_TRY(grid_build(h, &h->grid_b, c, cell, nullptr, false));
  const size_t n_max = c->n_max > 0 ? c->n_max : 1;
  B2S_TRY(c->nrm.ensure(n_max * 24, h->stream));
  int blocks = (int)((n_max + NK_THREADS - 1) / NK_THREADS);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  // phase-2 queue: counter + one slot per point
  B2S_TRY(h->tmp_i32.ensure((2 * n_max + 64) * 4, h->stream));
  int32_t* qn = h->tmp_i32.as<int32_t>() + 8;        // [0] phase-2 queue length, [1] query-list length
  int32_t* queue = h->tmp_i32.as<int32_t>() + 16;
  int32_t* qlist = flags ? queue + n_max : nullptr;
  const int32_t* qcount = flags ? qn + 1 : nullptr;
  // rings the thread-per-query kernel may walk before handing a query to the warp-cooperative kernel
  // (B2S_NORMALS_RING_LIMIT: tuning knob; -1 = every query goes to the warp-cooperative kernel)
  // measured on B200 (config 2, 11 k queries): all-warp 0.13 ms, thread kernel + warp stragglers 0.35 ms -> default -1
  static const int ring_limit_env = getenv("B2S_NORMALS_RING_LIMIT") ? atoi(getenv("B2S_NORMALS_RING_LIMIT")) : -1;
  const int ring_limit = ring_limit_env;
  ProfScope prof(h, PK_NORMALS);
  const GridHeader* hdr = h->grid_b.hdr.as<GridHeader>();
  const int32_t* cs = grid_starts(&h->grid_b);
  const double4* pts = h->grid_b.pts.as<double4>();
  double* out = c->nrm.as<double>();
  launch_pdl(zero_i32_kernel, 1, 1, 0, h->stream, qn);
  if (flags) {
    launch_pdl(normals_qlist_kernel, blocks, NK_THREADS, 0, h->stream, hdr, pts, flags, qlist, qn + 1);
    h->launches++;
  }
  // exact instantiations for the knn values the reference's presets use (Lua default 20, C++ struct default 5,
  // place-recognition normals 10); any other knn <= 32 takes the generic variants
  if (ring_limit == -1) {   // default: gather + select (one warp per query), stragglers to the general warp kernel
    // one warp per query, grid-stride (B2S_NS2_GRID: A/B knob of the grid size)
    static const int ns2_grid_env = getenv("B2S_NS2_GRID") ? atoi(getenv("B2S_NS2_GRID")) : 0;
    // measured at 16 chains / 1 chain: 2368 CTAs 10.21 k/s, 0.600 ms; 592: 10.28 k/s, 0.619 ms; 296: 10.41 k/s, 0.640 ms -> the large grid stays
    const int ns2_cap = ns2_grid_env > 0 ? ns2_grid_env : 148 * 16;
    int wblocks = (int)((n_max + (NK_THREADS / 32) - 1) / (NK_THREADS / 32));
    if (wblocks > ns2_cap) wblocks = ns2_cap;
    if (wblocks < 1) wblocks = 1;
    B2S_TRY(h->tmp_f64.ensure((n_max + 1) * 80, h->stream));
    double* cum = h->tmp_f64.as<double>();
    static const bool use_v1 = getenv("B2S_NORMALS_SELECT_V1") != nullptr;   // A/B knob: fixed 3x3x3 block, register gather
    if (use_v1) launch_pdl(normals_select_kernel, wblocks, NK_THREADS, 0, h->stream, hdr, cs, pts, knn, radius, qlist, qcount, queue, qn, cum);
    else launch_pdl(normals_select2_kernel, wblocks, NK_THREADS, 0, h->stream, hdr, cs, pts, knn, radius, qlist, qcount, queue, qn, cum);
    launch_pdl(normals_finish_kernel, blocks, NK_THREADS, 0, h->stream, hdr, pts, qlist, qcount, cum, out);
    h->launches++;
  } else if (knn == 20) launch_pdl(normals_kernel<20, true>, blocks, NK_THREADS, 0, h->stream, hdr, cs, pts, knn, radius, ring_limit, queue, qn, qlist, qcount, out);
  else if (knn == 10) launch_pdl(normals_kernel<10, true>, blocks, NK_THREADS, 0, h->stream, hdr, cs, pts, knn, radius, ring_limit, queue, qn, qlist, qcount, out);
  else if (knn == 5) launch_pdl(normals_kernel<5, true>, blocks, NK_THREADS, 0, h->stream, hdr, cs, pts, knn, radius, ring_limit, queue, qn, qlist, qcount, out);
  else if (knn <= 16) launch_pdl(normals_kernel<16, false>, blocks, NK_THREADS, 0, h->stream, hdr, cs, pts, knn, radius, ring_limit, queue, qn, qlist, qcount, out);
  else launch_pdl(normals_kernel<32, false>, blocks, NK_THREADS, 0, h->stream, hdr, cs, pts, knn, radius, ring_limit, queue, qn, qlist, qcount, out);
  static const bool dbg_counts = getenv("B2S_DEBUG_NORMALS") != nullptr;
  if (dbg_counts) {   // debug aid: how many queries the fast path left to the general kernel
    int32_t hq[6] = {0, 0, 0, 0, 0, 0};
    GridHeader gh;
    cudaMemcpyAsync(hq, qn, 24, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(&gh, hdr, sizeof(gh), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    fprintf(stderr, "[b2s normals] select2: over-capacity %d, resolved at R=2 %d, at R=3 %d\n", hq[2], hq[4], hq[5]);
    fprintf(stderr, "[b2s normals] indexed %d queries %d fallback %d cell %.3f dims %dx%dx%d\n", gh.n, flags ? hq[1] : gh.n, hq[0], gh.cell,
            gh.dims[0], gh.dims[1], gh.dims[2]);
  }
  launch_pdl(normals_phase2_kernel, 148 * 4, NK_THREADS, 0, h->stream, hdr, cs, pts, knn, radius, queue, qn, out);
  h->launches += 3;
  c->has_normals = true;
  B2S_CUDA(cudaGetLastError());
  return B2S_OK;
}

}  // namespace b2s
