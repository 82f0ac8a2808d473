// icp.cu -- K-icp-iter: the whole point-to-plane ICP loop of one registration inside ONE persistent kernel.
//
// Replaces [O3D] RegistrationICP + TransformationEstimationPointToPlane as called by
// RegistrationIcpPointToPlane::registerClouds (core/src/CloudRegistration.cpp:44-48), i.e. SURVEY.md 8a rows
// R3 (correspondence search), R4 (JtJ / Jtr), R5 (6x6 solve, SE(3) update, convergence test).
//
// Mapping to the machine:
//   * one thread-block CLUSTER (1..8 CTAs, one per SM) per registration, blockIdx.y = registration in the batch;
//   * the working copy of the source cloud lives in shared memory for the whole loop (each CTA owns a contiguous
//     chunk) and is advanced by the per-iteration update like [O3D] pcd.Transform(update);
//   * exact nearest neighbour with the strict d2 < r2 cut through the dense grid of grid_index.cu (ring expansion
//     with box-distance pruning; ties -> lower target index) -- gathers hit the L2-resident target;
//   * per-thread fp64 accumulation of the 21 + 6 + 2 sums, warp-shuffle tree, CTA tree, then a DSMEM exchange:
//     every CTA reads all cluster partials in rank order and redundantly solves the 6x6 system (LDLT with
//     diagonal pivoting), so one cluster barrier per iteration is enough and no host round trip ever happens;
//   * all arithmetic fp64; distances and the point transform use explicitly rounded ops (no FMA contraction) so
//     that correspondences are bit-identical to the CPU oracle.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b2s {

constexpr int ICP_THREADS = 512;
constexpr int ICP_WARPS = ICP_THREADS / 32;
constexpr int NACC = 29;  // 21 upper-triangular JtJ + 6 Jtr + sum d2 + count

struct GridView {
  double ox, oy, oz, cell, inv, eps;
  int nx, ny, nz;
  const int32_t* __restrict__ cs;
  const double4* __restrict__ pts;
  const double4* __restrict__ nrm;
};

__device__ __forceinline__ double slab_gap(double q, double o, double cell, int i, int n, double eps) {
  double g = 0.0;
  if (i > 0) { double lo = o + (double)i * cell; if (q < lo) g = lo - q; }
  if (i < n - 1) { double hi = o + (double)(i + 1) * cell; if (q > hi) g = q - hi; }
  g -= eps;  // slack: cell membership was decided with floor((p-o)*inv), which can disagree with o+i*cell by an ulp
  return g > 0.0 ? g : 0.0;
}

__device__ __forceinline__ void nn_scan_range(const double4* __restrict__ pts, int s, int e, double qx, double qy, double qz,
                                              double& best, int& bidx, int& bslot) {
  for (int j = s; j < e; ++j) {
    const double4 p = pts[j];
    const double d = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
    const int idx = (int)__double_as_longlong(p.w);
    if (d < best || (d == best && bslot >= 0 && idx < bidx)) { best = d; bidx = idx; bslot = j; }
  }
}

// exact nearest neighbour of q among the indexed points with d2 < r2 (strict). returns slot or -1.
__device__ __forceinline__ int nn_search(const GridView& g, double qx, double qy, double qz, double r2, double& d2_out) {
  if (!(qx == qx && qy == qy && qz == qz)) return -1;
  const int cx = (int)fmin(fmax(floor((qx - g.ox) * g.inv), 0.0), (double)(g.nx - 1));
  const int cy = (int)fmin(fmax(floor((qy - g.oy) * g.inv), 0.0), (double)(g.ny - 1));
  const int cz = (int)fmin(fmax(floor((qz - g.oz) * g.inv), 0.0), (double)(g.nz - 1));
  double best = r2;
  int bidx = 0x7fffffff, bslot = -1;
  for (int R = 0;; ++R) {
    const int z0 = max(cz - R, 0), z1 = min(cz + R, g.nz - 1);
    const int y0 = max(cy - R, 0), y1 = min(cy + R, g.ny - 1);
    const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
    for (int z = z0; z <= z1; ++z) {
      const double gz = slab_gap(qz, g.oz, g.cell, z, g.nz, g.eps);
      const double gz2 = gz * gz;
      if (gz2 > best) continue;
      const bool zface = (z == cz - R) || (z == cz + R);
      for (int y = y0; y <= y1; ++y) {
        const double gy = slab_gap(qy, g.oy, g.cell, y, g.ny, g.eps);
        if (gz2 + gy * gy > best) continue;
        const int row = (z * g.ny + y) * g.nx;
        if (zface || y == cy - R || y == cy + R) {
          nn_scan_range(g.pts, g.cs[row + x0], g.cs[row + x1 + 1], qx, qy, qz, best, bidx, bslot);
        } else {
          if (cx - R >= 0) nn_scan_range(g.pts, g.cs[row + cx - R], g.cs[row + cx - R + 1], qx, qy, qz, best, bidx, bslot);
          if (cx + R <= g.nx - 1) nn_scan_range(g.pts, g.cs[row + cx + R], g.cs[row + cx + R + 1], qx, qy, qz, best, bidx, bslot);
        }
      }
    }
    // every unvisited point lies beyond the faces of the (2R+1)^3 block that still have cells behind them
    double bound = INFINITY;
    if (cx - R > 0) bound = fmin(bound, qx - (g.ox + (double)(cx - R) * g.cell));
    if (cx + R < g.nx - 1) bound = fmin(bound, (g.ox + (double)(cx + R + 1) * g.cell) - qx);
    if (cy - R > 0) bound = fmin(bound, qy - (g.oy + (double)(cy - R) * g.cell));
    if (cy + R < g.ny - 1) bound = fmin(bound, (g.oy + (double)(cy + R + 1) * g.cell) - qy);
    if (cz - R > 0) bound = fmin(bound, qz - (g.oz + (double)(cz - R) * g.cell));
    if (cz + R < g.nz - 1) bound = fmin(bound, (g.oz + (double)(cz + R + 1) * g.cell) - qz);
    bound -= g.eps;
    if (bound < 0.0) bound = 0.0;
    if (bound == INFINITY || bound * bound > best) break;
  }
  d2_out = best;
  return bslot;
}

// ---- small fp64 linear algebra on one thread ----------------------------------------------------------------------
__device__ void ldlt6_solve_dev(const double* A_in, const double* b_in, double* x) {
  double A[36];
  for (int i = 0; i < 36; i++) A[i] = A_in[i];
  int perm[6] = {0, 1, 2, 3, 4, 5};
  for (int k = 0; k < 6; k++) {
    int piv = k;
    double best = fabs(A[7 * k]);
    for (int i = k + 1; i < 6; i++) if (fabs(A[7 * i]) > best) { best = fabs(A[7 * i]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < 6; j++) { double t = A[6 * k + j]; A[6 * k + j] = A[6 * piv + j]; A[6 * piv + j] = t; }
      for (int j = 0; j < 6; j++) { double t = A[6 * j + k]; A[6 * j + k] = A[6 * j + piv]; A[6 * j + piv] = t; }
      int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    const double d = A[7 * k];
    if (d != 0.0) {
      for (int i = k + 1; i < 6; i++) A[6 * i + k] /= d;
      for (int i = k + 1; i < 6; i++)
        for (int j = k + 1; j <= i; j++) {
          A[6 * i + j] -= A[6 * i + k] * d * A[6 * j + k];
          A[6 * j + i] = A[6 * i + j];
        }
    }
  }
  double y[6];
  for (int i = 0; i < 6; i++) y[i] = b_in[perm[i]];
  for (int i = 0; i < 6; i++) for (int j = 0; j < i; j++) y[i] -= A[6 * i + j] * y[j];
  for (int i = 0; i < 6; i++) { const double d = A[7 * i]; y[i] = (fabs(d) > 2.2250738585072014e-308) ? y[i] / d : 0.0; }
  for (int i = 5; i >= 0; i--) for (int j = i + 1; j < 6; j++) y[i] -= A[6 * j + i] * y[j];
  for (int i = 0; i < 6; i++) x[perm[i]] = y[i];
}

// [O3D] TransformVector6dToMatrix4d: R = Rz(x2) Ry(x1) Rx(x0), t = x[3..5]
__device__ void vec6_to_mat4_dev(const double* x, double* T) {
  double sa, ca, sb, cb, sg, cgm;
  sincos(x[0], &sa, &ca); sincos(x[1], &sb, &cb); sincos(x[2], &sg, &cgm);
  T[0] = cgm * cb; T[1] = cgm * sb * sa - sg * ca; T[2] = cgm * sb * ca + sg * sa; T[3] = x[3];
  T[4] = sg * cb;  T[5] = sg * sb * sa + cgm * ca; T[6] = sg * sb * ca - cgm * sa; T[7] = x[4];
  T[8] = -sb;      T[9] = cb * sa;                 T[10] = cb * ca;                T[11] = x[5];
  T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

__device__ void mat4_mul_dev(const double* A, const double* B, double* C) {
  double t[16];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += A[4 * i + k] * B[4 * k + j];
      t[4 * i + j] = s;
    }
  for (int i = 0; i < 16; i++) C[i] = t[i];
}

__device__ bool mat4_is_identity_dev(const double* T) {  // Eigen isIdentity(1e-12)
  const double prec = 1e-12;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      const double v = T[4 * i + j];
      if (i == j) { if (!(fabs(v - 1.0) <= prec * fmin(fabs(v), 1.0))) return false; }
      else if (!(fabs(v) <= prec)) return false;
    }
  return true;
}

constexpr int ICP_FIXED_SMEM_DOUBLES = ICP_WARPS * NACC + 2 * NACC + NACC + 16 + 16 + 8;

__global__ void __launch_bounds__(ICP_THREADS, 1) icp_p2plane_kernel(const IcpProblem* __restrict__ problems, int smem_pts_cap) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank();
  const unsigned csize = cluster.num_blocks();
  const IcpProblem& P = problems[blockIdx.y];

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_red = reinterpret_cast<double*>(smem_raw);  // [ICP_WARPS][NACC]
  double* s_part = s_red + ICP_WARPS * NACC;            // [2][NACC]  (read by the other CTAs through DSMEM)
  double* s_tot = s_part + 2 * NACC;                    // [NACC]
  double* s_U = s_tot + NACC;                           // [16] update of the current iteration
  double* s_T = s_U + 16;                               // [16] accumulated transformation
  double* s_misc = s_T + 16;                            // [0] prev fitness [1] prev rmse [2] done [3] apply
  GridHeader* s_g = reinterpret_cast<GridHeader*>(s_misc + 8);
  double* s_pts = reinterpret_cast<double*>(s_g + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = *P.src_n;
  const int chunk = (n + (int)csize - 1) / (int)csize;
  const int lo = min((int)crank * chunk, n), hi = min(lo + chunk, n);
  const int cnt = hi - lo;
  double* work = (cnt <= smem_pts_cap) ? s_pts : (P.work_xyz + 3 * (size_t)lo);

  if (tid == 0) {
    *s_g = *P.ghdr;
    const double* init = P.init_dev ? P.init_dev : P.init;
    for (int i = 0; i < 16; i++) { s_T[i] = init[i]; s_U[i] = init[i]; }
    s_misc[0] = 0.0; s_misc[1] = 0.0; s_misc[2] = 0.0;
    s_misc[3] = mat4_is_identity_dev(init) ? 0.0 : 1.0;  // [O3D]: if (!init.isIdentity()) pcd.Transform(init)
  }
  {  // stage this CTA's chunk of the source cloud
    const double* src = P.src_xyz + 3 * (size_t)lo;
    for (int i = tid; i < 3 * cnt; i += ICP_THREADS) work[i] = src[i];
  }
  __syncthreads();

  GridView g;
  g.ox = s_g->origin[0]; g.oy = s_g->origin[1]; g.oz = s_g->origin[2];
  g.cell = s_g->cell; g.inv = s_g->inv_cell; g.eps = 1e-9 * s_g->cell;
  g.nx = s_g->dims[0]; g.ny = s_g->dims[1]; g.nz = s_g->dims[2];
  g.cs = P.cell_start;
  g.pts = reinterpret_cast<const double4*>(P.tgt_pts);
  g.nrm = reinterpret_cast<const double4*>(P.tgt_nrm);
  const double r2 = P.max_corr * P.max_corr;
  const int max_iter = P.max_iter;

  for (int e = 0;; ++e) {
    const bool apply = s_misc[3] != 0.0;
    double U[12];
#pragma unroll
    for (int i = 0; i < 12; i++) U[i] = s_U[i];

    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0.0;

    for (int i = tid; i < cnt; i += ICP_THREADS) {
      double px = work[3 * i], py = work[3 * i + 1], pz = work[3 * i + 2];
      if (apply) {  // [O3D] TransformPoints with w == 1 exactly for a rigid update; same association as Eigen's product
        const double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(U[0], px), __dmul_rn(U[1], py)), __dmul_rn(U[2], pz)), U[3]);
        const double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(U[4], px), __dmul_rn(U[5], py)), __dmul_rn(U[6], pz)), U[7]);
        const double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(U[8], px), __dmul_rn(U[9], py)), __dmul_rn(U[10], pz)), U[11]);
        px = x; py = y; pz = z;
        work[3 * i] = px; work[3 * i + 1] = py; work[3 * i + 2] = pz;
      }
      double d2;
      const int slot = nn_search(g, px, py, pz, r2, d2);
      if (slot >= 0) {
        const double4 q = g.pts[slot];
        const double4 nn = g.nrm[slot];
        const double r = (px - q.x) * nn.x + (py - q.y) * nn.y + (pz - q.z) * nn.z;
        double J[6];
        J[0] = py * nn.z - pz * nn.y; J[1] = pz * nn.x - px * nn.z; J[2] = px * nn.y - py * nn.x;
        J[3] = nn.x; J[4] = nn.y; J[5] = nn.z;
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
          for (int b = a; b < 6; b++) acc[k++] += J[a] * J[b];
#pragma unroll
        for (int a = 0; a < 6; a++) acc[21 + a] += J[a] * r;
        acc[27] += d2;
        acc[28] += 1.0;
      }
    }
    // warp tree -> CTA tree (fixed order => run-to-run deterministic)
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      const double v = warp_sum(acc[i]);
      if (lane == 0) s_red[warp * NACC + i] = v;
    }
    __syncthreads();
    const int buf = e & 1;
    if (tid < NACC) {
      double v = 0.0;
      for (int w = 0; w < ICP_WARPS; w++) v += s_red[w * NACC + tid];
      s_part[buf * NACC + tid] = v;
    }
    cluster.sync();
    if (tid < NACC) {
      double v = 0.0;
      for (unsigned r = 0; r < csize; r++) {
        const double* rp = cluster.map_shared_rank(s_part, r);
        v += rp[buf * NACC + tid];
      }
      s_tot[tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
      const double c = s_tot[28];
      const double fit = (c > 0.0 && n > 0) ? c / (double)n : 0.0;
      const double rmse = c > 0.0 ? sqrt(s_tot[27] / c) : 0.0;
      bool done = false;
      if (e > 0 && fabs(s_misc[0] - fit) < P.rel_fitness && fabs(s_misc[1] - rmse) < P.rel_rmse) done = true;
      if (e >= max_iter) done = true;
      if (!done) {
        double Upd[16];
        if (c > 0.0) {
          double A[36], b[6], x[6];
          int k = 0;
          for (int a = 0; a < 6; a++) for (int bb = a; bb < 6; bb++) { A[6 * a + bb] = s_tot[k]; A[6 * bb + a] = s_tot[k]; k++; }
          for (int a = 0; a < 6; a++) b[a] = -s_tot[21 + a];
          ldlt6_solve_dev(A, b, x);
          vec6_to_mat4_dev(x, Upd);
        } else {  // [O3D] ComputeTransformation: corres.empty() -> Identity
          for (int i = 0; i < 16; i++) Upd[i] = (i % 5 == 0) ? 1.0 : 0.0;
        }
        mat4_mul_dev(Upd, s_T, s_T);
        for (int i = 0; i < 16; i++) s_U[i] = Upd[i];
        s_misc[0] = fit; s_misc[1] = rmse; s_misc[3] = 1.0;
      } else {
        s_misc[2] = 1.0;
        if (crank == 0) {
          b2s_result* out = P.out;
          for (int i = 0; i < 16; i++) out->T[i] = s_T[i];
          out->fitness = fit; out->inlier_rmse = rmse; out->n_corr = (int32_t)c; out->iters = e;
        }
      }
    }
    __syncthreads();
    if (s_misc[2] != 0.0) break;
  }
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

static bool g_icp_attr_set = false;
constexpr int ICP_DYN_SMEM = 200 * 1024;

int32_t icp_launch(b2s_handle* h, const IcpProblem* problems_dev, int n_problems, size_t max_src_points) {
  if (n_problems <= 0) return B2S_OK;
  if (!g_icp_attr_set) {
    B2S_CUDA(cudaFuncSetAttribute(icp_p2plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ICP_DYN_SMEM));
    g_icp_attr_set = true;
  }
  int csize = 1;
  while (csize < 8 && (size_t)csize * ICP_THREADS * 2 < max_src_points) csize *= 2;
  const int fixed = ICP_FIXED_SMEM_DOUBLES * 8 + (int)sizeof(GridHeader) + 16;
  const int pts_cap = (ICP_DYN_SMEM - fixed) / 24;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(csize, n_problems, 1);
  cfg.blockDim = dim3(ICP_THREADS, 1, 1);
  cfg.dynamicSmemBytes = ICP_DYN_SMEM;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof(h, PK_ICP);
  B2S_CUDA(cudaLaunchKernelEx(&cfg, icp_p2plane_kernel, problems_dev, pts_cap));
  h->launches++;
  return B2S_OK;
}

}  // namespace b2s
