/* config1.c -- a plain C program against include/b2s.h and libb2s.so (no Python, no ctypes): config 1 of BASELINE.json.
 *
 * target: 2 000 points on three mutually orthogonal 10 m planes with analytic normals; source = target moved by yaw 3 deg,
 * pitch 1 deg, t = (0.10, -0.05, 0.02) (no noise): point-to-plane ICP (r = 1.0, <= 50 iterations, init = I) must return the
 * inverse displacement.  Exercises the boundary the way the reference-side shim does: b2s_create, clouds from host arrays
 * (std::vector<Eigen::Vector3d> layout), b2s_register, b2s_register_host, error codes and b2s_last_error.
 * Exit code 0 = pass.  Built and run by tests/test_c_abi.py (gcc -std=c99 ... -lb2s). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2s.h"

#define N 2000
#define CHECK(call)                                                                      \
  do {                                                                                   \
    int32_t rc_ = (call);                                                                \
    if (rc_ != B2S_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)rc_, b2s_last_error()); return 1; } \
  } while (0)

static unsigned long long rng_state = 88172645463325252ULL;
static double uniform01(void) { /* xorshift64* */
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return (double)((rng_state * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}

static void rot_zyx(double roll, double pitch, double yaw, double R[9]) {
  const double ca = cos(roll), sa = sin(roll), cb = cos(pitch), sb = sin(pitch), cg = cos(yaw), sg = sin(yaw);
  R[0] = cg * cb; R[1] = cg * sb * sa - sg * ca; R[2] = cg * sb * ca + sg * sa;
  R[3] = sg * cb; R[4] = sg * sb * sa + cg * ca; R[5] = sg * sb * ca - cg * sa;
  R[6] = -sb;     R[7] = cb * sa;                R[8] = cb * ca;
}

int main(void) {
  static double tgt[3 * N], nrm[3 * N], src[3 * N];
  for (int i = 0; i < N; i++) {
    const double a = 10.0 * uniform01(), b = 10.0 * uniform01();
    const int plane = i < N / 2 ? 0 : (i < 3 * N / 4 ? 1 : 2);
    double* p = tgt + 3 * i; double* n = nrm + 3 * i;
    n[0] = n[1] = n[2] = 0.0;
    if (plane == 0) { p[0] = a; p[1] = b; p[2] = 0; n[2] = 1; }
    else if (plane == 1) { p[0] = a; p[1] = 0; p[2] = b; n[1] = 1; }
    else { p[0] = 0; p[1] = a; p[2] = b; n[0] = 1; }
  }
  double R[9];
  const double deg = 3.14159265358979323846 / 180.0, t[3] = {0.10, -0.05, 0.02};
  rot_zyx(0.0, 1.0 * deg, 3.0 * deg, R);
  for (int i = 0; i < N; i++)
    for (int r = 0; r < 3; r++) src[3 * i + r] = R[3 * r] * tgt[3 * i] + R[3 * r + 1] * tgt[3 * i + 1] + R[3 * r + 2] * tgt[3 * i + 2] + t[r];

  if (b2s_device_count() <= 0) { fprintf(stderr, "no CUDA device: the engine has no CPU fallback\n"); return 77; }
  b2s_config cfg;
  b2s_default_config(&cfg);
  cfg.icp.max_corr_dist = 1.0; cfg.icp.max_iter = 50; cfg.icp.reg_type = B2S_REG_POINT_TO_PLANE;
  b2s_handle* h = NULL;
  CHECK(b2s_create(&cfg, 0, NULL, &h));
  b2s_cloud *cs = NULL, *ct = NULL;
  CHECK(b2s_cloud_create(h, &cs));
  CHECK(b2s_cloud_create(h, &ct));
  CHECK(b2s_cloud_upload_f64(h, cs, src, NULL, N));
  CHECK(b2s_cloud_upload_f64(h, ct, tgt, nrm, N));
  const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  b2s_result res, res2;
  CHECK(b2s_register(h, cs, ct, I, &res));
  CHECK(b2s_register_host(h, src, N, tgt, nrm, N, I, &res2));
  /* expected: T = inverse of (R, t): R^T, -R^T t */
  double worst = 0.0;
  for (int r = 0; r < 3; r++) {
    double et = 0.0;
    for (int c = 0; c < 3; c++) {
      const double e = fabs(res.T[4 * r + c] - R[3 * c + r]);
      if (e > worst) worst = e;
      et -= R[3 * c + r] * t[c];
    }
    if (fabs(res.T[4 * r + 3] - et) > worst) worst = fabs(res.T[4 * r + 3] - et);
  }
  double dmax = 0.0;
  for (int i = 0; i < 16; i++) if (fabs(res.T[i] - res2.T[i]) > dmax) dmax = fabs(res.T[i] - res2.T[i]);
  printf("config1 via the C ABI: iters %d, n_corr %d, fitness %.6f, rmse %.3e, max |T - T_true^-1| = %.3e, host-pointer form differs by %.1e\n",
         (int)res.iters, (int)res.n_corr, res.fitness, res.inlier_rmse, worst, dmax);
  int fail = 0;
  if (!(worst < 1e-6)) { fprintf(stderr, "transform off by %.3e\n", worst); fail = 1; }
  if (!(res.fitness > 0.999) || res.n_corr < N - 2 || res.iters < 2 || res.iters > 50) { fprintf(stderr, "unexpected fitness / counts\n"); fail = 1; }
  if (dmax > 1e-12 || res.iters != res2.iters) { fprintf(stderr, "b2s_register_host disagrees with b2s_register\n"); fail = 1; }   /* (outlier order: last bits) */
  /* error behaviour: point-to-plane without target normals -> B2S_E_NO_NORMALS with a message; r <= 0 -> B2S_E_INVALID */
  CHECK(b2s_cloud_upload_f64(h, ct, tgt, NULL, N));
  if (b2s_register(h, cs, ct, I, &res) != B2S_E_NO_NORMALS || strlen(b2s_last_error()) == 0) { fprintf(stderr, "missing-normals case not reported\n"); fail = 1; }
  cfg.icp.max_corr_dist = 0.0;
  CHECK(b2s_set_config(h, &cfg));
  CHECK(b2s_cloud_upload_f64(h, ct, tgt, nrm, N));
  if (b2s_register(h, cs, ct, I, &res) != B2S_E_INVALID) { fprintf(stderr, "invalid max_correspondence_distance not reported\n"); fail = 1; }
  b2s_cloud_destroy(cs); b2s_cloud_destroy(ct);
  b2s_destroy(h);
  printf(fail ? "FAIL\n" : "PASS\n");
  return fail;
}
