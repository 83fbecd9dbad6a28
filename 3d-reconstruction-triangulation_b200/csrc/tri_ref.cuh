// tri_ref.cuh -- device functions that follow the reference's floating-point operation ORDER, for the
// results that feed exact compares (classifier gating, reference-LM trajectory).  Include only from
// translation units compiled with -fmad=false: every product and sum below must round separately,
// exactly as the reference's x86-64 build (and the oracle, built with -ffp-contract=off) does.
// IEEE division and sqrt are nvcc's defaults for double.
//
// Restated third-party arithmetic (OpenCV 4.x, modules/core/src/lapack.cpp and
// modules/calib3d/src/levmarq.cpp): JacobiImpl_ (cv::solve / cv::invert with DECOMP_EIG),
// SVBkSb back substitution, LMSolverImpl::run.
#pragma once
#include <float.h>

#include "tri_common.cuh"

namespace tri {
namespace ref {

// calculateRayDirectionForPixel + rotatePointByQuaternion (Triangulator.cpp:15-44); origin = rig.pos[c]
__device__ inline void make_dir(const RayRig& rig, int c, double px, double py, double dir[3]) {
  const double x0 = px + 0.5, y0 = py + 0.5;
  double vx = rig.aspect[c] * ((2 * x0 / rig.width[c]) - 1);
  double vy = (2 * y0 / rig.height[c]) - 1;
  double vz = rig.depth[c];
  const double s = 1.0 / sqrt(vx * vx + vy * vy + vz * vz);
  vx *= s; vy *= s; vz *= s;
  const double a1 = rig.quat[c][0], b1 = rig.quat[c][1], c1 = rig.quat[c][2], d1 = rig.quat[c][3];
  double a2 = 0.0, b2 = vx, c2 = vy, d2 = vz;
  const double ta = a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2;
  const double tb = a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2;
  const double tc = a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2;
  const double td = a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2;
  a2 = a1; b2 = -b1; c2 = -c1; d2 = -d1;
  dir[0] = ta * b2 + tb * a2 + tc * d2 - td * c2;
  dir[1] = ta * c2 - tb * d2 + tc * a2 + td * b2;
  dir[2] = ta * d2 + tb * c2 - tc * b2 + td * a2;
}

// distToRay (Triangulator.cpp:3-9)
__device__ __forceinline__ double dist_to_ray(const double o[3], const double d[3], double p0, double p1, double p2) {
  const double wx = p0 - o[0], wy = p1 - o[1], wz = p2 - o[2];
  const double cx = d[1] * wz - d[2] * wy;
  const double cy = d[2] * wx - d[0] * wz;
  const double cz = d[0] * wy - d[1] * wx;
  return sqrt(cx * cx + cy * cy + cz * cz);
}

// getDistFromRay (Triangulator.cpp:57-61)
__device__ inline double dist_from_ray(const RayRig& rig, int c, double px, double py, const double p[3]) {
  double d[3];
  make_dir(rig, c, px, py, d);
  return dist_to_ray(rig.pos[c], d, p[0], p[1], p[2]);
}

__device__ __forceinline__ double cv_hypot(double a, double b) {
  a = fabs(a); b = fabs(b);
  if (a > b) { b /= a; return a * sqrt(1 + b * b); }
  if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
  return 0;
}

// JacobiImpl_ for n = 3 (cv::eigen machinery behind DECOMP_EIG): A is destroyed; W descending; rows of V
__device__ inline void jacobi_eig3(double A[9], double W[3], double V[9]) {
  const int n = 3;
  const double eps = DBL_EPSILON;
  int i, j, k, m, iters, indR[3], indC[3];
  double mv = 0;
  for (i = 0; i < 9; i++) V[i] = 0;
  V[0] = V[4] = V[8] = 1;
  for (k = 0; k < n; k++) {
    W[k] = A[(n + 1) * k];
    if (k < n - 1) {
      for (m = k + 1, mv = fabs(A[n * k + m]), i = k + 2; i < n; i++) {
        double val = fabs(A[n * k + i]);
        if (mv < val) mv = val, m = i;
      }
      indR[k] = m;
    }
    if (k > 0) {
      for (m = 0, mv = fabs(A[k]), i = 1; i < k; i++) {
        double val = fabs(A[n * i + k]);
        if (mv < val) mv = val, m = i;
      }
      indC[k] = m;
    }
  }
  for (iters = 0; iters < n * n * 30; iters++) {
    for (k = 0, mv = fabs(A[indR[0]]), i = 1; i < n - 1; i++) {
      double val = fabs(A[n * i + indR[i]]);
      if (mv < val) mv = val, k = i;
    }
    int l = indR[k];
    for (i = 1; i < n; i++) {
      double val = fabs(A[n * indC[i] + i]);
      if (mv < val) mv = val, k = indC[i], l = i;
    }
    const double p = A[n * k + l];
    if (fabs(p) <= eps) break;
    const double y = (W[l] - W[k]) * 0.5;
    double t = fabs(y) + cv_hypot(p, y);
    double s = cv_hypot(p, t);
    const double c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) s = -s, t = -t;
    A[n * k + l] = 0;
    W[k] -= t;
    W[l] += t;
    double a0, b0;
#define TRI_ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
    for (i = 0; i < k; i++) TRI_ROT(A[n * i + k], A[n * i + l]);
    for (i = k + 1; i < l; i++) TRI_ROT(A[n * k + i], A[n * i + l]);
    for (i = l + 1; i < n; i++) TRI_ROT(A[n * k + i], A[n * l + i]);
    for (i = 0; i < n; i++) TRI_ROT(V[n * k + i], V[n * l + i]);
#undef TRI_ROT
    for (j = 0; j < 2; j++) {
      const int idx = j == 0 ? k : l;
      if (idx < n - 1) {
        for (m = idx + 1, mv = fabs(A[n * idx + m]), i = idx + 2; i < n; i++) {
          double val = fabs(A[n * idx + i]);
          if (mv < val) mv = val, m = i;
        }
        indR[idx] = m;
      }
      if (idx > 0) {
        for (m = 0, mv = fabs(A[idx]), i = 1; i < idx; i++) {
          double val = fabs(A[n * i + idx]);
          if (mv < val) mv = val, m = i;
        }
        indC[idx] = m;
      }
    }
  }
  for (k = 0; k < n - 1; k++) {
    m = k;
    for (i = k + 1; i < n; i++)
      if (W[m] < W[i]) m = i;
    if (k != m) {
      double t = W[m]; W[m] = W[k]; W[k] = t;
      for (i = 0; i < n; i++) { t = V[n * m + i]; V[n * m + i] = V[n * k + i]; V[n * k + i] = t; }
    }
  }
}

// cv::solve(A, b, x, DECOMP_EIG), symmetric 3x3
__device__ inline void eig_solve3(const double A_[9], const double b[3], double x[3]) {
  double A[9], W[3], V[9];
  for (int i = 0; i < 9; i++) A[i] = A_[i];
  jacobi_eig3(A, W, V);
  const double threshold = (W[0] + W[1] + W[2]) * (DBL_EPSILON * 2);
  x[0] = x[1] = x[2] = 0;
  for (int i = 0; i < 3; i++) {
    double wi = W[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    double s = 0;
    for (int j = 0; j < 3; j++) s += V[i * 3 + j] * b[j];
    s *= wi;
    for (int j = 0; j < 3; j++) x[j] = x[j] + s * V[i * 3 + j];
  }
}

// diag(cv::invert(A, DECOMP_EIG))
__device__ inline void eig_inv_diag3(const double A_[9], double diag[3]) {
  double A[9], W[3], V[9];
  for (int i = 0; i < 9; i++) A[i] = A_[i];
  jacobi_eig3(A, W, V);
  const double threshold = (W[0] + W[1] + W[2]) * (DBL_EPSILON * 2);
  diag[0] = diag[1] = diag[2] = 0;
  for (int i = 0; i < 3; i++) {
    double wi = W[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    for (int j = 0; j < 3; j++) diag[j] += V[i * 3 + j] * (V[i * 3 + j] * wi);
  }
}

// One view of a solve: the camera (for its origin) and the rotated direction.
struct RaySet {
  int n;
  int cam[TRI_MAX_CAMS];
  double d[TRI_MAX_CAMS][3];
};

// RayClosestPoint::compute without the Jacobian (RayTriangulator.cpp:16-26): S = sum r^2 with
// cv::norm(NORM_L2SQR)'s summation (groups of four in order, n%4 tail fused), max|r|, mean r.
__device__ inline void residual_pass(const RayRig& rig, const RaySet& rs, const double p[3], double& S, double& rmax,
                                     double& mean) {
  double s = 0, sum = 0, mx = 0;
  const int n = rs.n, k = n / 4 * 4;
  for (int i = 0; i < n; i++) {
    const double r = dist_to_ray(rig.pos[rs.cam[i]], rs.d[i], p[0], p[1], p[2]);
    sum += r;
    mx = fmax(mx, fabs(r));
    s = i < k ? s + r * r : fma(r, r, s);
  }
  S = s; rmax = mx; mean = sum / (double)n;
}

// compute() with the central-difference Jacobian (epsilon = THRESHOLD = 1e-4, RayTriangulator.cpp:28-44)
// folded straight into A = J^T J (cv::mulTransposed) and v = J^T r (cv::gemm GEMM_1_T: four partial
// sums over the rows, tail into the first) -- same summation order as the materialised version.
__device__ inline void normal_pass(const RayRig& rig, const RaySet& rs, const double p[3], double A[9], double v[3],
                                   double& S, double& rmax, double& mean) {
  const double e = 1e-4;
  const double x = p[0], y = p[1], z = p[2];
  double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
  double vs[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
  double s = 0, sum = 0, mx = 0;
  const int n = rs.n, k4 = n / 4 * 4;
  for (int i = 0; i < n; i++) {
    const double* o = rig.pos[rs.cam[i]];
    const double* d = rs.d[i];
    const double r = dist_to_ray(o, d, x, y, z);
    sum += r;
    mx = fmax(mx, fabs(r));
    s = i < k4 ? s + r * r : fma(r, r, s);
    const double j0 = (dist_to_ray(o, d, x + e, y, z) - dist_to_ray(o, d, x - e, y, z)) / (2 * e);
    const double j1 = (dist_to_ray(o, d, x, y + e, z) - dist_to_ray(o, d, x, y - e, z)) / (2 * e);
    const double j2 = (dist_to_ray(o, d, x, y, z + e) - dist_to_ray(o, d, x, y, z - e)) / (2 * e);
    a00 += j0 * j0; a01 += j0 * j1; a02 += j0 * j2; a11 += j1 * j1; a12 += j1 * j2; a22 += j2 * j2;
    const int slot = i < k4 ? (i & 3) : 0;
    vs[0][slot] += j0 * r; vs[1][slot] += j1 * r; vs[2][slot] += j2 * r;
  }
  A[0] = a00; A[1] = A[3] = a01; A[2] = A[6] = a02; A[4] = a11; A[5] = A[7] = a12; A[8] = a22;
  for (int a = 0; a < 3; a++) v[a] = ((vs[a][0] + vs[a][1]) + vs[a][2]) + vs[a][3];
  S = s; rmax = mx; mean = sum / (double)n;
}

// RayTriangulator::triangulatePoint (RayTriangulator.cpp:83-107) = LMSolverImpl::run from the mean of
// the origins, MAX_ITERATIONS = 1000, eps = FLT_EPSILON.  Returns the error left by the last compute().
__device__ inline double lm_point(const RayRig& rig, const RaySet& rs, double X[3], int& iters) {
  const int n = rs.n;
  double x[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) {
    const double* o = rig.pos[rs.cam[i]];
    x[0] += o[0]; x[1] += o[1]; x[2] += o[2];
  }
  x[0] /= n; x[1] /= n; x[2] /= n;
  double A[9], v[3], S, rmax, last_err;
  normal_pass(rig, rs, x, A, v, S, rmax, last_err);
  const double D[3] = {A[0], A[4], A[8]};
  double lambda = 1, lc = 0.75;
  int iter = 0;
  for (;;) {
    double Ap[9], d[3], xd[3];
    for (int i = 0; i < 9; i++) Ap[i] = A[i];
    Ap[0] += lambda * D[0]; Ap[4] += lambda * D[1]; Ap[8] += lambda * D[2];
    eig_solve3(Ap, v, d);
    for (int i = 0; i < 3; i++) xd[i] = x[i] - d[i];
    double Sd, rmax_d, err_d;
    residual_pass(rig, rs, xd, Sd, rmax_d, err_d);
    last_err = err_d;
    double temp_d[3];
    for (int i = 0; i < 3; i++)
      temp_d[i] = -1 * (A[i * 3] * d[0] + A[i * 3 + 1] * d[1] + A[i * 3 + 2] * d[2]) + 2 * v[i];
    const double dS = d[0] * temp_d[0] + d[1] * temp_d[1] + d[2] * temp_d[2];
    const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
    if (R > 0.75) {
      lambda *= 0.5;
      if (lambda < lc) lambda = 0;
    } else if (R < 0.25) {
      const double t = d[0] * v[0] + d[1] * v[1] + d[2] * v[2];
      double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
      nu = fmin(fmax(nu, 2.), 10.);
      if (lambda == 0) {
        double dg[3];
        eig_inv_diag3(A, dg);
        double maxval = DBL_EPSILON;
        for (int i = 0; i < 3; i++) maxval = fmax(maxval, fabs(dg[i]));
        lambda = lc = 1. / maxval;
        nu *= 0.5;
      }
      lambda *= nu;
    }
    if (Sd < S) {
      S = Sd;
      for (int i = 0; i < 3; i++) x[i] = xd[i];
      normal_pass(rig, rs, x, A, v, S, rmax, last_err);
      S = Sd;
    }
    iter++;
    const double dmax = fmax(fabs(d[0]), fmax(fabs(d[1]), fabs(d[2])));
    if (!(iter < 1000 && dmax >= (double)FLT_EPSILON && rmax >= (double)FLT_EPSILON)) break;
  }
  X[0] = x[0]; X[1] = x[1]; X[2] = x[2];
  iters = iter;
  return last_err;
}

// MatrixTriangulator::triangulatePoint (MatrixTriangulator.cpp:3-62) for an arbitrary camera subset:
// normal equations + adjugate, error from the residual rows themselves.
// (Pm[c] = camera c's 3x4 matrix, row-major: the rig's own copy in the constant bank, or a shared-memory copy where every
// thread indexes a different camera -- divergent constant-bank reads are serialised)
template <typename Rows>
__device__ inline double dlt_point_rows(const Rows& Pm, int n, const int* cam, const double* px, const double* py, double X[3]) {
  double M[6] = {0, 0, 0, 0, 0, 0}, v[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) {
    const double* P = Pm[cam[i]];
    const double x = px[i], y = py[i];
    double a0 = P[0] - x * P[8], a1 = P[1] - x * P[9], a2 = P[2] - x * P[10], b = x * P[11] - P[3];
    M[0] += a0 * a0; M[1] += a0 * a1; M[2] += a0 * a2; M[3] += a1 * a1; M[4] += a1 * a2; M[5] += a2 * a2;
    v[0] += a0 * b; v[1] += a1 * b; v[2] += a2 * b;
    a0 = P[4] - y * P[8]; a1 = P[5] - y * P[9]; a2 = P[6] - y * P[10]; b = y * P[11] - P[7];
    M[0] += a0 * a0; M[1] += a0 * a1; M[2] += a0 * a2; M[3] += a1 * a1; M[4] += a1 * a2; M[5] += a2 * a2;
    v[0] += a0 * b; v[1] += a1 * b; v[2] += a2 * b;
  }
  solve_sym3<double>(M, v, X);
  double ss = 0;
  for (int i = 0; i < n; i++) {
    const double* P = Pm[cam[i]];
    const double x = px[i], y = py[i];
    const double e0 = ((P[0] - x * P[8]) * X[0] + (P[1] - x * P[9]) * X[1] + (P[2] - x * P[10]) * X[2]) - (x * P[11] - P[3]);
    const double e1 = ((P[4] - y * P[8]) * X[0] + (P[5] - y * P[9]) * X[1] + (P[6] - y * P[10]) * X[2]) - (y * P[11] - P[7]);
    ss += e0 * e0;
    ss += e1 * e1;
  }
  return sqrt(ss / (2 * n));
}
__device__ inline double dlt_point(const DltRig<double>& rig, int n, const int* cam, const double* px, const double* py, double X[3]) {
  return dlt_point_rows(rig.P, n, cam, px, py, X);
}

}  // namespace ref
}  // namespace tri
