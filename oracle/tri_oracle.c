/* tri_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see tri_oracle.h).  Plain C, no OpenCV.  Citations are file:line under /root/reference/.
 *
 * Third-party arithmetic restated here (OpenCV 4.x, version unpinned by the reference's
 * CMakeLists.txt:4; restated from the published algorithms in modules/core/src/lapack.cpp and
 * modules/calib3d/src/levmarq.cpp and validated against the cv2 4.13 wheel by tests/):
 *   cv::invert(DECOMP_SVD)       -> orc_pinv_svd      (one-sided Jacobi SVD + back substitution)
 *   cv::solve/invert(DECOMP_EIG) -> jacobi_eig3 + svbksb (cyclic-pivot Jacobi eigen solver)
 *   cv::LMSolver::run            -> lm_run3
 *   cv::Mat::inv() 3x3           -> inv3 (closed-form adjugate branch of cv::invert)
 *
 * Compile with -ffp-contract=off: the CUDA "reference-LM" kernel is bit-compared to this file.
 */
#include "tri_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_CAMS 64

/* ---------------------------------------------------------------- P0: tdr::Camera ---- */

static void rot_from_quat(const double q[4], double R[9]) {
  /* Camera.h:274-287: raw (a,b,c,d) = (w,i,j,k); the normalised copy qTemp is never used. */
  double a = q[0], b = q[1], c = q[2], d = q[3];
  R[0] = 1 - 2 * (c * c + d * d); R[1] = 2 * (b * c - a * d);     R[2] = 2 * (b * d + a * c);
  R[3] = 2 * (b * c + a * d);     R[4] = 1 - 2 * (b * b + d * d); R[5] = 2 * (c * d - a * b);
  R[6] = 2 * (b * d - a * c);     R[7] = 2 * (c * d + a * b);     R[8] = 1 - 2 * (b * b + c * c);
}

static int inv3(const double S[9], double T[9]) {
  /* cv::Mat::inv() (DECOMP_LU) on a 3x3 CV_64F takes cv::invert's closed-form branch. */
#define S_(i, j) S[(i) * 3 + (j)]
  double d = S_(0, 0) * (S_(1, 1) * S_(2, 2) - S_(1, 2) * S_(2, 1)) -
             S_(0, 1) * (S_(1, 0) * S_(2, 2) - S_(1, 2) * S_(2, 0)) +
             S_(0, 2) * (S_(1, 0) * S_(2, 1) - S_(1, 1) * S_(2, 0));
  if (d == 0.) { memset(T, 0, 9 * sizeof(double)); return 0; }
  d = 1. / d;
  T[0] = (S_(1, 1) * S_(2, 2) - S_(1, 2) * S_(2, 1)) * d;
  T[1] = (S_(0, 2) * S_(2, 1) - S_(0, 1) * S_(2, 2)) * d;
  T[2] = (S_(0, 1) * S_(1, 2) - S_(0, 2) * S_(1, 1)) * d;
  T[3] = (S_(1, 2) * S_(2, 0) - S_(1, 0) * S_(2, 2)) * d;
  T[4] = (S_(0, 0) * S_(2, 2) - S_(0, 2) * S_(2, 0)) * d;
  T[5] = (S_(0, 2) * S_(1, 0) - S_(0, 0) * S_(1, 2)) * d;
  T[6] = (S_(1, 0) * S_(2, 1) - S_(1, 1) * S_(2, 0)) * d;
  T[7] = (S_(0, 1) * S_(2, 0) - S_(0, 0) * S_(2, 1)) * d;
  T[8] = (S_(0, 0) * S_(1, 1) - S_(0, 1) * S_(1, 0)) * d;
#undef S_
  return 1;
}

/* exported for oracle/shim (cv::Mat::inv of a 3x3) */
int orc_inv3(const double S[9], double T[9]) { return inv3(S, T); }

int orc_camera_make(orc_camera* c, int id, int width, int height, double focal, const double pos[3],
                    const double quat[4]) {
  /* createCamera, src/utils.cpp:94-107, then Camera::compCamParams, src/Camera.h:177-187 */
  const double RAD_TO_DEG = 57.29577951308232087679, DEG_TO_RAD = 0.01745329251994329576;
  memset(c, 0, sizeof(*c));
  c->id = id; c->width = width; c->height = height; c->focal = focal;
  memcpy(c->pos, pos, sizeof(c->pos));
  memcpy(c->quat, quat, sizeof(c->quat));
  if (width == 0 || height == 0) return ORC_ERR_CAMERA;           /* Camera.h:79-80 */
  c->cx = (int)round(width / 2.0);                                 /* Camera.h:81-82 */
  c->cy = (int)round(height / 2.0);
  c->fx = focal;
  if (c->fx == 0) return ORC_ERR_CAMERA;                           /* Camera.h:90 */
  c->fovx = 2 * atan(width / (2 * c->fx)) * 57.2958;               /* Camera.h:91 */
  c->fovy = 2.0 * atan(tan(c->fovx * 0.5 * DEG_TO_RAD) / ((double)width / (double)height)) * RAD_TO_DEG;
  c->fx = (width / 2.0) / (tan((c->fovx / 2.0) * DEG_TO_RAD));     /* Camera.h:114-115 */
  c->fy = (height / 2.0) / (tan((c->fovy / 2.0) * DEG_TO_RAD));
  double R0[9], R[9];
  rot_from_quat(quat, R0);
  inv3(R0, R);                                                     /* Camera.h:134,168 */
  for (int i = 0; i < 3; i++)                                      /* camPos = -R*tvec */
    c->cam_pos[i] = -(R[i * 3 + 0] * pos[0] + R[i * 3 + 1] * pos[1] + R[i * 3 + 2] * pos[2]);
  if (c->fx == 0 || c->fy == 0 || c->cx == 0 || c->cy == 0) return ORC_ERR_CAMERA; /* Camera.h:124 */
  double K[9] = {c->fx, 0, (double)c->cx, 0, c->fy, (double)c->cy, 0, 0, 1};
  memcpy(c->K, K, sizeof(K));
  for (int i = 0; i < 3; i++) {
    c->E[i * 4 + 0] = R[i * 3 + 0]; c->E[i * 4 + 1] = R[i * 3 + 1]; c->E[i * 4 + 2] = R[i * 3 + 2];
    c->E[i * 4 + 3] = c->cam_pos[i];
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += K[i * 3 + k] * c->E[k * 4 + j];
      c->P[i * 4 + j] = s;
    }
  return ORC_OK;
}

/* ------------------------------------------- cv::invert(DECOMP_SVD): one-sided Jacobi ---- */

static double cv_hypot(double a, double b) {
  /* lapack.cpp's own hypot */
  a = fabs(a); b = fabs(b);
  if (a > b) { b /= a; return a * sqrt(1 + b * b); }
  if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
  return 0;
}

void orc_pinv_svd(const double* A, int m, int n, double* pinv) {
  /* JacobiSVDImpl_ on At (n rows of length m = the columns of A), eps = 10*DBL_EPSILON;
   * then SVBkSb with threshold 2*DBL_EPSILON*sum(w): pinv = V diag(1/w) U^T. */
  double At[4 * 2 * ORC_MAX_CAMS], Vt[16], W[4];
  const double eps = DBL_EPSILON * 10;
  int i, j, k, iter, max_iter = m > 30 ? m : 30;
  for (i = 0; i < n; i++)
    for (k = 0; k < m; k++) At[i * m + k] = A[k * n + i];
  for (i = 0; i < n; i++) {
    double sd = 0;
    for (k = 0; k < m; k++) { double t = At[i * m + k]; sd += t * t; }
    W[i] = sd;
    for (k = 0; k < n; k++) Vt[i * n + k] = 0;
    Vt[i * n + i] = 1;
  }
  for (iter = 0; iter < max_iter; iter++) {
    int changed = 0;
    for (i = 0; i < n - 1; i++)
      for (j = i + 1; j < n; j++) {
        double *Ai = At + i * m, *Aj = At + j * m;
        double a = W[i], p = 0, b = W[j], c, s;
        for (k = 0; k < m; k++) p += Ai[k] * Aj[k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        double beta = a - b, gamma = cv_hypot(p, beta);
        if (beta < 0) {
          double delta = (gamma - beta) * 0.5;
          s = sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        a = b = 0;
        for (k = 0; k < m; k++) {
          double t0 = c * Ai[k] + s * Aj[k];
          double t1 = -s * Ai[k] + c * Aj[k];
          Ai[k] = t0; Aj[k] = t1;
          a += t0 * t0; b += t1 * t1;
        }
        W[i] = a; W[j] = b;
        changed = 1;
        double *Vi = Vt + i * n, *Vj = Vt + j * n;
        for (k = 0; k < n; k++) {
          double t0 = c * Vi[k] + s * Vj[k];
          double t1 = -s * Vi[k] + c * Vj[k];
          Vi[k] = t0; Vj[k] = t1;
        }
      }
    if (!changed) break;
  }
  for (i = 0; i < n; i++) {
    double sd = 0;
    for (k = 0; k < m; k++) { double t = At[i * m + k]; sd += t * t; }
    W[i] = sqrt(sd);
  }
  for (i = 0; i < n - 1; i++) {
    j = i;
    for (k = i + 1; k < n; k++)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      double t = W[i]; W[i] = W[j]; W[j] = t;
      for (k = 0; k < m; k++) { t = At[i * m + k]; At[i * m + k] = At[j * m + k]; At[j * m + k] = t; }
      for (k = 0; k < n; k++) { t = Vt[i * n + k]; Vt[i * n + k] = Vt[j * n + k]; Vt[j * n + k] = t; }
    }
  }
  for (i = 0; i < n; i++) { /* rows of At become the left singular vectors */
    double sd = W[i], s = sd > DBL_MIN ? 1 / sd : 0.;
    for (k = 0; k < m; k++) At[i * m + k] *= s;
  }
  double threshold = 0;
  for (i = 0; i < n; i++) threshold += W[i];
  threshold *= DBL_EPSILON * 2;
  for (i = 0; i < n * m; i++) pinv[i] = 0;
  for (i = 0; i < n; i++) {
    double wi = W[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    for (j = 0; j < n; j++) {
      double vj = Vt[i * n + j];
      for (k = 0; k < m; k++) pinv[j * m + k] += vj * (At[i * m + k] * wi);
    }
  }
}

/* ------------------------------------------------ cv::solve / cv::invert (DECOMP_EIG) ---- */

static void jacobi_eig3(double A[9], double W[3], double V[9]) {
  /* JacobiImpl_ (lapack.cpp) for n = 3: pivot = largest off-diagonal element tracked through
   * indR/indC, <= n*n*30 rotations, eigenvalues sorted descending; V rows are eigenvectors. */
  const int n = 3;
  const double eps = DBL_EPSILON;
  int i, j, k, m, iters, maxIters = n * n * 30, indR[3], indC[3];
  double mv = 0;
  for (i = 0; i < n; i++) { for (j = 0; j < n; j++) V[i * n + j] = 0; V[i * n + i] = 1; }
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
  for (iters = 0; iters < maxIters; iters++) {
    for (k = 0, mv = fabs(A[indR[0]]), i = 1; i < n - 1; i++) {
      double val = fabs(A[n * i + indR[i]]);
      if (mv < val) mv = val, k = i;
    }
    int l = indR[k];
    for (i = 1; i < n; i++) {
      double val = fabs(A[n * indC[i] + i]);
      if (mv < val) mv = val, k = indC[i], l = i;
    }
    double p = A[n * k + l];
    if (fabs(p) <= eps) break;
    double y = (W[l] - W[k]) * 0.5;
    double t = fabs(y) + cv_hypot(p, y);
    double s = cv_hypot(p, t);
    double c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) s = -s, t = -t;
    A[n * k + l] = 0;
    W[k] -= t;
    W[l] += t;
    double a0, b0;
#define ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
    for (i = 0; i < k; i++) ROT(A[n * i + k], A[n * i + l]);
    for (i = k + 1; i < l; i++) ROT(A[n * k + i], A[n * i + l]);
    for (i = l + 1; i < n; i++) ROT(A[n * k + i], A[n * l + i]);
    for (i = 0; i < n; i++) ROT(V[n * k + i], V[n * l + i]);
#undef ROT
    for (j = 0; j < 2; j++) {
      int idx = j == 0 ? k : l;
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

void orc_eig_solve3(const double A_[9], const double b[3], double x[3]) {
  /* cv::solve(DECOMP_EIG): Jacobi, then SVBkSb with u = v = eigenvectors, nb = 1 */
  double A[9], W[3], V[9];
  memcpy(A, A_, sizeof(A));
  jacobi_eig3(A, W, V);
  double threshold = (W[0] + W[1] + W[2]) * (DBL_EPSILON * 2);
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

void orc_eig_inv_diag3(const double A_[9], double diag[3]) {
  /* diag(cv::invert(A, DECOMP_EIG)): SVBkSb with b = identity */
  double A[9], W[3], V[9];
  memcpy(A, A_, sizeof(A));
  jacobi_eig3(A, W, V);
  double threshold = (W[0] + W[1] + W[2]) * (DBL_EPSILON * 2);
  diag[0] = diag[1] = diag[2] = 0;
  for (int i = 0; i < 3; i++) {
    double wi = W[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    for (int j = 0; j < 3; j++) diag[j] += V[i * 3 + j] * (V[i * 3 + j] * wi);
  }
}

/* ------------------------------------------ K2a-c: ray helpers (src/Triangulator.cpp) ---- */

void orc_make_ray(const orc_camera* cam, double px, double py, double origin[3], double dir[3]) {
  /* calculateRayDirectionForPixel, Triangulator.cpp:27-44 */
  double x0 = px + 0.5, y0 = py + 0.5;
  double d = 1 / tan(cam->fovy * 0.0174533 / 2);
  double vx = ((double)cam->width / (double)cam->height) * ((2 * x0 / (double)cam->width) - 1);
  double vy = (2 * y0 / (double)cam->height) - 1;
  double vz = d;
  double s = 1.0 / sqrt(vx * vx + vy * vy + vz * vz); /* cv::normalize(Vec3d) */
  vx *= s; vy *= s; vz *= s;
  /* rotatePointByQuaternion, Triangulator.cpp:15-25: q*(0,v)*conj(q), Hamilton, w first */
  double a1 = cam->quat[0], b1 = cam->quat[1], c1 = cam->quat[2], d1 = cam->quat[3];
  double a2 = 0.0, b2 = vx, c2 = vy, d2 = vz;
  double ta = a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2;
  double tb = a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2;
  double tc = a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2;
  double td = a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2;
  a2 = a1; b2 = -b1; c2 = -c1; d2 = -d1;
  dir[0] = ta * b2 + tb * a2 + tc * d2 - td * c2;
  dir[1] = ta * c2 - tb * d2 + tc * a2 + td * b2;
  dir[2] = ta * d2 + tb * c2 - tc * b2 + td * a2;
  origin[0] = cam->pos[0]; origin[1] = cam->pos[1]; origin[2] = cam->pos[2]; /* :50-54 */
}

double orc_dist_to_ray(const double o[3], const double d[3], const double p[3]) {
  /* distToRay, Triangulator.cpp:3-9 */
  double wx = p[0] - o[0], wy = p[1] - o[1], wz = p[2] - o[2];
  double cx = d[1] * wz - d[2] * wy;
  double cy = d[2] * wx - d[0] * wz;
  double cz = d[0] * wy - d[1] * wx;
  return sqrt(cx * cx + cy * cy + cz * cz);
}

double orc_dist_from_ray(const orc_camera* cam, double x, double y, const double p[3]) {
  double o[3], d[3]; /* getDistFromRay, Triangulator.cpp:57-61 */
  orc_make_ray(cam, x, y, o, d);
  return orc_dist_to_ray(o, d, p);
}

/* ------------------------------------- K1a: MatrixTriangulator::triangulatePoint ---- */

double orc_matrix_point(const orc_camera* cams, int n, const int* cam_idx, const double* xy, double X[3]) {
  /* MatrixTriangulator.cpp:3-62 */
  double A[2 * ORC_MAX_CAMS * 3], b[2 * ORC_MAX_CAMS], pinv[3 * 2 * ORC_MAX_CAMS];
  for (int i = 0; i < n; i++) {
    const double* P = cams[cam_idx[i]].P;
    double x = xy[2 * i], y = xy[2 * i + 1];
    for (int k = 0; k < 3; k++) {
      A[(2 * i) * 3 + k] = P[0 * 4 + k] - x * P[2 * 4 + k];
      A[(2 * i + 1) * 3 + k] = P[1 * 4 + k] - y * P[2 * 4 + k];
    }
    b[2 * i] = x * P[2 * 4 + 3] - P[0 * 4 + 3];
    b[2 * i + 1] = y * P[2 * 4 + 3] - P[1 * 4 + 3];
  }
  int m = 2 * n;
  orc_pinv_svd(A, m, 3, pinv);
  for (int j = 0; j < 3; j++) {
    double s = 0;
    for (int k = 0; k < m; k++) s += pinv[j * m + k] * b[k];
    X[j] = s;
  }
  double ss = 0;
  for (int k = 0; k < m; k++) {
    double e = (A[k * 3] * X[0] + A[k * 3 + 1] * X[1] + A[k * 3 + 2] * X[2]) - b[k];
    ss += e * e;
  }
  return sqrt(ss / (2 * n));
}

/* ---------------------------- K2d/e: RayClosestPoint::compute + cv::LMSolver::run ---- */

typedef struct { int n; double o[ORC_MAX_CAMS][3], d[ORC_MAX_CAMS][3]; double last_err; } ray_set;

static void rays_residual(ray_set* rs, const double p[3], double* r) {
  /* RayTriangulator.cpp:16-26; summation in index order (the reference's forEach is racy) */
  double s = 0;
  for (int i = 0; i < rs->n; i++) { r[i] = orc_dist_to_ray(rs->o[i], rs->d[i], p); s += r[i]; }
  rs->last_err = s / (double)rs->n;
}

static void rays_jacobian(const ray_set* rs, const double p[3], double* J) {
  /* central differences, epsilon = THRESHOLD = 1e-4, RayTriangulator.cpp:28-44 */
  const double e = 1e-4;
  double x = p[0], y = p[1], z = p[2];
  for (int i = 0; i < rs->n; i++) {
    double a[3], b[3];
    a[0] = x + e; a[1] = y; a[2] = z; b[0] = x - e; b[1] = y; b[2] = z;
    J[i * 3 + 0] = (orc_dist_to_ray(rs->o[i], rs->d[i], a) - orc_dist_to_ray(rs->o[i], rs->d[i], b)) / (2 * e);
    a[0] = x; a[1] = y + e; b[0] = x; b[1] = y - e;
    J[i * 3 + 1] = (orc_dist_to_ray(rs->o[i], rs->d[i], a) - orc_dist_to_ray(rs->o[i], rs->d[i], b)) / (2 * e);
    a[1] = y; a[2] = z + e; b[1] = y; b[2] = z - e;
    J[i * 3 + 2] = (orc_dist_to_ray(rs->o[i], rs->d[i], a) - orc_dist_to_ray(rs->o[i], rs->d[i], b)) / (2 * e);
  }
}

static void normal_eq(int n, const double* J, const double* r, double A[9], double v[3]) {
  /* mulTransposed(J, A, true); gemm(J, r, 1, noArray(), 0, v, GEMM_1_T) */
  for (int a = 0; a < 3; a++) {
    for (int b = a; b < 3; b++) {
      double s = 0;
      for (int i = 0; i < n; i++) s += J[i * 3 + a] * J[i * 3 + b];
      A[a * 3 + b] = A[b * 3 + a] = s;
    }
    /* cv::gemm(GEMM_1_T) on (n x 3)^T (n x 1): four partial sums over k (GEMMSingleMul's unrolled
     * loop), the k tail goes to the first one; probed against cv2 4.13 (tests/test_oracle_pin.py) */
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
      s0 += J[i * 3 + a] * r[i];
      s1 += J[(i + 1) * 3 + a] * r[i + 1];
      s2 += J[(i + 2) * 3 + a] * r[i + 2];
      s3 += J[(i + 3) * 3 + a] * r[i + 3];
    }
    for (; i < n; i++) s0 += J[i * 3 + a] * r[i];
    v[a] = ((s0 + s1) + s2) + s3;
  }
}

static double sumsq(int n, const double* r) {
  /* cv::norm(r, NORM_L2SQR): as probed on cv2 4.13 (AVX2 build) for n <= 15 the groups of four are
   * summed in order without contraction and the n%4 tail is fused (fma); from n = 16 a SIMD block
   * changes the order again -- not restated (only the last bits of S differ). */
  double s = 0;
  int i = 0, k = n / 4 * 4;
  for (; i < k; i++) s += r[i] * r[i];
  for (; i < n; i++) s = fma(r[i], r[i], s);
  return s;
}
static double maxabs(int n, const double* r) { double s = 0; for (int i = 0; i < n; i++) { double a = fabs(r[i]); if (a > s) s = a; } return s; }

static int lm_run3(ray_set* rs, double x[3], int maxIters, double eps) {
  /* LMSolverImpl::run (calib3d/src/levmarq.cpp), 3 parameters */
  double r[ORC_MAX_CAMS], rd[ORC_MAX_CAMS], J[ORC_MAX_CAMS * 3];
  double A[9], Ap[9], v[3], D[3], d[3], xd[3], temp_d[3];
  int n = rs->n, iter = 0;
  rays_residual(rs, x, r);
  rays_jacobian(rs, x, J);
  double S = sumsq(n, r);
  normal_eq(n, J, r, A, v);
  D[0] = A[0]; D[1] = A[4]; D[2] = A[8];
  const double Rlo = 0.25, Rhi = 0.75;
  double lambda = 1, lc = 0.75;
  for (;;) {
    memcpy(Ap, A, sizeof(A));
    for (int i = 0; i < 3; i++) Ap[i * 3 + i] += lambda * D[i];
    orc_eig_solve3(Ap, v, d);
    for (int i = 0; i < 3; i++) xd[i] = x[i] - d[i];
    rays_residual(rs, xd, rd);
    double Sd = sumsq(n, rd);
    for (int i = 0; i < 3; i++) /* gemm(A, d, -1, v, 2, temp_d) */
      temp_d[i] = -1 * (A[i * 3] * d[0] + A[i * 3 + 1] * d[1] + A[i * 3 + 2] * d[2]) + 2 * v[i];
    double dS = d[0] * temp_d[0] + d[1] * temp_d[1] + d[2] * temp_d[2];
    double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
    if (R > Rhi) {
      lambda *= 0.5;
      if (lambda < lc) lambda = 0;
    } else if (R < Rlo) {
      double t = d[0] * v[0] + d[1] * v[1] + d[2] * v[2];
      double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
      nu = fmin(fmax(nu, 2.), 10.);
      if (lambda == 0) {
        double dg[3];
        orc_eig_inv_diag3(A, dg);
        double maxval = DBL_EPSILON;
        for (int i = 0; i < 3; i++) maxval = fmax(maxval, fabs(dg[i]));
        lambda = lc = 1. / maxval;
        nu *= 0.5;
      }
      lambda *= nu;
    }
    if (Sd < S) {
      S = Sd;
      for (int i = 0; i < 3; i++) { double t = x[i]; x[i] = xd[i]; xd[i] = t; }
      rays_residual(rs, x, r);
      rays_jacobian(rs, x, J);
      normal_eq(n, J, r, A, v);
    }
    iter++;
    int proceed = iter < maxIters && maxabs(3, d) >= eps && maxabs(n, r) >= eps;
    if (!proceed) break;
  }
  return iter;
}

static void build_rays(const orc_camera* cams, int n, const int* cam_idx, const double* xy, ray_set* rs) {
  rs->n = n;
  for (int i = 0; i < n; i++) orc_make_ray(&cams[cam_idx[i]], xy[2 * i], xy[2 * i + 1], rs->o[i], rs->d[i]);
}

double orc_ray_point(const orc_camera* cams, int n, const int* cam_idx, const double* xy, double X[3],
                     int* iters) {
  /* RayTriangulator::triangulatePoint, RayTriangulator.cpp:83-107 */
  ray_set rs;
  build_rays(cams, n, cam_idx, xy, &rs);
  double g[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) { g[0] += rs.o[i][0]; g[1] += rs.o[i][1]; g[2] += rs.o[i][2]; }
  g[0] /= n; g[1] /= n; g[2] /= n;
  int it = lm_run3(&rs, g, 1000 /* MAX_ITERATIONS */, (double)FLT_EPSILON);
  if (iters) *iters = it;
  X[0] = g[0]; X[1] = g[1]; X[2] = g[2];
  return rs.last_err; /* value left by the last compute() call */
}

void orc_ray_closed_form(const orc_camera* cams, int n, const int* cam_idx, const double* xy, double X[3]) {
  ray_set rs;
  build_rays(cams, n, cam_idx, xy, &rs);
  double M[9] = {0}, c[3] = {0}, Mi[9];
  for (int i = 0; i < n; i++) {
    const double* d = rs.d[i];
    double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) { Mi[a * 3 + b] = (a == b ? dd : 0) - d[a] * d[b]; M[a * 3 + b] += Mi[a * 3 + b]; }
    for (int a = 0; a < 3; a++) c[a] += Mi[a * 3] * rs.o[i][0] + Mi[a * 3 + 1] * rs.o[i][1] + Mi[a * 3 + 2] * rs.o[i][2];
  }
  double Minv[9];
  inv3(M, Minv);
  for (int a = 0; a < 3; a++) X[a] = Minv[a * 3] * c[0] + Minv[a * 3 + 1] * c[1] + Minv[a * 3 + 2] * c[2];
}

/* -------------------------------------------------- K1b / K2f: triangulatePoints ---- */

static int tri_points_impl(const orc_camera* cams, int n_cams, int n_point_cams, int mode, const double* xyd,
                           const float* xyf, int64_t n_frames, int allow_too_few, double* out_xyz,
                           double* out_err, uint32_t* out_mask, int32_t* out_iters, int nthreads) {
  /* MatrixTriangulator.cpp:70-100 loops n_cam < min(points.size(), cameras.size());
   * RayTriangulator.cpp:51-81 loops n_cam < points.size() (OOB if more point rows than cameras). */
  int use = n_point_cams;
  if (mode == ORC_MATRIX && n_cams < use) use = n_cams;
  if (use > n_cams || use > ORC_MAX_CAMS) return ORC_ERR_DIM;
  int status = ORC_OK;
  (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int64_t f = 0; f < n_frames; f++) {
    int idx[ORC_MAX_CAMS], n = 0;
    double pix[2 * ORC_MAX_CAMS];
    uint32_t mask = 0;
    for (int c = 0; c < use; c++) {
      double x = xyd ? xyd[((int64_t)c * n_frames + f) * 2] : (double)xyf[((int64_t)c * n_frames + f) * 2];
      double y = xyd ? xyd[((int64_t)c * n_frames + f) * 2 + 1] : (double)xyf[((int64_t)c * n_frames + f) * 2 + 1];
      if (x == -1 || y == -1) continue;
      idx[n] = c; pix[2 * n] = x; pix[2 * n + 1] = y; n++;
      if (c < 32) mask |= 1u << c;
    }
    double X[3] = {0, 0, 0}, err = 0;
    int it = 0;
    if (n < 2) {
      if (!allow_too_few) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
        status = ORC_ERR_TOO_FEW;
      }
    } else if (mode == ORC_MATRIX) {
      err = orc_matrix_point(cams, n, idx, pix, X);
    } else {
      err = orc_ray_point(cams, n, idx, pix, X, &it);
    }
    out_xyz[3 * f] = X[0]; out_xyz[3 * f + 1] = X[1]; out_xyz[3 * f + 2] = X[2];
    if (out_err) out_err[f] = err;
    if (out_mask) out_mask[f] = mask;
    if (out_iters) out_iters[f] = it;
  }
  return status;
}

int orc_triangulate_points(const orc_camera* cams, int n_cams, int n_point_cams, int mode, const double* xy,
                           int64_t n_frames, int allow_too_few, double* out_xyz, double* out_err,
                           uint32_t* out_mask, int32_t* out_iters, int nthreads) {
  return tri_points_impl(cams, n_cams, n_point_cams, mode, xy, NULL, n_frames, allow_too_few, out_xyz, out_err,
                         out_mask, out_iters, nthreads);
}

int orc_triangulate_points_f32(const orc_camera* cams, int n_cams, int n_point_cams, int mode,
                               const float* xy, int64_t n_frames, int allow_too_few, double* out_xyz,
                               double* out_err, uint32_t* out_mask, int32_t* out_iters, int nthreads) {
  return tri_points_impl(cams, n_cams, n_point_cams, mode, NULL, xy, n_frames, allow_too_few, out_xyz, out_err,
                         out_mask, out_iters, nthreads);
}

/* ----------------------------------------------------------- K3: DroneClassifier ---- */

#define MAX_ERROR_MATRIX 1e+5 /* DroneClassifier.h:11 */
#define MAX_ERROR_RAY 120     /* :12 */
#define MAX_STEP 200          /* :13 */
#define MIN_CAMERAS 2         /* :16 */
#define PATH_TAIL 3           /* :17 */

typedef struct { int8_t c[ORC_MAX_CAMS]; double p[3]; double err; int zeros; int order; } comb_t;
typedef struct { comb_t* v; int n, cap; } comb_vec;

static void cv_push(comb_vec* q, const comb_t* c) {
  if (q->n == q->cap) { q->cap = q->cap ? q->cap * 2 : 256; q->v = (comb_t*)realloc(q->v, (size_t)q->cap * sizeof(comb_t)); }
  q->v[q->n++] = *c;
}

/* Combination::operator< (DroneClassifier.cpp:12-20): a < b  <=>  a has lower priority */
static int comb_less(const comb_t* a, const comb_t* b) {
  if (a->zeros != b->zeros) return a->zeros > b->zeros;
  return a->err > b->err;
}

/* libstdc++ std::priority_queue = std::push_heap / std::pop_heap on a vector (bits/stl_heap.h) */
static void heap_push_up(comb_t* first, int hole, int top, const comb_t* value) {
  int parent = (hole - 1) / 2;
  while (hole > top && comb_less(&first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = *value;
}
static void pq_push(comb_vec* q, const comb_t* c) {
  cv_push(q, c);
  comb_t value = q->v[q->n - 1];
  heap_push_up(q->v, q->n - 1, 0, &value);
}
static void pq_pop(comb_vec* q) {
  if (q->n > 1) {
    comb_t value = q->v[q->n - 1];
    q->v[q->n - 1] = q->v[0];
    int len = q->n - 1, hole = 0, top = 0, second = 0;
    while (second < (len - 1) / 2) {
      second = 2 * (second + 1);
      if (comb_less(&q->v[second], &q->v[second - 1])) second--;
      q->v[hole] = q->v[second];
      hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
      second = 2 * (second + 1);
      q->v[hole] = q->v[second - 1];
      hole = second - 1;
    }
    heap_push_up(q->v, hole, top, &value);
  }
  q->n--;
}

/* Margin audit: the smallest |lhs - rhs| of every floating-point compare that decides an index --
 * a flip needs an arithmetic difference of at least that much.  [0] error vs error_ (DroneClassifier.cpp:185,209,243),
 * [1] |c.point - pos| vs MAX_STEP (:244), [2] getDistFromRay vs MAX_STEP (:231-233), [3] error vs error between
 * queue entries with the same number of unused cameras (Combination::operator<, :12-20), [4] the tail distances
 * compared in classifyPaths (:291, :299-300). */
static _Thread_local double g_margin[5];
static void margin(int k, double lhs, double rhs) { double d = fabs(lhs - rhs); if (d < g_margin[k]) g_margin[k] = d; }
void orc_last_margins(double out[5]) { memcpy(out, g_margin, sizeof(g_margin)); }
static int cmp_zero_err(const void* a, const void* b) {
  const double* x = (const double*)a; const double* y = (const double*)b;
  if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
  return x[1] < y[1] ? -1 : x[1] > y[1] ? 1 : 0;
}

typedef struct {
  const orc_camera* cams; int n_cams, mode, n_drones; double error_;
  orc_stats st;
} clf_t;

/* One frame's detections: per camera count + pixel list + map local index -> original index */
typedef struct { int n[ORC_MAX_CAMS]; const double* xy[ORC_MAX_CAMS]; double gated_xy[ORC_MAX_CAMS][2 * 127];
                 int remap[ORC_MAX_CAMS][128]; int use_remap; } frame_dets;

/* Iterator, DroneClassifier.cpp:43-89 */
typedef struct { int c[ORC_MAX_CAMS], sizes[ORC_MAX_CAMS], n, skip; } iter_t;
static int it_increment(iter_t* it) {
  if (it->skip) { it->skip = 0; return 1; }
  for (int i = 0; i < it->n; i++) if (it->c[i] == -1) { it->c[i]++; return 1; }
  for (int i = it->n - 1; i >= 0; i--) {
    if (it->c[i] < it->sizes[i] - 1) { it->c[i]++; return 1; }
    it->c[i] = -1;
  }
  return 0;
}
static int it_cut(iter_t* it) {
  for (int i = it->n - 1; i >= 0; i--) {
    if (it->c[i] != -1 && it->c[i] < it->sizes[i] - 1) { it->c[i]++; it->skip = 1; return 1; }
    it->c[i] = -1;
  }
  return 0;
}

/* fillCombinationQueue, DroneClassifier.cpp:156-198.  heap = 1: push into a libstdc++ heap;
 * heap = 0: append in DFS order. */
static void fill_queue(clf_t* cl, const frame_dets* fd, comb_vec* q, int heap) {
  iter_t it;
  it.n = cl->n_cams; it.skip = 0;
  for (int i = 0; i < it.n; i++) { it.c[i] = -1; it.sizes[i] = fd->n[i] + 1; }
  int order = 0;
  while (it_increment(&it)) {
    cl->st.nodes++;
    int count = 0, has_unset = 0;
    for (int i = 0; i < it.n; i++) { if (it.c[i] > 0) count++; if (it.c[i] == -1) has_unset = 1; }
    if (count < 2) continue;
    int idx[ORC_MAX_CAMS], n = 0;
    double pix[2 * ORC_MAX_CAMS];
    for (int i = 0; i < it.n; i++) {
      if (it.c[i] <= 0) continue;
      idx[n] = i; pix[2 * n] = fd->xy[i][2 * (it.c[i] - 1)]; pix[2 * n + 1] = fd->xy[i][2 * (it.c[i] - 1) + 1]; n++;
    }
    comb_t c;
    int iters = 0;
    cl->st.solves++;
    c.err = cl->mode == ORC_MATRIX ? orc_matrix_point(cl->cams, n, idx, pix, c.p)
                                   : orc_ray_point(cl->cams, n, idx, pix, c.p, &iters);
    cl->st.lm_iters += iters;
    margin(0, c.err, cl->error_);
    if (c.err > cl->error_) {
      if (!it_cut(&it)) break;
    } else if (!has_unset && count >= MIN_CAMERAS) {
      c.zeros = 0;
      memset(c.c, 0, sizeof(c.c));
      for (int i = 0; i < it.n; i++) { /* getOriginalCombination, DetectionsContainer.cpp:173-186 */
        c.c[i] = (int8_t)(fd->use_remap ? fd->remap[i][it.c[i]] : it.c[i]);
        if (c.c[i] == 0) c.zeros++;
      }
      c.order = order++;
      cl->st.leaves++;
      if (heap) pq_push(q, &c); else cv_push(q, &c);
    }
  }
}

/* isCombinationUnique, DroneClassifier.cpp:32-41 */
static int comb_unique(const comb_t* c, const comb_t* list, int n, int n_cams) {
  for (int k = 0; k < n; k++)
    for (int i = 0; i < n_cams; i++)
      if (c->c[i] == list[k].c[i] && c->c[i] != 0) return 0;
  return 1;
}

static void tie_audit(clf_t* cl, const comb_vec* q) {
  /* count (zeros, error) keys that occur more than once: there the pop order is an artefact of
   * the heap algorithm, and the CUDA engine's documented tie rule (DFS order) may differ */
  if (q->n > 1) { /* ordering margin: smallest error gap between entries of equal priority class */
    double* key = (double*)malloc(sizeof(double) * 2 * (size_t)q->n);
    for (int i = 0; i < q->n; i++) { key[2 * i] = q->v[i].zeros; key[2 * i + 1] = q->v[i].err; }
    qsort(key, (size_t)q->n, 2 * sizeof(double), cmp_zero_err);
    for (int i = 1; i < q->n; i++)
      if (key[2 * i] == key[2 * i - 2] && key[2 * i + 1] != key[2 * i - 1]) margin(3, key[2 * i + 1], key[2 * i - 1]);
    free(key);
  }
  for (int i = 0; i < q->n; i++)
    for (int j = i + 1; j < q->n; j++)
      if (q->v[i].zeros == q->v[j].zeros && q->v[i].err == q->v[j].err) { cl->st.ties++; break; }
}

int orc_enumerate_frame(const orc_camera* cams, int n_cams, int mode, const int32_t* offs, const double* dets,
                        int n_frames, int frame, int max_leaves, int8_t* out_comb, double* out_xyz,
                        double* out_err, orc_stats* stats) {
  clf_t cl; memset(&cl, 0, sizeof(cl));
  cl.cams = cams; cl.n_cams = n_cams; cl.mode = mode;
  cl.error_ = mode == ORC_MATRIX ? MAX_ERROR_MATRIX : MAX_ERROR_RAY;
  frame_dets fd; fd.use_remap = 0;
  for (int c = 0; c < n_cams; c++) {
    int32_t a = offs[c * (n_frames + 1) + frame], b = offs[c * (n_frames + 1) + frame + 1];
    fd.n[c] = b - a; fd.xy[c] = dets + 2 * (int64_t)a;
  }
  comb_vec q = {0, 0, 0};
  fill_queue(&cl, &fd, &q, 0);
  for (int i = 0; i < q.n && i < max_leaves; i++) {
    for (int c = 0; c < n_cams; c++) out_comb[(int64_t)i * n_cams + c] = q.v[i].c[c];
    out_xyz[3 * i] = q.v[i].p[0]; out_xyz[3 * i + 1] = q.v[i].p[1]; out_xyz[3 * i + 2] = q.v[i].p[2];
    out_err[i] = q.v[i].err;
  }
  int n = q.n;
  free(q.v);
  if (stats) *stats = cl.st;
  return n;
}

typedef struct { double (*p)[3]; int n, cap; } path_t;
static void path_push(path_t* p, const double v[3]) {
  if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 1024; p->p = (double(*)[3])realloc(p->p, (size_t)p->cap * 3 * sizeof(double)); }
  memcpy(p->p[p->n++], v, 3 * sizeof(double));
}
static int in_list(const int* l, int n, int v) { for (int i = 0; i < n; i++) if (l[i] == v) return 1; return 0; }
static double dist3(const double a[3], const double b[3]) {
  double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2]; /* cv::norm(Point3d) */
  return sqrt(x * x + y * y + z * z);
}

int orc_classify(const orc_camera* cams, int n_cams, int mode, int n_drones, const int32_t* offs,
                 const double* dets, int n_frames, double* out_paths, int8_t* out_assign, uint8_t* out_phase,
                 orc_stats* stats) {
  if (n_cams > ORC_MAX_CAMS || n_drones > 64) return ORC_ERR_DIM;
  clf_t cl; memset(&cl, 0, sizeof(cl));
  for (int k = 0; k < 5; k++) g_margin[k] = HUGE_VAL;
  cl.cams = cams; cl.n_cams = n_cams; cl.mode = mode; cl.n_drones = n_drones;
  cl.error_ = mode == ORC_MATRIX ? MAX_ERROR_MATRIX : MAX_ERROR_RAY; /* DroneClassifier.cpp:3-10 */
  path_t* paths = (path_t*)calloc((size_t)n_drones, sizeof(path_t));
  char* empty = (char*)calloc((size_t)n_drones * n_frames, 1); /* emptyFrames[i] contains frame */
  memset(out_assign, -1, (size_t)n_drones * n_frames * n_cams);
  memset(out_phase, 0, (size_t)n_drones * n_frames);
  comb_vec q = {0, 0, 0};
  comb_t* used = (comb_t*)malloc(sizeof(comb_t) * 64);
  comb_vec fin = {0, 0, 0};
  frame_dets* fd = (frame_dets*)malloc(sizeof(frame_dets));
  frame_dets* gd = (frame_dets*)malloc(sizeof(frame_dets));

  for (int frame = 0; frame < n_frames; frame++) { /* classifyDrones, DroneClassifier.cpp:112-144 */
    int processed[128], n_proc = 0, n_used = 0;
    fd->use_remap = 0;
    for (int c = 0; c < n_cams; c++) {
      int32_t a = offs[c * (n_frames + 1) + frame], b = offs[c * (n_frames + 1) + frame + 1];
      fd->n[c] = b - a; fd->xy[c] = dets + 2 * (int64_t)a;
      if (fd->n[c] > 127) return ORC_ERR_DIM;
    }
    for (int np = 0; np < n_drones; np++) { /* phase 1, :119-135 */
      path_t* cur = &paths[np];
      if (cur->n == 0) continue;
      const double* last = cur->p[cur->n - 1];
      if (last[0] == 0 && last[1] == 0 && last[2] == 0) continue;
      /* triangulateWithLastPos, :219-250 */
      gd->use_remap = 1;
      for (int c = 0; c < n_cams; c++) {
        gd->n[c] = 0; gd->xy[c] = gd->gated_xy[c]; gd->remap[c][0] = 0;
        for (int d = 0; d < fd->n[c]; d++) {
          double x = fd->xy[c][2 * d], y = fd->xy[c][2 * d + 1];
          const double gate = orc_dist_from_ray(&cams[c], x, y, last);
          margin(2, gate, MAX_STEP);
          if (gate < MAX_STEP) {
            gd->gated_xy[c][2 * gd->n[c]] = x; gd->gated_xy[c][2 * gd->n[c] + 1] = y;
            gd->n[c]++;
            gd->remap[c][gd->n[c]] = d + 1;
          }
        }
      }
      q.n = 0;
      fill_queue(&cl, gd, &q, 1);
      tie_audit(&cl, &q);
      int found = 0;
      comb_t best;
      while (q.n > 0) {
        comb_t c = q.v[0];
        if (comb_unique(&c, used, n_used, n_cams) && c.err < cl.error_) {
          const double step = dist3(c.p, last);
          margin(1, step, MAX_STEP);
          if (step < MAX_STEP) { best = c; found = 1; break; }
        }
        pq_pop(&q);
      }
      if (found) {
        processed[n_proc++] = np;
        used[n_used++] = best;
        path_push(cur, best.p);
        for (int c = 0; c < n_cams; c++) out_assign[((int64_t)np * n_frames + frame) * n_cams + c] = best.c[c];
        out_phase[(int64_t)np * n_frames + frame] = 1;
        cl.st.phase1++;
      }
    }
    if (n_proc == n_drones) continue; /* :137 */
    /* pickBestCombinations, :200-217 */
    q.n = 0; fin.n = 0;
    fill_queue(&cl, fd, &q, 1);
    tie_audit(&cl, &q);
    while (q.n > 0) {
      comb_t c = q.v[0];
      if (comb_unique(&c, fin.v, fin.n, n_cams) && comb_unique(&c, used, n_used, n_cams) && c.err < cl.error_) cv_push(&fin, &c);
      pq_pop(&q);
    }
    /* classifyPaths, :262-332 */
    int nf = fin.n;
    int* cp_comb = (int*)malloc(sizeof(int) * (nf + 1));
    int* cp_path = (int*)malloc(sizeof(int) * (nf + 1));
    double* cp_err = (double*)malloc(sizeof(double) * (nf + 1));
    for (int i = 0; i < nf; i++) {
      int bestPath = 0; double bestDist = -1;
      for (int j = 0; j < n_drones; j++) {
        if (in_list(processed, n_proc, j)) continue;
        int npc = paths[j].n < PATH_TAIL ? paths[j].n : PATH_TAIL;
        if (npc == 0) continue;
        double dist = 0;
        for (int t = paths[j].n - npc; t < paths[j].n; t++) dist += dist3(paths[j].p[t], fin.v[i].p);
        dist /= (double)npc;
        if (bestDist != -1) margin(4, dist, bestDist);
        if (dist < bestDist || bestDist == -1) { bestDist = dist; bestPath = j; }
      }
      cp_comb[i] = i; cp_path[i] = bestPath; cp_err[i] = bestDist;
    }
    /* std::sort(..., greater<CombinationPath>()) with operator> = error < elem.error.  libstdc++
     * uses insertion sort for <= 16 elements (stable); restated as a stable insertion sort. */
    for (int i = 1; i < nf; i++) {
      int c0 = cp_comb[i], p0 = cp_path[i]; double e0 = cp_err[i]; int j = i - 1;
      for (int k = 0; k < i; k++) if (e0 != -1 && cp_err[k] != -1) margin(4, e0, cp_err[k]);
      while (j >= 0 && e0 < cp_err[j]) { cp_comb[j + 1] = cp_comb[j]; cp_path[j + 1] = cp_path[j]; cp_err[j + 1] = cp_err[j]; j--; }
      cp_comb[j + 1] = c0; cp_path[j + 1] = p0; cp_err[j + 1] = e0;
    }
    for (int k = 0; k < nf; k++) {
      int target = -1;
      if (in_list(processed, n_proc, cp_path[k])) {
        for (int i = 0; i < n_drones; i++) if (paths[i].n == 0) { target = i; break; }
      } else target = cp_path[k];
      if (target != -1 && n_proc < 128) {
        const comb_t* c = &fin.v[cp_comb[k]];
        path_push(&paths[target], c->p);
        processed[n_proc++] = target;
        for (int cc = 0; cc < n_cams; cc++) out_assign[((int64_t)target * n_frames + frame) * n_cams + cc] = c->c[cc];
        out_phase[(int64_t)target * n_frames + frame] = 2;
        cl.st.phase2++;
      }
    }
    for (int i = 0; i < n_drones; i++)
      if (!in_list(processed, n_proc, i)) empty[(int64_t)i * n_frames + frame] = 1;
    free(cp_comb); free(cp_path); free(cp_err);
  }
  /* :147-153: re-insert (0,0,0) at the recorded empty frames */
  for (int i = 0; i < n_drones; i++) {
    int k = 0;
    for (int f = 0; f < n_frames; f++) {
      double* o = out_paths + ((int64_t)i * n_frames + f) * 3;
      if (empty[(int64_t)i * n_frames + f] || k >= paths[i].n) { o[0] = o[1] = o[2] = 0; }
      else { memcpy(o, paths[i].p[k++], 3 * sizeof(double)); }
    }
    free(paths[i].p);
  }
  free(paths); free(empty); free(q.v); free(used); free(fin.v); free(fd); free(gd);
  if (stats) *stats = cl.st;
  return ORC_OK;
}
