// oracle/shim/opencv2/opencv.hpp -- the slice of OpenCV's C++ API the reference's hot path compiles against.
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// OpenCV C++ is not installed in the build image, so the reference's own sources
// (/root/reference/src/{Triangulator,MatrixTriangulator,RayTriangulator,DroneClassifier,DetectionsContainer,
// utils,main}.cpp and Camera.h) cannot be built against the real library.  This header lets them compile
// UNMODIFIED (oracle/Makefile -> oracle/_ref/): cv::Mat (CV_64F, 2-D, shared storage like cv::Mat), Mat_<double>
// with the comma initialiser, Matx, Vec, Point_, Point3_, Ptr, InputArray/OutputArray, invert, transpose, norm,
// normalize and LMSolver.  The control flow that is executed is therefore the reference's own; the third-party
// arithmetic below restates OpenCV 4.x (modules/core/src/lapack.cpp, matmul, modules/calib3d/src/levmarq.cpp)
// and forwards to the primitives of tri_oracle.c that tests/test_oracle_pin.py and
// tests/test_twin_lm_vs_opencv.py pin against the real cv2 4.13 wheel:
//     cv::invert(DECOMP_SVD) -> orc_pinv_svd       cv::solve / cv::invert(DECOMP_EIG) -> orc_eig_solve3 / orc_eig_inv_diag3
//     cv::Mat::inv() 3x3     -> orc_inv3           cv::gemm / mulTransposed / norm    -> the summation orders probed on cv2
//
// Taps: Point_ and Point3_ carry a hidden 64-bit `tag` that copies along with the value and is ignored by every
// operator.  The harness (oracle/ref_harness.cpp) tags each detection with its index and each triangulated point
// with the triangulatePoint call that produced it, and so reads the assignment decisions of
// DroneClassifier.cpp:130 and :315-321 off the returned paths without touching the reference sources.
#ifndef TRI_ORACLE_SHIM_OPENCV_HPP
#define TRI_ORACLE_SHIM_OPENCV_HPP

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" {
void orc_pinv_svd(const double* A, int m, int n, double* pinv);
void orc_eig_solve3(const double A[9], const double b[3], double x[3]);
void orc_eig_inv_diag3(const double A[9], double diag[3]);
int orc_inv3(const double S[9], double T[9]);
}

#define CV_64F 6
#define CV_64FC1 6

namespace cv {

// ---- taps (defined in ref_harness.cpp; no-ops when unset) ----
namespace tap {
extern void (*on_norm3)(uint64_t tag_a, uint64_t tag_b, double value);  // cv::norm(Point3d a - b)
}

enum DecompTypes { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3, DECOMP_QR = 4, DECOMP_NORMAL = 16 };
enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_L2SQR = 5 };

template <typename T, int N>
struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
  Vec(T a, T b) : Vec() { val[0] = a; val[1] = b; }
  Vec(T a, T b, T c) : Vec() { val[0] = a; val[1] = b; val[2] = c; }
  Vec(T a, T b, T c, T d) : Vec() { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  Vec conj() const {  // Vec<T,4>: quaternion conjugate
    static_assert(N == 4, "conj() is defined for 4-vectors");
    return Vec(val[0], -val[1], -val[2], -val[3]);
  }
};
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 4> Vec4d;
template <typename T, int N> Vec<T, N> operator-(const Vec<T, N>& a, const Vec<T, N>& b) { Vec<T, N> r; for (int i = 0; i < N; i++) r[i] = a[i] - b[i]; return r; }
template <typename T, int N> Vec<T, N> operator+(const Vec<T, N>& a, const Vec<T, N>& b) { Vec<T, N> r; for (int i = 0; i < N; i++) r[i] = a[i] + b[i]; return r; }
template <typename T, int N> Vec<T, N> operator*(const Vec<T, N>& a, double s) { Vec<T, N> r; for (int i = 0; i < N; i++) r[i] = a[i] * s; return r; }
// Vec<T,4> * Vec<T,4>: Hamilton product, element 0 = w (core/matx.hpp)
template <typename T>
Vec<T, 4> operator*(const Vec<T, 4>& v1, const Vec<T, 4>& v2) {
  return Vec<T, 4>(v1[0] * v2[0] - v1[1] * v2[1] - v1[2] * v2[2] - v1[3] * v2[3], v1[0] * v2[1] + v1[1] * v2[0] + v1[2] * v2[3] - v1[3] * v2[2],
                   v1[0] * v2[2] - v1[1] * v2[3] + v1[2] * v2[0] + v1[3] * v2[1], v1[0] * v2[3] + v1[1] * v2[2] - v1[2] * v2[1] + v1[3] * v2[0]);
}
// cv::normalize(Vec): v * (1 / norm) (core/matx.hpp)
template <typename T, int N>
Vec<T, N> normalize(const Vec<T, N>& v) {
  double s = 0;
  for (int i = 0; i < N; i++) s += (double)v[i] * v[i];
  const double nv = std::sqrt(s);
  return v * (nv ? 1. / nv : 0.);
}

template <typename T>
struct Point_ {
  T x, y;
  uint64_t tag = 0;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<double> Point2d;
template <typename T> bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> bool operator!=(const Point_<T>& a, const Point_<T>& b) { return !(a == b); }

template <typename T>
struct Point3_ {
  T x, y, z;
  uint64_t tag = 0, tag_b = 0;  // tag_b: the right operand's tag of a difference
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
  Point3_(const Vec<T, 3>& v) : x(v[0]), y(v[1]), z(v[2]) {}
  operator Vec<T, 3>() const { return Vec<T, 3>(x, y, z); }
  Point3_ cross(const Point3_& p) const { return Point3_(y * p.z - z * p.y, z * p.x - x * p.z, x * p.y - y * p.x); }
  T dot(const Point3_& p) const { return x * p.x + y * p.y + z * p.z; }
  double ddot(const Point3_& p) const { return (double)x * p.x + (double)y * p.y + (double)z * p.z; }
  Point3_& operator+=(const Point3_& b) { x += b.x; y += b.y; z += b.z; return *this; }
};
typedef Point3_<double> Point3d;
template <typename T> bool operator==(const Point3_<T>& a, const Point3_<T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> bool operator!=(const Point3_<T>& a, const Point3_<T>& b) { return !(a == b); }
template <typename T> Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) {
  Point3_<T> r(a.x - b.x, a.y - b.y, a.z - b.z);
  r.tag = a.tag; r.tag_b = b.tag;
  return r;
}
template <typename T> Point3_<T> operator*(const Point3_<T>& a, double s) { return Point3_<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> Point3_<T> operator*(double s, const Point3_<T>& a) { return a * s; }
template <typename T> Point3_<T> operator/(const Point3_<T>& a, double s) { return Point3_<T>(a.x / s, a.y / s, a.z / s); }
// cv::norm(Point3_): sqrt(x^2 + y^2 + z^2)
template <typename T>
double norm(const Point3_<T>& p) {
  const double v = std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z);
  if (tap::on_norm3) tap::on_norm3(p.tag, p.tag_b, v);
  return v;
}
template <typename T, int N>
double norm(const Vec<T, N>& v) { double s = 0; for (int i = 0; i < N; i++) s += (double)v[i] * v[i]; return std::sqrt(s); }

template <typename T, int M, int N>
struct Matx {
  T val[M * N];
  Matx() { for (int i = 0; i < M * N; i++) val[i] = T(0); }
  Matx(T v0, T v1, T v2, T v3, T v4, T v5, T v6, T v7, T v8) { const T v[9] = {v0, v1, v2, v3, v4, v5, v6, v7, v8}; static_assert(M * N == 9, "3x3"); for (int i = 0; i < 9; i++) val[i] = v[i]; }
  T& operator()(int i, int j) { return val[i * N + j]; }
  const T& operator()(int i, int j) const { return val[i * N + j]; }
};

// ---- cv::Mat: CV_64F, 2-D, reference-counted storage (copies are shallow, like cv::Mat) ----
class Mat {
 public:
  int rows = 0, cols = 0;
  std::shared_ptr<std::vector<double>> buf;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  template <typename T, int M, int N>
  explicit Mat(const Matx<T, M, N>& m) { create(M, N, CV_64F); for (int i = 0; i < M * N; i++) (*buf)[i] = (double)m.val[i]; }
  void create(int r, int c, int type) {
    if (type != CV_64F) throw std::runtime_error("shim cv::Mat: CV_64F only");
    if (buf && rows == r && cols == c) return;
    rows = r; cols = c;
    buf = std::make_shared<std::vector<double>>((size_t)r * c, 0.0);
  }
  bool empty() const { return !buf || rows * cols == 0; }
  int type() const { return CV_64F; }
  size_t total() const { return (size_t)rows * cols; }
  double* data() { return buf->data(); }
  const double* data() const { return buf->data(); }
  template <typename T> T& at(int i, int j) { static_assert(sizeof(T) == 8, "CV_64F"); return (*buf)[(size_t)i * cols + j]; }
  template <typename T> const T& at(int i, int j) const { return (*buf)[(size_t)i * cols + j]; }
  // single index: element i of a vector (row or column), else row i column 0 (core/mat.inl.hpp)
  template <typename T> T& at(int i) { return (*buf)[(rows == 1 || cols == 1) ? (size_t)i : (size_t)i * cols]; }
  template <typename T> const T& at(int i) const { return (*buf)[(rows == 1 || cols == 1) ? (size_t)i : (size_t)i * cols]; }
  template <typename T> T* ptr(int row = 0) { return buf->data() + (size_t)row * cols; }
  template <typename T> const T* ptr(int row = 0) const { return buf->data() + (size_t)row * cols; }
  Mat clone() const { Mat m(rows, cols, CV_64F); if (buf) *m.buf = *buf; return m; }
  void copyTo(Mat& dst) const { dst = clone(); }
  Mat reshape(int cn, int new_rows) const {  // same storage, new shape
    (void)cn;
    Mat m = *this;
    const int total = rows * cols;
    m.rows = new_rows; m.cols = new_rows ? total / new_rows : 0;
    return m;
  }
  Mat t() const { Mat m(cols, rows, CV_64F); for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) m.at<double>(j, i) = at<double>(i, j); return m; }
  Mat diag() const { Mat d(std::min(rows, cols), 1, CV_64F); for (int i = 0; i < d.rows; i++) d.at<double>(i) = at<double>(i, i); return d; }
  double dot(const Mat& m) const { double s = 0; for (size_t i = 0; i < total(); i++) s += (*buf)[i] * (*m.buf)[i]; return s; }
  Mat inv(int method = DECOMP_LU) const;
  // Mat::forEach: the functor gets (element, position); serial here (the reference's use sums into a captured
  // variable from OpenCV's parallel_for_, a data race there -- SURVEY.md 5 -- whose serial order this is)
  template <typename T, typename F>
  void forEach(const F& f) {
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < cols; j++) { const int pos[2] = {i, j}; f(at<T>(i, j), pos); }
  }
};

template <typename T> class Mat_;
template <typename T>
class MatCommaInitializer_ {
 public:
  Mat_<T>* m;
  size_t idx;
  MatCommaInitializer_(Mat_<T>* m_) : m(m_), idx(0) {}
  template <typename V> MatCommaInitializer_& operator,(V v) { (*m->buf)[idx++] = (T)v; return *this; }
  operator Mat_<T>() const { return *m; }
};
template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, CV_64F) { static_assert(sizeof(T) == 8, "CV_64F"); }
  Mat_(const Mat& m) : Mat(m) {}
};
template <typename T, typename V>
MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v) {
  // OpenCV: the temporary Mat_ is shared by the initialiser; storage is reference counted, so a copy aliases it
  Mat_<T>* held = new Mat_<T>(m);  // leaked on purpose: a handful of 3x3 / 4x1 constants per camera
  MatCommaInitializer_<T> ci(held);
  return (ci, v);
}

// ---- matrix expressions: evaluated eagerly ----
// A * B: cv::gemm, sums over k in index order for these small CV_64F operands (the order tri_oracle.c restates and
// tests/test_oracle_pin.py checks against cv2)
inline Mat operator*(const Mat& a, const Mat& b) {
  if (a.cols != b.rows) throw std::runtime_error("shim cv::Mat: gemm size mismatch");
  Mat r(a.rows, b.cols, CV_64F);
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < b.cols; j++) {
      double s = 0;
      for (int k = 0; k < a.cols; k++) s += a.at<double>(i, k) * b.at<double>(k, j);
      r.at<double>(i, j) = s;
    }
  return r;
}
inline Mat operator*(double s, const Mat& a) { Mat r = a.clone(); for (double& v : *r.buf) v *= s; return r; }
inline Mat operator*(const Mat& a, double s) { return s * a; }
inline Mat operator-(const Mat& a) { Mat r = a.clone(); for (double& v : *r.buf) v = -v; return r; }
inline Mat operator-(const Mat& a, const Mat& b) { Mat r = a.clone(); for (size_t i = 0; i < r.total(); i++) (*r.buf)[i] -= (*b.buf)[i]; return r; }
inline Mat operator+(const Mat& a, const Mat& b) { Mat r = a.clone(); for (size_t i = 0; i < r.total(); i++) (*r.buf)[i] += (*b.buf)[i]; return r; }
inline void transpose(const Mat& src, Mat& dst) { dst = src.t(); }
inline void subtract(const Mat& a, const Mat& b, Mat& dst) { dst = a - b; }

// cv::invert: DECOMP_SVD = pseudo-inverse through the one-sided Jacobi SVD (JacobiSVDImpl_ + SVBkSb);
// DECOMP_LU on 3x3 = the closed-form adjugate branch; DECOMP_EIG = symmetric Jacobi eigen-decomposition.
inline double invert(const Mat& src, Mat& dst, int method = DECOMP_LU) {
  const int m = src.rows, n = src.cols;
  if (method == DECOMP_SVD) {
    if (n > 4 || m < n) throw std::runtime_error("shim cv::invert(SVD): m x n with n <= 4 <= m only");
    Mat r(n, m, CV_64F);
    orc_pinv_svd(src.data(), m, n, r.data());
    dst = r;
    return 1;
  }
  if (m != 3 || n != 3) throw std::runtime_error("shim cv::invert: 3x3 only for LU / EIG");
  Mat r(3, 3, CV_64F);
  if (method == DECOMP_EIG) {
    // SVBkSb against the identity: column j of the inverse = solve(A, e_j)
    for (int j = 0; j < 3; j++) {
      double e[3] = {0, 0, 0}, x[3];
      e[j] = 1;
      orc_eig_solve3(src.data(), e, x);
      for (int i = 0; i < 3; i++) r.at<double>(i, j) = x[i];
    }
    double dg[3];  // the diagonal exactly as the LM loop reads it
    orc_eig_inv_diag3(src.data(), dg);
    for (int i = 0; i < 3; i++) r.at<double>(i, i) = dg[i];
    dst = r;
    return 1;
  }
  const int ok = orc_inv3(src.data(), r.data());
  dst = r;
  return ok;
}
inline Mat Mat::inv(int method) const { Mat r; invert(*this, r, method); return r; }
inline bool solve(const Mat& A, const Mat& b, Mat& x, int method) {
  if (method != DECOMP_EIG || A.rows != 3 || A.cols != 3 || b.total() != 3) throw std::runtime_error("shim cv::solve: 3x3 DECOMP_EIG only");
  Mat r(3, 1, CV_64F);
  orc_eig_solve3(A.data(), b.data(), r.data());
  x = r;
  return true;
}

// ---- InputArray / OutputArray (what LMSolver::Callback::compute receives) ----
class _InputArray {
 public:
  Mat* m = nullptr;
  std::vector<double>* v = nullptr;
  _InputArray() {}
  _InputArray(const Mat& m_) : m(const_cast<Mat*>(&m_)) {}
  _InputArray(const std::vector<double>& v_) : v(const_cast<std::vector<double>*>(&v_)) {}
  Mat getMat() const {
    if (m) return *m;
    if (v) { Mat r((int)v->size(), 1, CV_64F); *r.buf = *v; return r; }  // a copy; LMSolver::run writes the result back
    return Mat();
  }
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() {}
  _OutputArray(Mat& m_) : _InputArray(m_) {}
  _OutputArray(std::vector<double>& v_) : _InputArray(v_) {}
  bool needed() const { return m != nullptr || v != nullptr; }
  void create(int rows, int cols, int type) const { if (m) m->create(rows, cols, type); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline const _OutputArray& noArray() { static _OutputArray none; return none; }

template <typename T>
struct Ptr : public std::shared_ptr<T> {
  Ptr() {}
  Ptr(T* p) : std::shared_ptr<T>(p) {}
  Ptr(const std::shared_ptr<T>& p) : std::shared_ptr<T>(p) {}
  template <typename Y> Ptr(const Ptr<Y>& p) : std::shared_ptr<T>(p) {}
};
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return Ptr<T>(new T(std::forward<A>(a)...)); }

// ---- cv::LMSolver (calib3d/src/levmarq.cpp, LMSolverImpl::run), restated on the Mat operations above ----
class LMSolver {
 public:
  class Callback {
   public:
    virtual ~Callback() {}
    virtual bool compute(InputArray param, OutputArray err, OutputArray J) const = 0;
  };
  virtual ~LMSolver() {}
  virtual int run(InputOutputArray param) const = 0;
  static Ptr<LMSolver> create(const Ptr<LMSolver::Callback>& cb, int maxIters);
  static Ptr<LMSolver> create(const Ptr<LMSolver::Callback>& cb, int maxIters, double eps);
};

namespace shim {
// mulTransposed(J, A, true): A = J^T J, each entry summed over the rows in order
inline void mulTransposed_ata(const Mat& J, Mat& A) {
  const int n = J.rows, p = J.cols;
  Mat r(p, p, CV_64F);
  for (int a = 0; a < p; a++)
    for (int b = a; b < p; b++) {
      double s = 0;
      for (int i = 0; i < n; i++) s += J.at<double>(i, a) * J.at<double>(i, b);
      r.at<double>(a, b) = r.at<double>(b, a) = s;
    }
  A = r;
}
// gemm(J, r, 1, noArray(), 0, v, GEMM_1_T): four partial sums over k, the tail into the first (GEMMSingleMul's
// unrolled loop, probed against cv2 4.13 in tests/test_oracle_pin.py)
inline void gemm_atb(const Mat& J, const Mat& r, Mat& v) {
  const int n = J.rows, p = J.cols;
  Mat out(p, 1, CV_64F);
  for (int a = 0; a < p; a++) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
      s0 += J.at<double>(i, a) * r.at<double>(i);
      s1 += J.at<double>(i + 1, a) * r.at<double>(i + 1);
      s2 += J.at<double>(i + 2, a) * r.at<double>(i + 2);
      s3 += J.at<double>(i + 3, a) * r.at<double>(i + 3);
    }
    for (; i < n; i++) s0 += J.at<double>(i, a) * r.at<double>(i);
    out.at<double>(a) = ((s0 + s1) + s2) + s3;
  }
  v = out;
}
// norm(r, NORM_L2SQR) as probed on cv2 4.13 for n <= 15: groups of four in order, the n % 4 tail fused
inline double norm_l2sqr(const Mat& r) {
  const int n = (int)r.total(), k = n / 4 * 4;
  double s = 0;
  int i = 0;
  for (; i < k; i++) s += r.at<double>(i) * r.at<double>(i);
  for (; i < n; i++) s = std::fma(r.at<double>(i), r.at<double>(i), s);
  return s;
}
inline double norm_inf(const Mat& r) { double s = 0; for (size_t i = 0; i < r.total(); i++) s = std::max(s, std::fabs(r.data()[i])); return s; }

class LMSolverImpl : public LMSolver {
 public:
  Ptr<LMSolver::Callback> cb;
  int maxIters;
  double epsx, epsf;
  LMSolverImpl(const Ptr<LMSolver::Callback>& cb_, int maxIters_, double eps) : cb(cb_), maxIters(maxIters_), epsx(eps), epsf(eps) {}
  int run(InputOutputArray param0) const override {
    Mat x = param0.getMat().clone(), xd, r, rd, J, A, Ap, v, temp_d, d;
    const int lx = (int)x.total();
    if (lx != 3) throw std::runtime_error("shim cv::LMSolver: 3 parameters only");
    x = x.reshape(1, lx);
    if (!cb->compute(x, r, J)) return -1;
    double S = norm_l2sqr(r);
    mulTransposed_ata(J, A);
    gemm_atb(J, r, v);
    Mat D = A.diag().clone();
    const double Rlo = 0.25, Rhi = 0.75;
    double lambda = 1, lc = 0.75;
    int iter = 0;
    for (;;) {
      A.copyTo(Ap);
      for (int i = 0; i < lx; i++) Ap.at<double>(i, i) += lambda * D.at<double>(i);
      solve(Ap, v, d, DECOMP_EIG);
      subtract(x, d, xd);
      if (!cb->compute(xd, rd, noArray())) return -1;
      const double Sd = norm_l2sqr(rd);
      {  // gemm(A, d, -1, v, 2, temp_d)
        Mat t(lx, 1, CV_64F);
        for (int i = 0; i < lx; i++) {
          double s = 0;
          for (int k = 0; k < lx; k++) s += A.at<double>(i, k) * d.at<double>(k);
          t.at<double>(i) = -1 * s + 2 * v.at<double>(i);
        }
        temp_d = t;
      }
      const double dS = d.dot(temp_d);
      const double R = (S - Sd) / (std::fabs(dS) > DBL_EPSILON ? dS : 1);
      if (R > Rhi) {
        lambda *= 0.5;
        if (lambda < lc) lambda = 0;
      } else if (R < Rlo) {
        const double t = d.dot(v);
        double nu = (Sd - S) / (std::fabs(t) > DBL_EPSILON ? t : 1) + 2;
        nu = std::min(std::max(nu, 2.), 10.);
        if (lambda == 0) {
          invert(A, Ap, DECOMP_EIG);
          double maxval = DBL_EPSILON;
          for (int i = 0; i < lx; i++) maxval = std::max(maxval, std::abs(Ap.at<double>(i, i)));
          lambda = lc = 1. / maxval;
          nu *= 0.5;
        }
        lambda *= nu;
      }
      if (Sd < S) {
        S = Sd;
        std::swap(x, xd);
        if (!cb->compute(x, r, J)) return -1;
        mulTransposed_ata(J, A);
        gemm_atb(J, r, v);
      }
      iter++;
      const bool proceed = iter < maxIters && norm_inf(d) >= epsx && norm_inf(r) >= epsf;
      if (!proceed) break;
    }
    last_iters() = iter;
    if (param0.v) for (int i = 0; i < lx; i++) (*param0.v)[i] = x.at<double>(i);
    else if (param0.m) for (int i = 0; i < lx; i++) param0.m->data()[i] = x.at<double>(i);
    if (iter == maxIters) iter = -iter;
    return iter;
  }
  static int& last_iters() { static thread_local int n = 0; return n; }  // read by the harness (the reference drops run()'s result)
};
}  // namespace shim

inline Ptr<LMSolver> LMSolver::create(const Ptr<LMSolver::Callback>& cb, int maxIters) { return Ptr<LMSolver>(new shim::LMSolverImpl(cb, maxIters, (double)FLT_EPSILON)); }
inline Ptr<LMSolver> LMSolver::create(const Ptr<LMSolver::Callback>& cb, int maxIters, double eps) { return Ptr<LMSolver>(new shim::LMSolverImpl(cb, maxIters, eps)); }

// ---- calib3d entry points Camera.h names in members the hot path never calls ----
inline void Rodrigues(const Mat&, Mat&) { throw std::runtime_error("shim: cv::Rodrigues is not on the hot path"); }
template <typename A, typename B>
inline void projectPoints(const A&, const Mat&, const Mat&, const Mat&, const Mat&, B&) { throw std::runtime_error("shim: cv::projectPoints is not on the hot path"); }
inline void undistort(const Mat&, Mat&, const Mat&, const Mat&) { throw std::runtime_error("shim: cv::undistort is not on the hot path"); }

}  // namespace cv
#endif
