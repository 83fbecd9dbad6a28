// tri_matrix.cu -- kernel 1: batched MatrixTriangulator::triangulatePoints for sm_100a.
//
// Reference semantics (src/MatrixTriangulator.cpp): per frame, every camera whose pixel is not the
// (-1,-1) sentinel (:86-87) contributes the two rows
//     P[0,0:3] - x P[2,0:3] | x P[2,3] - P[0,3]        (:16-27, :42-45)
//     P[1,0:3] - y P[2,0:3] | y P[2,3] - P[1,3]        (:29-40, :46-49)
// of a 2n x 3 system A X = b solved by X = pinv_SVD(A) b (:53-54); the reported error is
// sqrt(|A X - b|^2 / 2n) (:55-59).  One thread owns two frames (tri_batch.cuh), accumulates the
// 3x3 normal matrix A^T A and A^T b in registers -- 26 FMA per view with the camera constants read
// from the constant bank -- and solves by the adjugate.  The full-rank pseudo-inverse solution IS
// the normal-equation solution; cond(A) <= ~5 on real rigs so nothing is lost.
// HBM traffic per frame: 8 B x n_cams in, 12 B out.
#include "tri_batch.cuh"

namespace tri {

template <typename T_, bool CENTRED>
struct DltPolicy {
  using T = T_;
  using Rig = DltRig<T>;
  struct Acc {
    T M[6] = {0, 0, 0, 0, 0, 0};
    T v[3] = {0, 0, 0};
  };
  static __device__ __forceinline__ void add(const Rig& rig, int c, T x, T y, Acc& a) {
    const T(&P)[12] = rig.P[c];
    if constexpr (CENTRED) { x -= rig.pix0[c][0]; y -= rig.pix0[c][1]; }
    T a0 = P[0] - x * P[8], a1 = P[1] - x * P[9], a2 = P[2] - x * P[10], b = x * P[11] - P[3];
    a.M[0] += a0 * a0; a.M[1] += a0 * a1; a.M[2] += a0 * a2; a.M[3] += a1 * a1; a.M[4] += a1 * a2; a.M[5] += a2 * a2;
    a.v[0] += a0 * b; a.v[1] += a1 * b; a.v[2] += a2 * b;
    a0 = P[4] - y * P[8]; a1 = P[5] - y * P[9]; a2 = P[6] - y * P[10]; b = y * P[11] - P[7];
    a.M[0] += a0 * a0; a.M[1] += a0 * a1; a.M[2] += a0 * a2; a.M[3] += a1 * a1; a.M[4] += a1 * a2; a.M[5] += a2 * a2;
    a.v[0] += a0 * b; a.v[1] += a1 * b; a.v[2] += a2 * b;
  }
  static __device__ __forceinline__ void solve(const Rig&, const Acc& a, int, T (&X)[3], int, int&) {
    solve_sym3<T>(a.M, a.v, X);
  }
  // |A X - b|^2 contribution of one camera (MatrixTriangulator.cpp:55-58)
  static __device__ __forceinline__ T residual(const Rig& rig, int c, T x, T y, const T (&X)[3]) {
    const T(&P)[12] = rig.P[c];
    if constexpr (CENTRED) { x -= rig.pix0[c][0]; y -= rig.pix0[c][1]; }
    T e0 = (P[0] - x * P[8]) * X[0] + (P[1] - x * P[9]) * X[1] + (P[2] - x * P[10]) * X[2] - (x * P[11] - P[3]);
    T e1 = (P[4] - y * P[8]) * X[0] + (P[5] - y * P[9]) * X[1] + (P[6] - y * P[10]) * X[2] - (y * P[11] - P[7]);
    return e0 * e0 + e1 * e1;
  }
  static __device__ __forceinline__ double error(T sum, int n) { return sqrt((double)sum / (double)(2 * n)); }
  static __device__ __forceinline__ void to_world(const Rig& rig, T (&X)[3]) {
    if constexpr (CENTRED) { X[0] += rig.origin[0]; X[1] += rig.origin[1]; X[2] += rig.origin[2]; }
  }
};

cudaError_t launch_dlt(const LaunchCtx& ctx, bool f32, int pixfmt, const DltRig<double>& rig64,
                       const DltRig<float>& rig32, const void* d_xy, int n_use, int64_t n_frames,
                       int64_t cam_stride, const BatchOut& out) {
  if (n_frames <= 0) return cudaSuccess;
  using P32 = DltPolicy<float, true>;
  using P64 = DltPolicy<double, false>;
  if (f32) {
    switch (pixfmt) {
      case PIX_F32: return launch_batch_policy<P32, PIX_F32>(ctx, rig32, d_xy, n_use, n_frames, cam_stride, out, 0);
      case PIX_F64: return launch_batch_policy<P32, PIX_F64>(ctx, rig32, d_xy, n_use, n_frames, cam_stride, out, 0);
      default: return launch_batch_policy<P32, PIX_U16>(ctx, rig32, d_xy, n_use, n_frames, cam_stride, out, 0);
    }
  }
  switch (pixfmt) {
    case PIX_F32: return launch_batch_policy<P64, PIX_F32>(ctx, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);
    case PIX_F64: return launch_batch_policy<P64, PIX_F64>(ctx, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);
    default: return launch_batch_policy<P64, PIX_U16>(ctx, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);
  }
}

}  // namespace tri
