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
    T a0 = fma_(-x, P[8], P[0]), a1 = fma_(-x, P[9], P[1]), a2 = fma_(-x, P[10], P[2]), b = fma_(x, P[11], -P[3]);
    a.M[0] = fma_(a0, a0, a.M[0]); a.M[1] = fma_(a0, a1, a.M[1]); a.M[2] = fma_(a0, a2, a.M[2]);
    a.M[3] = fma_(a1, a1, a.M[3]); a.M[4] = fma_(a1, a2, a.M[4]); a.M[5] = fma_(a2, a2, a.M[5]);
    a.v[0] = fma_(a0, b, a.v[0]); a.v[1] = fma_(a1, b, a.v[1]); a.v[2] = fma_(a2, b, a.v[2]);
    a0 = fma_(-y, P[8], P[4]); a1 = fma_(-y, P[9], P[5]); a2 = fma_(-y, P[10], P[6]); b = fma_(y, P[11], -P[7]);
    a.M[0] = fma_(a0, a0, a.M[0]); a.M[1] = fma_(a0, a1, a.M[1]); a.M[2] = fma_(a0, a2, a.M[2]);
    a.M[3] = fma_(a1, a1, a.M[3]); a.M[4] = fma_(a1, a2, a.M[4]); a.M[5] = fma_(a2, a2, a.M[5]);
    a.v[0] = fma_(a0, b, a.v[0]); a.v[1] = fma_(a1, b, a.v[1]); a.v[2] = fma_(a2, b, a.v[2]);
  }
  static __device__ __forceinline__ void solve(const Rig&, const Acc& a, int, T (&X)[3], int, int&) {
    solve_sym3<T>(a.M, a.v, X);
  }
  // |A X - b|^2 contribution of one camera (MatrixTriangulator.cpp:55-58)
  static __device__ __forceinline__ T residual(const Rig& rig, int c, T x, T y, const T (&X)[3]) {
    const T(&P)[12] = rig.P[c];
    if constexpr (CENTRED) { x -= rig.pix0[c][0]; y -= rig.pix0[c][1]; }
    const T e0 = fma_(fma_(-x, P[8], P[0]), X[0], fma_(fma_(-x, P[9], P[1]), X[1], fma_(fma_(-x, P[10], P[2]), X[2], -fma_(x, P[11], -P[3]))));
    const T e1 = fma_(fma_(-y, P[8], P[4]), X[0], fma_(fma_(-y, P[9], P[5]), X[1], fma_(fma_(-y, P[10], P[6]), X[2], -fma_(y, P[11], -P[7]))));
    return fma_(e0, e0, mul_(e1, e1));
  }
  static __device__ __forceinline__ double error(T sum, int n) { return sqrt((double)sum / (double)(2 * n)); }
  static __device__ __forceinline__ void to_world(const Rig& rig, T (&X)[3]) {
    if constexpr (CENTRED) { X[0] += rig.origin[0]; X[1] += rig.origin[1]; X[2] += rig.origin[2]; }
  }
};

// ---- FP32 main path: packed-pair SIMD (FFMA2 / FMUL2 / FADD2, new on sm_100) ----
// One thread owns frames (2g, 2g+1) as the two halves of a float2; every multiply-add of the
// accumulation is one FFMA2, so the kernel needs half the issue slots of the scalar form (which is
// issue-bound: ncu r1a, 80 % issue-active at 59 % DRAM).  Branch-free: an absent view is weighted by
// w = 0 instead of skipped (w a is exact for w in {0,1}, so the points are bit-identical to the
// scalar DltPolicy<float> path).  Rig constants are pre-duplicated/negated float2 so the operands come
// from the uniform datapath.
struct DltRigX2 {
  float2 A[TRI_MAX_CAMS][8];   // P0 P1 P2 -P3 | P4 P5 P6 -P7   (addends of the x row, the y row)
  float2 N[TRI_MAX_CAMS][4];   // -P8 -P9 -P10 P11              (multiplicands)
  float2 npix0[TRI_MAX_CAMS][2];
  float origin[3];
};

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

template <int NC, int PIX>
__global__ void __launch_bounds__(BATCH_THREADS, 3)
dlt_f32x2_kernel(const __grid_constant__ DltRigX2 rig, const char* __restrict__ xy, int64_t row_bytes, int64_t n_pairs,
                 int n_use, BatchOut out, unsigned long long* first_bad, int64_t frame_base) {
  __shared__ __align__(16) float tile[BATCH_THREADS * 6];
  const int64_t block_pair0 = (int64_t)blockIdx.x * BATCH_THREADS;
  const int64_t pair = block_pair0 + threadIdx.x;
  const int nc = NC > 0 ? NC : n_use;
  float2 X[3] = {{0, 0}, {0, 0}, {0, 0}};
  if (pair < n_pairs) {
    float2 M[6] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}}, v[3] = {{0, 0}, {0, 0}, {0, 0}};
    uint32_t mask0 = 0, mask1 = 0;
    auto view = [&](int c, const View2<float>& q) {
      const float2 w = make_float2(q.v0 ? 1.0f : 0.0f, q.v1 ? 1.0f : 0.0f);
      mask0 |= (q.v0 ? 1u : 0u) << c;
      mask1 |= (q.v1 ? 1u : 0u) << c;
      const float2 x = __fadd2_rn(make_float2(q.x0, q.x1), rig.npix0[c][0]);
      const float2 y = __fadd2_rn(make_float2(q.y0, q.y1), rig.npix0[c][1]);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const float2 u = r == 0 ? x : y;
        const float2 a0 = fma2(u, rig.N[c][0], rig.A[c][4 * r]), a1 = fma2(u, rig.N[c][1], rig.A[c][4 * r + 1]),
                     a2 = fma2(u, rig.N[c][2], rig.A[c][4 * r + 2]), b = fma2(u, rig.N[c][3], rig.A[c][4 * r + 3]);
        const float2 w0 = mul2(a0, w), w1 = mul2(a1, w), w2 = mul2(a2, w);
        M[0] = fma2(w0, a0, M[0]); M[1] = fma2(w0, a1, M[1]); M[2] = fma2(w0, a2, M[2]);
        M[3] = fma2(w1, a1, M[3]); M[4] = fma2(w1, a2, M[4]); M[5] = fma2(w2, a2, M[5]);
        v[0] = fma2(w0, b, v[0]); v[1] = fma2(w1, b, v[1]); v[2] = fma2(w2, b, v[2]);
      }
    };
    if constexpr (NC > 0) {
      View2<float> q[NC];
#pragma unroll
      for (int c = 0; c < NC; c++) q[c] = fetch2<float, PIX>(xy + c * row_bytes, pair);
#pragma unroll
      for (int c = 0; c < NC; c++) view(c, q[c]);
    } else {
#pragma unroll 2
      for (int c = 0; c < nc; c++) view(c, fetch2<float, PIX>(xy + c * row_bytes, pair));
    }
    // adjugate solve, both frames at once (same operation order as solve_sym3<float>)
    const float2 c00 = fma2(M[3], M[5], neg2(mul2(M[4], M[4]))), c01 = fma2(M[2], M[4], neg2(mul2(M[1], M[5]))),
                 c02 = fma2(M[1], M[4], neg2(mul2(M[2], M[3]))), c11 = fma2(M[0], M[5], neg2(mul2(M[2], M[2]))),
                 c12 = fma2(M[1], M[2], neg2(mul2(M[0], M[4]))), c22 = fma2(M[0], M[3], neg2(mul2(M[1], M[1])));
    const float2 det = fma2(M[0], c00, fma2(M[1], c01, mul2(M[2], c02)));
    const float2 inv = make_float2(1.0f / det.x, 1.0f / det.y);
    X[0] = mul2(fma2(c00, v[0], fma2(c01, v[1], mul2(c02, v[2]))), inv);
    X[1] = mul2(fma2(c01, v[0], fma2(c11, v[1], mul2(c12, v[2]))), inv);
    X[2] = mul2(fma2(c02, v[0], fma2(c12, v[1], mul2(c22, v[2]))), inv);
    const bool ok0 = __popc(mask0) >= 2, ok1 = __popc(mask1) >= 2;
    X[0].x = ok0 ? X[0].x + rig.origin[0] : 0.f; X[1].x = ok0 ? X[1].x + rig.origin[1] : 0.f; X[2].x = ok0 ? X[2].x + rig.origin[2] : 0.f;
    X[0].y = ok1 ? X[0].y + rig.origin[0] : 0.f; X[1].y = ok1 ? X[1].y + rig.origin[1] : 0.f; X[2].y = ok1 ? X[2].y + rig.origin[2] : 0.f;
    if (!ok0) atomicMin(first_bad, (unsigned long long)(frame_base + 2 * pair));
    if (!ok1) atomicMin(first_bad, (unsigned long long)(frame_base + 2 * pair + 1));
    if (out.xyz_f64) {
      double* o = out.xyz_f64 + 6 * pair;
      o[0] = X[0].x; o[1] = X[1].x; o[2] = X[2].x; o[3] = X[0].y; o[4] = X[1].y; o[5] = X[2].y;
    }
    if (out.mask) reinterpret_cast<uint2*>(out.mask)[pair] = make_uint2(mask0, mask1);
  }
  if (out.xyz_f32) {
    float2* t2 = reinterpret_cast<float2*>(tile) + 3 * threadIdx.x;
    t2[0] = make_float2(X[0].x, X[1].x);
    t2[1] = make_float2(X[2].x, X[0].y);
    t2[2] = make_float2(X[1].y, X[2].y);
    __syncthreads();
    const int64_t remaining = n_pairs - block_pair0;
    float* dst = out.xyz_f32 + 6 * block_pair0;
    if (remaining >= BATCH_THREADS && ((uintptr_t)dst & 15) == 0) {
      float4* d4 = reinterpret_cast<float4*>(dst);
      const float4* s4 = reinterpret_cast<const float4*>(tile);
#pragma unroll
      for (int i = threadIdx.x; i < BATCH_THREADS * 6 / 4; i += BATCH_THREADS) __stcs(d4 + i, s4[i]);
    } else {
      const int cnt = (int)(remaining < BATCH_THREADS ? remaining : BATCH_THREADS) * 6;
      for (int i = threadIdx.x; i < cnt; i += BATCH_THREADS) dst[i] = tile[i];
    }
  }
}

static DltRigX2 make_x2(const DltRig<float>& r) {
  DltRigX2 x;
  for (int c = 0; c < TRI_MAX_CAMS; c++) {
    const float* P = r.P[c];
    const float A[8] = {P[0], P[1], P[2], -P[3], P[4], P[5], P[6], -P[7]};
    const float N[4] = {-P[8], -P[9], -P[10], P[11]};
    for (int k = 0; k < 8; k++) x.A[c][k] = make_float2(A[k], A[k]);
    for (int k = 0; k < 4; k++) x.N[c][k] = make_float2(N[k], N[k]);
    for (int k = 0; k < 2; k++) x.npix0[c][k] = make_float2(-r.pix0[c][k], -r.pix0[c][k]);
  }
  for (int k = 0; k < 3; k++) x.origin[k] = r.origin[k];
  return x;
}

// xyz / mask outputs only; vector-aligned rows only; the caller falls back to the scalar policy otherwise
template <int PIX>
static cudaError_t launch_dlt_x2(const LaunchCtx& ctx, const DltRig<float>& rig32, const char* xy, int64_t row_bytes,
                                 int n_use, int64_t n_pairs, const BatchOut& out) {
  const DltRigX2 rig = make_x2(rig32);
  const unsigned grid = (unsigned)((n_pairs + BATCH_THREADS - 1) / BATCH_THREADS);
#define TRI_CASE(N)                                                                                                    \
  case N:                                                                                                              \
    dlt_f32x2_kernel<N, PIX><<<grid, BATCH_THREADS, 0, ctx.stream>>>(rig, xy, row_bytes, n_pairs, n_use, out,          \
                                                                     ctx.d_first_bad, ctx.frame_base);                  \
    break;
  switch (n_use) {
    TRI_CASE(2) TRI_CASE(3) TRI_CASE(4) TRI_CASE(5) TRI_CASE(6) TRI_CASE(7) TRI_CASE(8)
    default:
      dlt_f32x2_kernel<0, PIX><<<grid, BATCH_THREADS, 0, ctx.stream>>>(rig, xy, row_bytes, n_pairs, n_use, out, ctx.d_first_bad,
                                                                       ctx.frame_base);
  }
#undef TRI_CASE
  ++*ctx.launches;
  return cudaGetLastError();
}

template <int PIX>
static cudaError_t launch_dlt_f32(const LaunchCtx& ctx, const DltRig<float>& rig32, const void* d_xy, int n_use,
                                  int64_t n_frames, int64_t cam_stride, const BatchOut& out) {
  using P32 = DltPolicy<float, true>;
  const char* xy = static_cast<const char*>(d_xy);
  const int64_t row_bytes = cam_stride * pix_bytes(PIX), need = 2 * pix_bytes(PIX);
  const bool packed = !out.err && !out.iters && PIX != PIX_F64 && ((uintptr_t)xy % need == 0) && (row_bytes % need == 0) &&
                      (out.mask == nullptr || (uintptr_t)out.mask % 8 == 0) && n_frames >= 2;
  if (!packed) return launch_batch_policy<P32, PIX, 2, 3>(ctx, rig32, d_xy, n_use, n_frames, cam_stride, out, 0);
  const int64_t n_pairs = n_frames / 2;
  cudaError_t err = launch_dlt_x2<PIX>(ctx, rig32, xy, row_bytes, n_use, n_pairs, out);
  if (err != cudaSuccess || 2 * n_pairs == n_frames) return err;
  const int64_t done = 2 * n_pairs;  // odd tail frame: scalar kernel
  batch_single_kernel<P32, PIX><<<1, BATCH_THREADS, 0, ctx.stream>>>(rig32, xy, row_bytes, done, n_frames, n_use, out, 0,
                                                                     ctx.d_first_bad, ctx.frame_base);
  ++*ctx.launches;
  return cudaGetLastError();
}

cudaError_t launch_dlt(const LaunchCtx& ctx, bool f32, int pixfmt, const DltRig<double>& rig64,
                       const DltRig<float>& rig32, const void* d_xy, int n_use, int64_t n_frames,
                       int64_t cam_stride, const BatchOut& out) {
  if (n_frames <= 0) return cudaSuccess;
  using P64 = DltPolicy<double, false>;
  if (f32) {
    switch (pixfmt) {
      case PIX_F32: return launch_dlt_f32<PIX_F32>(ctx, rig32, d_xy, n_use, n_frames, cam_stride, out);
      case PIX_F64: return launch_dlt_f32<PIX_F64>(ctx, rig32, d_xy, n_use, n_frames, cam_stride, out);
      default: return launch_dlt_f32<PIX_U16>(ctx, rig32, d_xy, n_use, n_frames, cam_stride, out);
    }
  }
  switch (pixfmt) {
    case PIX_F32: return launch_batch_policy<P64, PIX_F32, 1, 3>(ctx, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);
    case PIX_F64: return launch_batch_policy<P64, PIX_F64, 1, 3>(ctx, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);
    default: return launch_batch_policy<P64, PIX_U16, 1, 3>(ctx, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);
  }
}

}  // namespace tri
