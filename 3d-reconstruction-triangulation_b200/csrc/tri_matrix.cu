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
#include "tri_pipe.cuh"


namespace tri {

// one camera's 3x4 matrix as 16-byte vector loads (LDS.128 from the shared-memory rig)
__device__ __forceinline__ void load12(const double (&src)[12], double (&P)[12]) {
  const double2* v = reinterpret_cast<const double2*>(src);
#pragma unroll
  for (int k = 0; k < 6; k++) { const double2 q = v[k]; P[2 * k] = q.x; P[2 * k + 1] = q.y; }
}
__device__ __forceinline__ void load12(const float (&src)[12], float (&P)[12]) {
  const float4* v = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int k = 0; k < 3; k++) { const float4 q = v[k]; P[4 * k] = q.x; P[4 * k + 1] = q.y; P[4 * k + 2] = q.z; P[4 * k + 3] = q.w; }
}

// One row (a, b) into the normal equations.  An absent view is masked by clearing the row (selects on
// the ALU pipe) rather than skipped or predicated: within a warp some lane almost always has the view,
// so the FMAs issue either way; a divergent branch costs BSSY/BSYNC barriers (27 % of the stall samples
// in profiles/ncu_r1c.md) and "@p fma.rn.f64" sequences measured 45 % slower than this form.
__device__ __forceinline__ void acc_row(bool valid, double a0, double a1, double a2, double b, double (&M)[6], double (&v)[3]) {
  a0 = valid ? a0 : 0.0; a1 = valid ? a1 : 0.0; a2 = valid ? a2 : 0.0;
  // (ptxas picks the order; an inline-PTX order built for operand reuse measured the same, and clearing
  // only the high word of the three values measured the same -- the FP64 pipe's register ports are the limit)
  M[0] = fma_(a0, a0, M[0]); M[1] = fma_(a0, a1, M[1]); M[2] = fma_(a0, a2, M[2]);
  M[3] = fma_(a1, a1, M[3]); M[4] = fma_(a1, a2, M[4]); M[5] = fma_(a2, a2, M[5]);
  v[0] = fma_(a0, b, v[0]); v[1] = fma_(a1, b, v[1]); v[2] = fma_(a2, b, v[2]);
}
__device__ __forceinline__ void acc_row(bool valid, float a0, float a1, float a2, float b, float (&M)[6], float (&v)[3]) {
  a0 = valid ? a0 : 0.f; a1 = valid ? a1 : 0.f; a2 = valid ? a2 : 0.f;  // exact zeros: identical to the packed FFMA2 path
  M[0] = fma_(a0, a0, M[0]); M[1] = fma_(a0, a1, M[1]); M[2] = fma_(a0, a2, M[2]);
  M[3] = fma_(a1, a1, M[3]); M[4] = fma_(a1, a2, M[4]); M[5] = fma_(a2, a2, M[5]);
  v[0] = fma_(a0, b, v[0]); v[1] = fma_(a1, b, v[1]); v[2] = fma_(a2, b, v[2]);
}

template <typename T_, bool CENTRED>
struct DltPolicy {
  using T = T_;
  using Rig = DltRig<T>;
  // more than 8 cameras stay on the generic batch_kernel: the chunked pipeline measured slower for the DLT
  // (32 cameras x 20 M frames: FP64 2.40 ms at one frame per thread, 2.03 ms at two, vs 1.84 ms; FP32 1.29 vs 0.99 ms)
  static constexpr bool CHUNKED = false;
  static constexpr int CHUNK_FPT = 2;
  struct Acc {
    T M[6] = {0, 0, 0, 0, 0, 0};
    T v[3] = {0, 0, 0};
  };
  // Branch-free: an absent view is masked inside acc_row (predicated FMAs in FP64, exact-zero rows in FP32)
  // instead of skipped -- the divergence barriers (BSSY/BSYNC, 27 % of the stall samples in
  // profiles/ncu_r1c.md) disappear.
  static __device__ __forceinline__ void add(const Rig& rig, int c, T x, T y, bool valid, Acc& a) {
    T P[12];
    load12(rig.P[c], P);
    if constexpr (CENTRED) { x -= rig.pix0[c][0]; y -= rig.pix0[c][1]; }
    acc_row(valid, fma_(-x, P[8], P[0]), fma_(-x, P[9], P[1]), fma_(-x, P[10], P[2]), fma_(x, P[11], -P[3]), a.M, a.v);
    acc_row(valid, fma_(-y, P[8], P[4]), fma_(-y, P[9], P[5]), fma_(-y, P[10], P[6]), fma_(y, P[11], -P[7]), a.M, a.v);
  }
  static __device__ __forceinline__ void add_chunk(const Rig& rig, int c, T x, T y, bool valid, Acc& a) { add(rig, c, x, y, valid, a); }
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
// accumulation is one FFMA2, so the solve needs half the issue slots of the scalar form (which was
// issue-bound: profiles/r1a, 80 % issue-active).  Branch-free: an absent view is weighted by w = 0
// instead of skipped (w a is exact for w in {0,1}, so the points are bit-identical to the scalar
// DltPolicy<float> path, which serves the tails and the optional outputs).  Rig constants are
// pre-duplicated / pre-negated float2 so the operands come from the uniform datapath.
struct __align__(16) DltRigX2 {
  float2 A[TRI_MAX_CAMS][8];   // P0 P1 P2 -P3 | P4 P5 P6 -P7   (addends of the x row, the y row)
  float2 N[TRI_MAX_CAMS][4];   // -P8 -P9 -P10 P11              (multiplicands)
  float2 npix0[TRI_MAX_CAMS][2];
  float origin[3];
};


template <bool SEL>
struct DltX2Tile {
  static constexpr int FPT = 2;
  using Rig = DltRigX2;
  template <int NC, int PIX>
  static __device__ __forceinline__ void run(const Rig& rig, const typename RawPix<PIX, 2>::type (&raw)[NC], int,
                                             float (&X)[2][3], uint32_t (&mask)[2]) {
    float2 M[6] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}}, v[3] = {{0, 0}, {0, 0}, {0, 0}};
    uint32_t mask0 = 0, mask1 = 0;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const Views<float, PIX, 2> q = decode<float, PIX, 2>(raw[c]);
      const float2 w = make_float2(q.v[0] ? 1.0f : 0.0f, q.v[1] ? 1.0f : 0.0f);
      mask0 |= (q.v[0] ? 1u : 0u) << c;
      mask1 |= (q.v[1] ? 1u : 0u) << c;
      float2 K[14];  // this camera's constants: 7 LDS.128
      {
        const float4* kv = reinterpret_cast<const float4*>(rig.A[c]);
#pragma unroll
        for (int k = 0; k < 4; k++) { const float4 t = kv[k]; K[2 * k] = make_float2(t.x, t.y); K[2 * k + 1] = make_float2(t.z, t.w); }
        const float4* nv = reinterpret_cast<const float4*>(rig.N[c]);
#pragma unroll
        for (int k = 0; k < 2; k++) { const float4 t = nv[k]; K[8 + 2 * k] = make_float2(t.x, t.y); K[9 + 2 * k] = make_float2(t.z, t.w); }
        const float4 t = *reinterpret_cast<const float4*>(rig.npix0[c]);
        K[12] = make_float2(t.x, t.y); K[13] = make_float2(t.z, t.w);
      }
      const float2 x = __fadd2_rn(make_float2(q.x[0], q.x[1]), K[12]);
      const float2 y = __fadd2_rn(make_float2(q.y[0], q.y[1]), K[13]);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const float2 u = r == 0 ? x : y;
        const float2 a0 = fma2(u, K[8], K[4 * r]), a1 = fma2(u, K[9], K[4 * r + 1]),
                     a2 = fma2(u, K[10], K[4 * r + 2]), b = fma2(u, K[11], K[4 * r + 3]);
        float2 w0, w1, w2;
        if constexpr (SEL) {  // zero an absent view's row on the ALU pipe instead of multiplying on the FMA pipe
          w0 = make_float2(q.v[0] ? a0.x : 0.f, q.v[1] ? a0.y : 0.f);
          w1 = make_float2(q.v[0] ? a1.x : 0.f, q.v[1] ? a1.y : 0.f);
          w2 = make_float2(q.v[0] ? a2.x : 0.f, q.v[1] ? a2.y : 0.f);
        } else {
          w0 = mul2(a0, w); w1 = mul2(a1, w); w2 = mul2(a2, w);
        }
        M[0] = fma2(w0, a0, M[0]); M[1] = fma2(w0, a1, M[1]); M[2] = fma2(w0, a2, M[2]);
        M[3] = fma2(w1, a1, M[3]); M[4] = fma2(w1, a2, M[4]); M[5] = fma2(w2, a2, M[5]);
        v[0] = fma2(w0, b, v[0]); v[1] = fma2(w1, b, v[1]); v[2] = fma2(w2, b, v[2]);
      }
    }
    float2 X0, X1, X2;
    solve_sym3_x2(M, v, X0, X1, X2);
    const bool ok0 = __popc(mask0) >= 2, ok1 = __popc(mask1) >= 2;
    X[0][0] = ok0 ? X0.x + rig.origin[0] : 0.f; X[0][1] = ok0 ? X1.x + rig.origin[1] : 0.f; X[0][2] = ok0 ? X2.x + rig.origin[2] : 0.f;
    X[1][0] = ok1 ? X0.y + rig.origin[0] : 0.f; X[1][1] = ok1 ? X1.y + rig.origin[1] : 0.f; X[1][2] = ok1 ? X2.y + rig.origin[2] : 0.f;
    mask[0] = mask0; mask[1] = mask1;
  }
};

// Memory-roofline probe (TRI_DEBUG_STREAM): the same pipeline with a near-empty solve -- x = sum of the
// valid pixel x, y = sum of y, z = number of views -- to measure what the streaming skeleton alone sustains.
struct StreamProbeTile {
  static constexpr int FPT = 2;
  using Rig = DltRigX2;
  template <int NC, int PIX>
  static __device__ __forceinline__ void run(const Rig&, const typename RawPix<PIX, 2>::type (&raw)[NC], int, float (&X)[2][3],
                                             uint32_t (&mask)[2]) {
    mask[0] = mask[1] = 0;
#pragma unroll
    for (int j = 0; j < 2; j++) X[j][0] = X[j][1] = X[j][2] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const Views<float, PIX, 2> q = decode<float, PIX, 2>(raw[c]);
#pragma unroll
      for (int j = 0; j < 2; j++)
        if (q.v[j]) { X[j][0] += q.x[j]; X[j][1] += q.y[j]; X[j][2] += 1.f; mask[j] |= 1u << c; }
    }
  }
};

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

cudaError_t launch_dlt(const LaunchCtx& ctx, bool f32, int pixfmt, const DltRig<double>& rig64,
                       const DltRig<float>& rig32, const void* d_xy, int n_use, int64_t n_frames,
                       int64_t cam_stride, const BatchOut& out) {
  if (n_frames <= 0) return cudaSuccess;
  using P32 = DltPolicy<float, true>;
  using P64 = DltPolicy<double, false>;
  using T64 = PolicyTile<P64, 1>;
  if (f32) {
    const DltRigX2 x2 = make_x2(rig32);
    if (ctx.debug_stream && pixfmt == PIX_F32)
      return launch_streamed<StreamProbeTile, P32, PIX_F32, 2, 3, 2, 2>(ctx, x2, rig32, d_xy, n_use, n_frames, cam_stride, out, 0);
    if (pixfmt == PIX_F32) {
      switch (ctx.variant) {  // TRI_VARIANT: earlier generations kept for A/B measurements (profiles/r1_variants*.log)
        case 1: return launch_streamed<DltX2Tile<false>, P32, PIX_F32, 2, 3, 2, 2>(ctx, x2, rig32, d_xy, n_use, n_frames, cam_stride, out, 0);  // gen 2: TMA + mbarrier ring
        case 2: return launch_batch_policy<P32, PIX_F32, 2, 3>(ctx, rig32, d_xy, n_use, n_frames, cam_stride, out, 0);                          // gen 1: scalar, one tile per CTA
        default: return launch_streamed<DltX2Tile<true>, P32, PIX_F32, 2, 2, 3, 0>(ctx, x2, rig32, d_xy, n_use, n_frames, cam_stride, out, 0);  // gen 3: cp.async per-thread pipeline
      }
    }
    if (pixfmt == PIX_F64) return launch_streamed<DltX2Tile<true>, P32, PIX_F64, 2, 2, 3, 0>(ctx, x2, rig32, d_xy, n_use, n_frames, cam_stride, out, 0);
    return launch_streamed<DltX2Tile<true>, P32, PIX_U16, 2, 2, 3, 0>(ctx, x2, rig32, d_xy, n_use, n_frames, cam_stride, out, 0);
  }
  if (pixfmt == PIX_F32) {
    switch (ctx.variant) {
      case 1: return launch_streamed<T64, P64, PIX_F32, 1, 4, 3, 2>(ctx, rig64, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);  // gen 2
      case 2: return launch_batch_policy<P64, PIX_F32, 1, 3>(ctx, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);                // gen 1
      case 3: return launch_streamed<T64, P64, PIX_F32, 1, 3, 4, 0>(ctx, rig64, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);  // gen 3, one frame per thread
      case 4: return launch_streamed<PolicyTile<P64, 2>, P64, PIX_F32, 2, 3, 2, 0>(ctx, rig64, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);  // 3-stage ring: 1.955 ms
      // two frames per thread, 2-stage ring, 2 CTAs per SM: 1.934 ms (3 CTAs at 80 registers spill: 2.196 ms)
      default: return launch_streamed<PolicyTile<P64, 2>, P64, PIX_F32, 2, 2, 2, 0>(ctx, rig64, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);
    }
  }
  if (pixfmt == PIX_F64) return launch_streamed<T64, P64, PIX_F64, 1, 4, 3, 2>(ctx, rig64, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);
  return launch_streamed<PolicyTile<P64, 2>, P64, PIX_U16, 2, 3, 2, 0>(ctx, rig64, rig64, d_xy, n_use, n_frames, cam_stride, out, 0);
}

}  // namespace tri
