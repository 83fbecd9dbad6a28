// tri_batch.cuh -- the streaming skeleton shared by the batched triangulatePoints kernels
// (kernel 4 of the north star: structure-of-arrays detections, validity bitmask, vector loads,
// shared-memory staged output).
//
// Layout in HBM: pixels xy[cam][frame] (float2 | double2 | ushort2), camera rows `row_bytes` apart;
// (-1,-1) -- or 0xFFFF,0xFFFF for ushort2 -- marks "no detection" (DetectionsContainer.cpp:145-171,
// MatrixTriangulator.cpp:86, RayTriangulator.cpp:66).  One thread owns two consecutive frames: one
// 16-byte load per camera (8 for ushort2), all cameras' loads issued before the first use.  The
// per-frame solve is a policy class S:
//     S::T, S::Rig (a __grid_constant__ parameter: operands come from the constant bank), S::Acc,
//     S::add(rig, c, x, y, acc)            one valid view into the accumulator
//     S::solve(rig, acc, n, X, opt, iters) the per-frame solve, X in rig coordinates
//     S::residual(rig, c, x, y, X)         that view's contribution to the reported error
//     S::error(sum, n)                     the reported error from the summed contributions
//     S::to_world(rig, X)
// Results: float3 packed (12 B) through a shared-memory tile -> coalesced 16-byte streaming stores;
// optional double3 / mask / err / iters outputs for parity tests and the classifier.
#pragma once
#include "tri_common.cuh"

namespace tri {

constexpr int BATCH_THREADS = 256;

template <typename T>
struct View2 {  // the pixels of two consecutive frames on one camera
  T x0, y0, x1, y1;
  bool v0, v1;
};

template <typename T, int PIX>
__device__ __forceinline__ View2<T> fetch2(const char* row, int64_t pair) {
  View2<T> r;
  if constexpr (PIX == PIX_F32) {
    float4 q = ld_stream(reinterpret_cast<const float4*>(row) + pair);
    r.v0 = pix_valid(q.x, q.y);
    r.v1 = pix_valid(q.z, q.w);
    r.x0 = to_real<T>(q.x); r.y0 = to_real<T>(q.y); r.x1 = to_real<T>(q.z); r.y1 = to_real<T>(q.w);
  } else if constexpr (PIX == PIX_F64) {
    double2 a = ld_stream(reinterpret_cast<const double2*>(row) + 2 * pair);
    double2 b = ld_stream(reinterpret_cast<const double2*>(row) + 2 * pair + 1);
    r.v0 = pix_valid(a.x, a.y);
    r.v1 = pix_valid(b.x, b.y);
    r.x0 = (T)a.x; r.y0 = (T)a.y; r.x1 = (T)b.x; r.y1 = (T)b.y;
  } else {
    uint2 q = ld_stream(reinterpret_cast<const uint2*>(row) + pair);
    r.v0 = q.x != 0xffffffffu;
    r.v1 = q.y != 0xffffffffu;
    if constexpr (sizeof(T) == 8) {  // integers < 2^16 -> double without touching the conversion pipe
      const double magic = 4503599627370496.0;  // 2^52
      r.x0 = __hiloint2double(0x43300000, (int)(q.x & 0xffffu)) - magic;
      r.y0 = __hiloint2double(0x43300000, (int)(q.x >> 16)) - magic;
      r.x1 = __hiloint2double(0x43300000, (int)(q.y & 0xffffu)) - magic;
      r.y1 = __hiloint2double(0x43300000, (int)(q.y >> 16)) - magic;
    } else {
      r.x0 = (T)(q.x & 0xffffu); r.y0 = (T)(q.x >> 16); r.x1 = (T)(q.y & 0xffffu); r.y1 = (T)(q.y >> 16);
    }
  }
  return r;
}

template <typename T, int PIX>
__device__ __forceinline__ void fetch1(const char* row, int64_t f, T& x, T& y, bool& v) {
  if constexpr (PIX == PIX_F32) {
    float2 q = ld_stream(reinterpret_cast<const float2*>(row) + f);
    v = pix_valid(q.x, q.y); x = (T)q.x; y = (T)q.y;
  } else if constexpr (PIX == PIX_F64) {
    double2 q = ld_stream(reinterpret_cast<const double2*>(row) + f);
    v = pix_valid(q.x, q.y); x = (T)q.x; y = (T)q.y;
  } else {
    unsigned q = ld_stream(reinterpret_cast<const unsigned*>(row) + f);
    v = q != 0xffffffffu; x = (T)(q & 0xffffu); y = (T)(q >> 16);
  }
}

// ---- two frames per thread; NC > 0 unrolls the camera loop and keeps every pixel in registers ----
template <class S, int NC, int PIX>
__global__ void __launch_bounds__(BATCH_THREADS)
batch_pairs_kernel(const __grid_constant__ typename S::Rig rig, const char* __restrict__ xy, int64_t row_bytes,
                   int64_t n_pairs, int n_use, BatchOut out, int opt, unsigned long long* first_bad, int64_t frame_base) {
  using T = typename S::T;
  __shared__ __align__(16) float tile[BATCH_THREADS * 6];
  const int64_t block_pair0 = (int64_t)blockIdx.x * BATCH_THREADS;
  const int64_t pair = block_pair0 + threadIdx.x;
  const int nc = NC > 0 ? NC : n_use;
  T X0[3] = {0, 0, 0}, X1[3] = {0, 0, 0};

  if (pair < n_pairs) {
    typename S::Acc a0, a1;
    uint32_t mask0 = 0, mask1 = 0;
    T e0 = 0, e1 = 0;
    int it0 = 0, it1 = 0;
    if constexpr (NC > 0) {
      View2<T> w[NC];
#pragma unroll
      for (int c = 0; c < NC; c++) w[c] = fetch2<T, PIX>(xy + c * row_bytes, pair);
#pragma unroll
      for (int c = 0; c < NC; c++) {
        if (w[c].v0) { S::add(rig, c, w[c].x0, w[c].y0, a0); mask0 |= 1u << c; }
        if (w[c].v1) { S::add(rig, c, w[c].x1, w[c].y1, a1); mask1 |= 1u << c; }
      }
      const int n0 = __popc(mask0), n1 = __popc(mask1);
      if (n0 >= 2) S::solve(rig, a0, n0, X0, opt, it0);
      if (n1 >= 2) S::solve(rig, a1, n1, X1, opt, it1);
      if (out.err) {
#pragma unroll
        for (int c = 0; c < NC; c++) {
          if (w[c].v0) e0 += S::residual(rig, c, w[c].x0, w[c].y0, X0);
          if (w[c].v1) e1 += S::residual(rig, c, w[c].x1, w[c].y1, X1);
        }
      }
    } else {
#pragma unroll 4
      for (int c = 0; c < nc; c++) {
        View2<T> w = fetch2<T, PIX>(xy + c * row_bytes, pair);
        if (w.v0) { S::add(rig, c, w.x0, w.y0, a0); mask0 |= 1u << c; }
        if (w.v1) { S::add(rig, c, w.x1, w.y1, a1); mask1 |= 1u << c; }
      }
      const int n0 = __popc(mask0), n1 = __popc(mask1);
      if (n0 >= 2) S::solve(rig, a0, n0, X0, opt, it0);
      if (n1 >= 2) S::solve(rig, a1, n1, X1, opt, it1);
      if (out.err) {
        for (int c = 0; c < nc; c++) {
          View2<T> w = fetch2<T, PIX>(xy + c * row_bytes, pair);
          if (w.v0) e0 += S::residual(rig, c, w.x0, w.y0, X0);
          if (w.v1) e1 += S::residual(rig, c, w.x1, w.y1, X1);
        }
      }
    }
    const int n0 = __popc(mask0), n1 = __popc(mask1);
    if (n0 >= 2) S::to_world(rig, X0);
    else atomicMin(first_bad, (unsigned long long)(frame_base + 2 * pair));
    if (n1 >= 2) S::to_world(rig, X1);
    else atomicMin(first_bad, (unsigned long long)(frame_base + 2 * pair + 1));

    if (out.xyz_f64) {
      double* o = out.xyz_f64 + 6 * pair;
      o[0] = X0[0]; o[1] = X0[1]; o[2] = X0[2]; o[3] = X1[0]; o[4] = X1[1]; o[5] = X1[2];
    }
    if (out.mask) reinterpret_cast<uint2*>(out.mask)[pair] = make_uint2(mask0, mask1);
    if (out.err)
      reinterpret_cast<double2*>(out.err)[pair] =
          make_double2(n0 >= 2 ? S::error(e0, n0) : 0.0, n1 >= 2 ? S::error(e1, n1) : 0.0);
    if (out.iters) reinterpret_cast<int2*>(out.iters)[pair] = make_int2(it0, it1);
  }

  if (out.xyz_f32) {  // uniform branch: 24 B per thread into the tile, the tile out as 16 B vectors
    float2* t2 = reinterpret_cast<float2*>(tile) + 3 * threadIdx.x;
    t2[0] = make_float2((float)X0[0], (float)X0[1]);
    t2[1] = make_float2((float)X0[2], (float)X1[0]);
    t2[2] = make_float2((float)X1[1], (float)X1[2]);
    __syncthreads();
    const int64_t remaining = n_pairs - block_pair0;
    float* dst = out.xyz_f32 + 6 * block_pair0;
    if (remaining >= BATCH_THREADS) {
      float4* d4 = reinterpret_cast<float4*>(dst);
      const float4* s4 = reinterpret_cast<const float4*>(tile);
#pragma unroll
      for (int i = threadIdx.x; i < BATCH_THREADS * 6 / 4; i += BATCH_THREADS) __stcs(d4 + i, s4[i]);
    } else {
      const int n = (int)remaining * 6;
      for (int i = threadIdx.x; i < n; i += BATCH_THREADS) dst[i] = tile[i];
    }
  }
}

// ---- fallback: one frame per thread, scalar loads (unaligned rows, odd tail frame) ----
template <class S, int PIX>
__global__ void __launch_bounds__(BATCH_THREADS)
batch_single_kernel(const __grid_constant__ typename S::Rig rig, const char* __restrict__ xy, int64_t row_bytes,
                    int64_t f_begin, int64_t f_end, int n_use, BatchOut out, int opt, unsigned long long* first_bad,
                    int64_t frame_base) {
  using T = typename S::T;
  const int64_t f = f_begin + (int64_t)blockIdx.x * BATCH_THREADS + threadIdx.x;
  if (f >= f_end) return;
  typename S::Acc acc;
  T X[3] = {0, 0, 0}, e = 0;
  uint32_t mask = 0;
  int it = 0;
  for (int c = 0; c < n_use; c++) {
    T x, y; bool ok;
    fetch1<T, PIX>(xy + c * row_bytes, f, x, y, ok);
    if (ok) { S::add(rig, c, x, y, acc); mask |= 1u << c; }
  }
  const int n = __popc(mask);
  if (n >= 2) {
    S::solve(rig, acc, n, X, opt, it);
    if (out.err)
      for (int c = 0; c < n_use; c++) {
        T x, y; bool ok;
        fetch1<T, PIX>(xy + c * row_bytes, f, x, y, ok);
        if (ok) e += S::residual(rig, c, x, y, X);
      }
    S::to_world(rig, X);
  } else {
    atomicMin(first_bad, (unsigned long long)(frame_base + f));
  }
  if (out.xyz_f32) { float* o = out.xyz_f32 + 3 * f; o[0] = (float)X[0]; o[1] = (float)X[1]; o[2] = (float)X[2]; }
  if (out.xyz_f64) { double* o = out.xyz_f64 + 3 * f; o[0] = X[0]; o[1] = X[1]; o[2] = X[2]; }
  if (out.mask) out.mask[f] = mask;
  if (out.err) out.err[f] = n >= 2 ? S::error(e, n) : 0.0;
  if (out.iters) out.iters[f] = it;
}

template <class S, int PIX>
static cudaError_t launch_batch_policy(const LaunchCtx& ctx, const typename S::Rig& rig, const void* d_xy, int n_use,
                                       int64_t n_frames, int64_t cam_stride, const BatchOut& out, int opt) {
  const char* xy = static_cast<const char*>(d_xy);
  const int64_t row_bytes = cam_stride * pix_bytes(PIX);
  const int64_t need = PIX == PIX_U16 ? 8 : 16;  // every camera row must start vector-aligned
  const bool vec = ((uintptr_t)xy % need == 0) && (row_bytes % need == 0) &&
                   (out.xyz_f32 == nullptr || (uintptr_t)out.xyz_f32 % 16 == 0) &&
                   (out.mask == nullptr || (uintptr_t)out.mask % 8 == 0) &&
                   (out.err == nullptr || (uintptr_t)out.err % 16 == 0) &&
                   (out.iters == nullptr || (uintptr_t)out.iters % 8 == 0);
  const int64_t n_pairs = vec ? n_frames / 2 : 0;
  if (n_pairs > 0) {
    const unsigned grid = (unsigned)((n_pairs + BATCH_THREADS - 1) / BATCH_THREADS);
#define TRI_CASE(N)                                                                                            \
  case N:                                                                                                      \
    batch_pairs_kernel<S, N, PIX><<<grid, BATCH_THREADS, 0, ctx.stream>>>(rig, xy, row_bytes, n_pairs, n_use, \
                                                                          out, opt, ctx.d_first_bad,           \
                                                                          ctx.frame_base);                     \
    break;
    switch (n_use) {
      TRI_CASE(2) TRI_CASE(3) TRI_CASE(4) TRI_CASE(5) TRI_CASE(6) TRI_CASE(7) TRI_CASE(8)
      default:
        batch_pairs_kernel<S, 0, PIX><<<grid, BATCH_THREADS, 0, ctx.stream>>>(rig, xy, row_bytes, n_pairs, n_use, out,
                                                                              opt, ctx.d_first_bad, ctx.frame_base);
    }
#undef TRI_CASE
    ++*ctx.launches;
  }
  const int64_t done = 2 * n_pairs;
  if (done < n_frames) {
    const unsigned grid = (unsigned)((n_frames - done + BATCH_THREADS - 1) / BATCH_THREADS);
    batch_single_kernel<S, PIX><<<grid, BATCH_THREADS, 0, ctx.stream>>>(rig, xy, row_bytes, done, n_frames, n_use, out,
                                                                        opt, ctx.d_first_bad, ctx.frame_base);
    ++*ctx.launches;
  }
  return cudaGetLastError();
}

}  // namespace tri
