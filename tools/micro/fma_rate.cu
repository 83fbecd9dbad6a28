// Micro-benchmark: issue rate of FFMA, FFMA2 (packed f32x2) and DFMA on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(float* out, int iters, float seed) {
  float a[8]; float2 b[8]; double d[8];
  for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x; b[i] = make_float2(a[i], a[i] + 1); d[i] = a[i]; }
  const float m = 1.0001f; const float2 m2 = make_float2(m, m); const double md = 1.0001;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (KIND == 0) a[i] = fmaf(a[i], m, 0.5f);
        if (KIND == 1) b[i] = __ffma2_rn(b[i], m2, m2);
        if (KIND == 2) d[i] = fma(d[i], md, 0.5);
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += a[i] + b[i].x + b[i].y + (float)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const char* names[3] = {"FFMA", "FFMA2", "DFMA"};
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int kind = 0; kind < 3; kind++) {
      const int iters = 4000, grid = p.multiProcessorCount * 2, block = warps * 32 / 2;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (kind == 0) k<0><<<grid, block>>>(out, iters, 1.f);
        if (kind == 1) k<1><<<grid, block>>>(out, iters, 1.f);
        if (kind == 2) k<2><<<grid, block>>>(out, iters, 1.f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double inst = (double)grid * block / 32 * iters * 64;  // warp-instructions
      double per_clk_sm = inst / (ms * 1e-3) / (p.clockRate * 1e3) / p.multiProcessorCount;
      printf("%-6s warps/SM=%2d  %.3f ms  %.2f warp-inst/clk/SM (clock %d MHz)\n", names[kind], warps, ms, per_clk_sm, p.clockRate / 1000);
    }
  }
  return 0;
}
