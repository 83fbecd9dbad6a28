// DFMA issue rate by operand source on sm_100a: register-only forms vs a uniform-register (constant bank)
// operand, with and without the LDCU that feeds it.  16 independent accumulator chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
struct K { double c[512]; };
template <int KIND>
__global__ void __launch_bounds__(512) k(const __grid_constant__ K kc, const double* in, double* out, int iters) {
  double x[4], acc[16];
  for (int i = 0; i < 4; i++) x[i] = in[threadIdx.x + 32 * i];
  for (int i = 0; i < 16; i++) acc[i] = in[threadIdx.x + 32 * (4 + i)];
  for (int it = 0; it < iters; it++) {
    const int base = (it * 32) & 511;  // uniform, loop-dependent: the constants cannot be hoisted
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (KIND == 0) acc[i] = fma(x[u], x[(u + 1) & 3], acc[i]);        // 3 distinct register pairs, two of them shared by 16 in a row
        if (KIND == 1) acc[i] = fma(x[u], acc[(i + 1) & 15], acc[i]);     // 3 distinct register pairs, one shared
        if (KIND == 2) acc[i] = fma(x[u], kc.c[i & 7], acc[i]);           // UR operand, 8 loop-invariant constants (no LDCU in the loop)
        if (KIND == 3) acc[i] = fma(x[u], kc.c[base + ((u * 16 + i) >> 1)], acc[i]);  // UR operand, one LDCU.64 per two DFMAs
        if (KIND == 4) acc[i] = fma(x[u], kc.c[base + ((u * 16 + i) >> 2)], acc[i]);  // UR operand, one LDCU.64 per four DFMAs
        if (KIND == 5) acc[i] = fma(acc[i], acc[i], x[u]);                // 2 distinct
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *in, *out; cudaMalloc(&in, 8192 * 8); cudaMalloc(&out, 148 * 8 * 1024 * 8);
  double h[8192]; for (int i = 0; i < 8192; i++) h[i] = 1.0 + 1e-9 * (i % 977);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  K kc; for (int i = 0; i < 512; i++) kc.c[i] = 1e-3 * (i + 1);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const char* names[6] = {"3 reg pairs, 2 shared", "3 reg pairs, 1 shared", "UR operand, invariant", "UR operand, LDCU per 2", "UR operand, LDCU per 4", "2 reg pairs"};
  for (int warps = 8; warps <= 16; warps *= 2)
    for (int kind = 0; kind < 6; kind++) {
      const int iters = 4000, grid = p.multiProcessorCount, block = warps * 32;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        switch (kind) {
          case 0: k<0><<<grid, block>>>(kc, in, out, iters); break;
          case 1: k<1><<<grid, block>>>(kc, in, out, iters); break;
          case 2: k<2><<<grid, block>>>(kc, in, out, iters); break;
          case 3: k<3><<<grid, block>>>(kc, in, out, iters); break;
          case 4: k<4><<<grid, block>>>(kc, in, out, iters); break;
          default: k<5><<<grid, block>>>(kc, in, out, iters); break;
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaError_t err = cudaGetLastError();
      double inst = (double)grid * block / 32 * iters * 64;
      printf("warps/SM=%2d %-26s %.3f ms  %.2f DFMA warp-inst/clk/SM %s\n", warps, names[kind], ms,
             inst / (ms * 1e-3) / (p.clockRate * 1e3) / p.multiProcessorCount, err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
  return 0;
}
