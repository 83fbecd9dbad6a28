// DFMA issue rate by operand source: constant-bank operand (2 register sources) vs three register pairs.
// Shapes follow the two candidate DLT accumulations: "direct" M_ij += a_i a_j (registers only) and the
// monomial form N_ij += x C1_ij + y C2_ij + s C3_ij (one constant-bank operand per DFMA).
#include <cstdio>
#include <cuda_runtime.h>
struct K { double c[64]; };
template <int KIND>
__global__ void k(const __grid_constant__ K kc, const double* in, double* out, int iters) {
  double x[6], y[6], acc[9];
  for (int i = 0; i < 6; i++) { x[i] = in[threadIdx.x + 32 * i]; y[i] = in[threadIdx.x + 32 * (i + 6)]; }
  for (int i = 0; i < 9; i++) acc[i] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (KIND == 0) {  // monomial: 27 DFMA, each with a constant operand
#pragma unroll
        for (int i = 0; i < 9; i++) {
          acc[i] = fma(x[0], kc.c[(u * 27 + i) & 63], acc[i]);
          acc[i] = fma(x[1], kc.c[(u * 27 + 9 + i) & 63], acc[i]);
          acc[i] = fma(x[2], kc.c[(u * 27 + 18 + i) & 63], acc[i]);
        }
      }
      if (KIND == 1) {  // direct: rows from constants (8) + 18 register-only
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const double a0 = fma(x[r], kc.c[(u * 8 + 0) & 63], kc.c[(u * 8 + 1) & 63]);
          const double a1 = fma(x[r], kc.c[(u * 8 + 2) & 63], kc.c[(u * 8 + 3) & 63]);
          const double a2 = fma(x[r], kc.c[(u * 8 + 4) & 63], kc.c[(u * 8 + 5) & 63]);
          const double b = fma(x[r], kc.c[(u * 8 + 6) & 63], kc.c[(u * 8 + 7) & 63]);
          acc[0] = fma(a0, a0, acc[0]); acc[1] = fma(a0, a1, acc[1]); acc[2] = fma(a0, a2, acc[2]);
          acc[3] = fma(a1, a1, acc[3]); acc[4] = fma(a1, a2, acc[4]); acc[5] = fma(a2, a2, acc[5]);
          acc[6] = fma(a0, b, acc[6]); acc[7] = fma(a1, b, acc[7]); acc[8] = fma(a2, b, acc[8]);
        }
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 9; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *in, *out; cudaMalloc(&in, 4096 * 8); cudaMalloc(&out, 148 * 8 * 1024 * 8);
  double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0 + 1e-9 * (i % 977);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  K kc; for (int i = 0; i < 64; i++) kc.c[i] = 1e-3 * (i + 1);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const char* names[2] = {"monomial (const operand)", "direct (registers)"};
  const int per_iter[2] = {8 * 27, 8 * 26};
  for (int warps = 4; warps <= 32; warps *= 2)
    for (int kind = 0; kind < 2; kind++) {
      const int iters = 2000, grid = p.multiProcessorCount, block = warps * 32;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (kind == 0) k<0><<<grid, block>>>(kc, in, out, iters); else k<1><<<grid, block>>>(kc, in, out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double inst = (double)grid * block / 32 * iters * per_iter[kind];
      printf("warps/SM=%2d %-26s %.3f ms  %.2f DFMA warp-inst/clk/SM  (%.1f clk per 8-view frame per SM-warp)\n", warps, names[kind], ms,
             inst / (ms * 1e-3) / (p.clockRate * 1e3) / p.multiProcessorCount, ms * 1e-3 * p.clockRate * 1e3 / iters / (block / 32));
    }
  return 0;
}
