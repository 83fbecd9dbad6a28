// How fast does the 9-FMA normal-equation update (M += a a^T, v += a b) run, and does source order matter?
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int FPT>
__global__ void k(const double* in, double* out, int iters) {
  double M[FPT][6], v[FPT][3], a[FPT][4];
  for (int j = 0; j < FPT; j++) {
    for (int i = 0; i < 6; i++) M[j][i] = 0;
    for (int i = 0; i < 3; i++) v[j][i] = 0;
    for (int i = 0; i < 4; i++) a[j][i] = in[threadIdx.x + 32 * (4 * j + i)];
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int j = 0; j < FPT; j++) {
        // cheap, FP64-free refresh of the row so nothing is loop invariant
        a[j][u & 3] = __hiloint2double(__double2hiint(a[j][u & 3]), __double2loint(a[j][u & 3]) + it);
        const double a0 = a[j][0], a1 = a[j][1], a2 = a[j][2], b = a[j][3];
        if (KIND == 0) {
          M[j][0] = fma(a0, a0, M[j][0]); M[j][1] = fma(a0, a1, M[j][1]); M[j][2] = fma(a0, a2, M[j][2]);
          M[j][3] = fma(a1, a1, M[j][3]); M[j][4] = fma(a1, a2, M[j][4]); M[j][5] = fma(a2, a2, M[j][5]);
          v[j][0] = fma(a0, b, v[j][0]); v[j][1] = fma(a1, b, v[j][1]); v[j][2] = fma(a2, b, v[j][2]);
        } else {
          M[j][0] = fma(a0, a0, M[j][0]); M[j][1] = fma(a0, a1, M[j][1]); M[j][3] = fma(a1, a1, M[j][3]);
          M[j][4] = fma(a1, a2, M[j][4]); M[j][5] = fma(a2, a2, M[j][5]); v[j][2] = fma(a2, b, v[j][2]);
          v[j][1] = fma(a1, b, v[j][1]); v[j][0] = fma(a0, b, v[j][0]); M[j][2] = fma(a0, a2, M[j][2]);
        }
      }
    }
  }
  double s = 0;
  for (int j = 0; j < FPT; j++) { for (int i = 0; i < 6; i++) s += M[j][i]; for (int i = 0; i < 3; i++) s += v[j][i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND, int FPT> void run(const double* in, double* out, const cudaDeviceProp& p, const char* name) {
  for (int warps = 8; warps <= 16; warps *= 2) {
    const int iters = 2000, grid = p.multiProcessorCount, block = warps * 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k<KIND, FPT><<<grid, block>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)grid * block / 32 * iters * 8 * FPT * 9;
    printf("%-28s warps/SM=%2d %.3f ms  %.2f DFMA warp-inst/clk/SM\n", name, warps, ms, inst / (ms * 1e-3) / (p.clockRate * 1e3) / p.multiProcessorCount);
  }
}
int main() {
  double *in, *out; cudaMalloc(&in, 4096 * 8); cudaMalloc(&out, 148 * 1024 * 8);
  double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0 + 1e-9 * (i % 977);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  run<0, 1>(in, out, p, "natural order, 1 frame");
  run<1, 1>(in, out, p, "reuse-chain order, 1 frame");
  run<0, 2>(in, out, p, "natural order, 2 frames");
  run<1, 2>(in, out, p, "reuse-chain order, 2 frames");
  return 0;
}
