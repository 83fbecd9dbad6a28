// Does a DFMA with three distinct register-pair sources run slower than one with a repeated / reused source?
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(const double* in, double* out, int iters) {
  double x[6], y[6], acc[6];
  for (int i = 0; i < 6; i++) { x[i] = in[threadIdx.x + 32 * i]; y[i] = in[threadIdx.x + 32 * (i + 6)]; acc[i] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < 6; i++) {
        if (KIND == 0) acc[i] = fma(x[i], y[i], acc[i]);            // 3 distinct pairs, nothing reusable
        if (KIND == 1) acc[i] = fma(x[0], y[i], acc[i]);            // slot A reused across the 6
        if (KIND == 2) acc[i] = fma(x[i], x[i], acc[i]);            // 2 distinct pairs
        if (KIND == 3) acc[i] = fma(acc[i], x[0], y[0]);            // chain, 3 pairs but two shared
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 6; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *in, *out; cudaMalloc(&in, 4096 * 8); cudaMalloc(&out, 148 * 8 * 1024 * 8);
  double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0 + 1e-9 * (i % 977);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const char* names[4] = {"3 distinct pairs", "A reused", "2 distinct pairs", "chain"};
  for (int warps = 4; warps <= 32; warps *= 2)
    for (int kind = 0; kind < 4; kind++) {
      const int iters = 4000, grid = p.multiProcessorCount, block = warps * 32;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        switch (kind) { case 0: k<0><<<grid, block>>>(in, out, iters); break; case 1: k<1><<<grid, block>>>(in, out, iters); break;
                        case 2: k<2><<<grid, block>>>(in, out, iters); break; default: k<3><<<grid, block>>>(in, out, iters); }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double inst = (double)grid * block / 32 * iters * 48;
      printf("warps/SM=%2d %-18s %.3f ms  %.2f DFMA warp-inst/clk/SM\n", warps, names[kind], ms, inst / (ms * 1e-3) / (p.clockRate * 1e3) / p.multiProcessorCount);
    }
  return 0;
}
