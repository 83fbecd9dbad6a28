// Micro-benchmark: does the register-file port limit (3 distinct 64-bit source operands) slow FFMA2 / DFMA?
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(float* out, int iters, float seed) {
  float2 acc[6], x[6], y[6]; double dacc[6], dx[6], dy[6];
  for (int i = 0; i < 6; i++) {
    x[i] = make_float2(seed + i, seed - i); y[i] = make_float2(1.0f + 1e-6f * i, 1.0f - 1e-6f * i); acc[i] = make_float2(0, 0);
    dx[i] = seed + i; dy[i] = 1.0 + 1e-9 * i; dacc[i] = 0;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < 6; i++) {
        if (KIND == 0) acc[i] = __ffma2_rn(x[i], y[i], acc[i]);                 // 3 distinct register pairs
        if (KIND == 1) acc[i] = __ffma2_rn(x[0], y[i], acc[i]);                 // one operand shared (reuse cache)
        if (KIND == 2) acc[i] = __ffma2_rn(acc[i], y[0], y[1]);                 // chain on acc, constants shared
        if (KIND == 3) dacc[i] = fma(dx[i], dy[i], dacc[i]);
        if (KIND == 4) dacc[i] = fma(dx[0], dy[i], dacc[i]);
        if (KIND == 5) dacc[i] = fma(dacc[i], dy[0], dy[1]);
      }
    }
    // keep the operands live and changing so nothing is hoisted
    x[it & 3].x += 1e-7f; dx[it & 3] += 1e-9;
  }
  float s = 0;
  for (int i = 0; i < 6; i++) s += acc[i].x + acc[i].y + (float)dacc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const char* names[6] = {"FFMA2 3 pairs", "FFMA2 shared A", "FFMA2 chain", "DFMA 3 pairs", "DFMA shared A", "DFMA chain"};
  for (int kind = 0; kind < 6; kind++) {
    const int iters = 4000, grid = p.multiProcessorCount * 2, block = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      switch (kind) {
        case 0: k<0><<<grid, block>>>(out, iters, 1.f); break; case 1: k<1><<<grid, block>>>(out, iters, 1.f); break;
        case 2: k<2><<<grid, block>>>(out, iters, 1.f); break; case 3: k<3><<<grid, block>>>(out, iters, 1.f); break;
        case 4: k<4><<<grid, block>>>(out, iters, 1.f); break; default: k<5><<<grid, block>>>(out, iters, 1.f);
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)grid * block / 32 * iters * 48;
    printf("%-16s %.3f ms  %.2f warp-inst/clk/SM\n", names[kind], ms, inst / (ms * 1e-3) / (p.clockRate * 1e3) / p.multiProcessorCount);
  }
  return 0;
}
