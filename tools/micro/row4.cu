// DFMA issue rate by operand source: register / uniform register / constant bank.
#include <cstdio>
#include <cuda_runtime.h>
struct Rig { double P[8][12]; };
template <int KIND>
__global__ void k(const __grid_constant__ Rig rig, const double* in, double* out, int iters) {
  double x[4], acc[8];
  for (int i = 0; i < 4; i++) x[i] = in[threadIdx.x + 32 * i];
  for (int i = 0; i < 8; i++) acc[i] = 0;
  double r[8];
  for (int i = 0; i < 8; i++) r[i] = in[threadIdx.x + 32 * (i + 4)];
  for (int it = 0; it < iters; it++) {
    x[it & 3] = __hiloint2double(__double2hiint(x[it & 3]), __double2loint(x[it & 3]) + it);
#pragma unroll
    for (int c = 0; c < 8; c++) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (KIND == 0) acc[2 * i] = fma(-x[i], rig.P[c][8 + i], rig.P[c][i]) + acc[2 * i] * 0;  // placeholder, see KIND 1..3
        if (KIND == 1) acc[i] = fma(x[i], rig.P[c][8 + i], acc[i]);                 // reg, const, acc(reg)
        if (KIND == 2) acc[i] = fma(x[i], r[(c + i) & 7], acc[i]);                  // reg, reg, acc(reg): 3 pairs
        if (KIND == 3) acc[i + 4 * (c & 1)] = fma(-x[i], rig.P[c][8 + i], rig.P[c][i]) ;  // the row form: reg, const, const
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *in, *out; cudaMalloc(&in, 4096 * 8); cudaMalloc(&out, 148 * 1024 * 8);
  double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0 + 1e-9 * (i % 977);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  Rig rig; for (int c = 0; c < 8; c++) for (int i = 0; i < 12; i++) rig.P[c][i] = 1.0 + 0.001 * (c * 12 + i);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const char* names[4] = {"", "fma(reg, const, acc)", "fma(reg, reg, acc)", "row: fma(reg, const, const)"};
  for (int kind = 1; kind <= 3; kind++) {
    const int iters = 4000, grid = p.multiProcessorCount, block = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (kind == 1) k<1><<<grid, block>>>(rig, in, out, iters);
      if (kind == 2) k<2><<<grid, block>>>(rig, in, out, iters);
      if (kind == 3) k<3><<<grid, block>>>(rig, in, out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)grid * block / 32 * iters * 32;
    printf("%-30s %.3f ms  %.2f DFMA warp-inst/clk/SM\n", names[kind], ms, inst / (ms * 1e-3) / (p.clockRate * 1e3) / p.multiProcessorCount);
  }
  return 0;
}
