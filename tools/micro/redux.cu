// Latency / throughput of warp-collective instructions on one SM: REDUX.OR (__reduce_or_sync), CREDUX.MIN (__reduce_min_sync),
// VOTE (__ballot_sync), SHFL, with 1 .. 16 warps of one CTA issuing dependent chains.   nvcc -arch=sm_100a -O3 -o redux redux.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(unsigned* out, long long* cyc, int n) {
  unsigned v = threadIdx.x * 2654435761u + 1, acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    unsigned r;
    if (OP == 0) r = __reduce_or_sync(0xffffffffu, v);
    else if (OP == 1) r = __reduce_min_sync(0xffffffffu, v);
    else if (OP == 2) r = __ballot_sync(0xffffffffu, v & 1);
    else if (OP == 3) r = __shfl_xor_sync(0xffffffffu, v, 1);
    else r = __reduce_add_sync(0xffffffffu, v);
    v = v * 3 + r + i;  // dependent
    acc ^= r;
  }
  const long long t1 = clock64();
  if (threadIdx.x % 32 == 0) cyc[threadIdx.x / 32] = t1 - t0;
  out[threadIdx.x] = acc + v;
}
int main() {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 256);
  const char* names[5] = {"REDUX.OR", "CREDUX.MIN", "VOTE", "SHFL", "REDUX.SUM"};
  const int n = 1000;
  for (int op = 0; op < 5; op++)
    for (int warps = 1; warps <= 16; warps *= 2) {
      for (int rep = 0; rep < 2; rep++) {
        if (op == 0) k<0><<<1, 32 * warps>>>(out, cyc, n);
        if (op == 1) k<1><<<1, 32 * warps>>>(out, cyc, n);
        if (op == 2) k<2><<<1, 32 * warps>>>(out, cyc, n);
        if (op == 3) k<3><<<1, 32 * warps>>>(out, cyc, n);
        if (op == 4) k<4><<<1, 32 * warps>>>(out, cyc, n);
      }
      long long h[16];
      cudaMemcpy(h, cyc, 8 * warps, cudaMemcpyDeviceToHost);
      printf("%-10s warps %2d: %.1f cycles per dependent op (warp 0)\n", names[op], warps, (double)h[0] / n);
    }
  return 0;
}
