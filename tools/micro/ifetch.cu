// Instruction-fetch cost for a lone warp: a loop whose body is N dependent IMADs (straight line) or N/8 blocks of 7 IMADs + one
// taken branch, body sizes around the L0 (6 KB = 384 instr) and L1.5 (32 KB = 2048 instr) instruction caches; 1 and 4 warps / SM
// sub-partition.   nvcc -arch=sm_100a -O3 -o ifetch ifetch.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void straight(unsigned* out, long long* cyc, int iters, unsigned a, unsigned b) {
  unsigned x = threadIdx.x;
  long long t0 = 0;
  for (int it = 0; it < iters + 2; it++) {
    if (it == 2) t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) x = x * a + b;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}
// blocks of 8: 6 IMADs, a compare and a branch that is always taken at run time (the compiler cannot know)
template <int N>
__global__ void branchy(unsigned* out, long long* cyc, int iters, unsigned a, unsigned b, unsigned never) {
  unsigned x = threadIdx.x;
  long long t0 = 0;
  for (int it = 0; it < iters + 2; it++) {
    if (it == 2) t0 = clock64();
#pragma unroll
    for (int i = 0; i < N / 8; i++) {
      x = x * a + b; x = x * a + b; x = x * a + b; x = x * a + b; x = x * a + b; x = x * a + b;
      if (x == never) { asm volatile("st.global.u32 [%0], %1;" ::"l"(out + 1024 + i), "r"(x) : "memory"); x ^= 5; }  // never runs: a forward branch over it
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}
template <int N>
void run(unsigned* out, long long* cyc) {
  for (int warps : {1, 4, 16}) {
    long long h;
    const int iters = 200;
    straight<N><<<1, 32 * warps>>>(out, cyc, iters, 3, 7);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double s = (double)h / iters / N;
    branchy<N><<<1, 32 * warps>>>(out, cyc, iters, 3, 7, 0xdeadbeefu);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("body %5d instr (%5.1f KB), %2d warps: straight %.2f cycles/instr   branchy %.2f cycles/instr\n", N, N * 16 / 1024.0, warps, s, (double)h / iters / N);
  }
}
int main() {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  run<128>(out, cyc); run<256>(out, cyc); run<384>(out, cyc); run<512>(out, cyc); run<768>(out, cyc); run<1024>(out, cyc);
  run<2048>(out, cyc); run<4096>(out, cyc); run<8192>(out, cyc);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
