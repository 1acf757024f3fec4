// Dependent-chain latencies of the fp64 building blocks used by the 8x8 factorisation (B200 probe): one warp.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(double *out, long long *clk, double seed) {
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  const int N = 256;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  t1 = clock64(); clk[0] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = 1.0 / x + 0.5;
  t1 = clock64(); clk[1] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 0.5;
  t1 = clock64(); clk[2] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  t1 = clock64(); clk[3] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 0.5; }
  t1 = clock64(); clk[4] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = sqrt(x) + 0.5;
  t1 = clock64(); clk[5] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64(); clk[6] = (t1 - t0);
  out[threadIdx.x] = x;
}
int main() {
  double *out; long long *clk; cudaMalloc(&out, 256); cudaMalloc(&clk, 64);
  k<<<1, 32>>>(out, clk, 1.5); k<<<1, 32>>>(out, clk, 1.5);
  long long h[8]; cudaMemcpy(h, clk, 56, cudaMemcpyDeviceToHost);
  const char *nm[] = {"dfma", "div+add", "rsqrt+add", "shfl64", "rcp.approx+add", "sqrt+add", "dmul"};
  for (int i = 0; i < 7; ++i) printf("%-16s %.1f clk per dependent op\n", nm[i], h[i] / 256.0);
  return 0;
}
