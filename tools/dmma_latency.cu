// DMMA m8n8k4 throughput versus warps per SM and independent accumulator chains per warp (B200 probe).
#include <cuda_runtime.h>
#include <cstdio>
template <int ILP>
__global__ void k(double *out, int iters) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(double *out, int warps_per_sm) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<148, warps_per_sm * 32>>>(out, iters / 10);
  cudaEventRecord(e0);
  k<ILP><<<148, warps_per_sm * 32>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double mmas = 148.0 * warps_per_sm * iters * ILP;
  double clk = ms * 1e-3 * 1.965e9;
  printf("warps/SM %2d ILP %d : %.2f TFLOP/s, %.1f clk per MMA per SMSP, chain step %.1f clk\n", warps_per_sm, ILP,
         mmas * 512 / (ms * 1e-3) / 1e12, clk / (iters * ILP * warps_per_sm / 4.0), clk / iters);
}
int main() {
  double *out; cudaMalloc(&out, 8 * 148 * 1024);
  for (int w : {4, 8, 16, 32}) { run<1>(out, w); run<2>(out, w); run<4>(out, w); run<8>(out, w); }
  return 0;
}
