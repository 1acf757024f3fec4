// FP64 throughput probe for B200: plain DFMA versus mma.sync DMMA (m8n8k4, m16n8k8, m16n8k16).
// Prints TFLOP/s for each; used once to fix the FP64 roofline denominator (MEASURED_PEAKS.json has none).
#include <cuda_runtime.h>
#include <cstdio>

__global__ void k_dfma(double *out, int iters) {
  double a[8];
  const double b = 1.000000001, c = 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double *out, int iters) {
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1688(double *out, int iters) {
  double c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  double a0 = 1.0 + threadIdx.x * 1e-6, a1 = 0.5, a2 = 0.25, a3 = 0.125, b0 = 1.0 - threadIdx.x * 1e-6, b1 = 0.75;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
double time_kernel(K k, double *out, int grid, int block, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<<<grid, block>>>(out, iters / 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<<<grid, block>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int grid = p.multiProcessorCount * 4, block = 512, iters = 20000;
  double *out;
  cudaMalloc(&out, sizeof(double) * grid * block);
  const double threads = (double)grid * block, warps = threads / 32;
  double t = time_kernel(k_dfma, out, grid, block, iters);
  printf("{\"probe\":\"dfma\",\"tflops\":%.2f}\n", threads * iters * 8 * 2 / t / 1e12);
  t = time_kernel(k_dmma884, out, grid, block, iters);
  printf("{\"probe\":\"dmma_m8n8k4\",\"tflops\":%.2f}\n", warps * iters * 4 * (8 * 8 * 4 * 2.0) / t / 1e12);
  t = time_kernel(k_dmma1688, out, grid, block, iters);
  printf("{\"probe\":\"dmma_m16n8k8\",\"tflops\":%.2f}\n", warps * iters * 4 * (16 * 8 * 8 * 2.0) / t / 1e12);
  printf("{\"sms\":%d,\"clock_khz\":%d,\"err\":\"%s\"}\n", p.multiProcessorCount, p.clockRate, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
