// DMMA throughput versus operand reuse: is the 16-cycle issue rate of mma.sync.m8n8k4.f64 reachable when every
// instruction reads different A / B registers (as the triangular solver does), and with only 4 warps per scheduler?
// Variants: same A,B | distinct A, same B | distinct A and B | m16n8k8 distinct.  Grid: 148 CTAs x (warps) warps.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double *out, int iters) {
  double c[12][2], a[12], b[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) { c[i][0] = c[i][1] = 0.0; a[i] = 1.0 + (threadIdx.x + i) * 1e-6; b[i] = 1.0 - (threadIdx.x + 3 * i) * 1e-6; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      if (MODE == 0) dmma(c[i][0], c[i][1], a[0], b[0]);
      if (MODE == 1) dmma(c[i][0], c[i][1], a[i], b[0]);
      if (MODE == 2) dmma(c[i][0], c[i][1], a[i], b[i]);
      if (MODE == 3) dmma(c[i][0], c[i][1], a[0], b[i]);
      if (MODE == 4) dmma(c[i & 1][0], c[i & 1][1], a[i], b[i]);   // two accumulation chains (backward sweep)
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k1688(double *out, int iters) {
  double c[6][4], a[6][4], b[6][2];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 4; ++j) { c[i][j] = 0.0; a[i][j] = 1.0 + (threadIdx.x + i + j) * 1e-6; }
    b[i][0] = 1.0 - (threadIdx.x + i) * 1e-6; b[i][1] = 0.5 + i * 1e-3;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[i][0]), "d"(a[i][1]), "d"(a[i][2]), "d"(a[i][3]), "d"(b[i][0]), "d"(b[i][1]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K>
double run(K kern, double *out, int grid, int block, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<grid, block>>>(out, iters / 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<<<grid, block>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3;
}
int main() {
  double *out;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  const int iters = 20000;
  const char *names[5] = {"same A, same B", "distinct A, same B", "distinct A, distinct B", "same A, distinct B", "distinct A,B, 2 chains"};
  for (int warps : {4, 8, 16, 32}) {
    const int block = 32 * warps;
    const double nw = 148.0 * warps;
    double t[5];
    t[0] = run(k<0>, out, 148, block, iters); t[1] = run(k<1>, out, 148, block, iters); t[2] = run(k<2>, out, 148, block, iters);
    t[3] = run(k<3>, out, 148, block, iters); t[4] = run(k<4>, out, 148, block, iters);
    for (int m = 0; m < 5; ++m)
      printf("warps/SM %2d  m8n8k4 %-26s %6.2f TFLOP/s\n", warps, names[m], nw * iters * 12 * 512.0 / t[m] / 1e12);
    const double t6 = run(k1688, out, 148, block, iters);
    printf("warps/SM %2d  m16n8k8 distinct A, B          %6.2f TFLOP/s\n", warps, nw * iters * 6 * 2048.0 / t6 / 1e12);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
