// Developer microbenchmark: per-SM throughput of the FP64 operations and conversions the bit-exact mixer uses
// (one CTA of 1024 threads per SM, independent chains).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/microbench/_bin/fp64_rates tools/microbench/fp64_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int kMode>
__global__ void __launch_bounds__(1024) k(float *out, long long *clk, int iters, double a, float fa) {
  double d[4]; float f[4];
  for (int i = 0; i < 4; ++i) { d[i] = threadIdx.x + i; f[i] = threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (kMode == 0) d[i] = d[i] * a;                         // DMUL
        if (kMode == 1) d[i] = d[i] + a;                         // DADD
        if (kMode == 2) d[i] = fma(d[i], a, a);                  // DFMA
        if (kMode == 3) { d[i] = (double)f[i]; f[i] = f[i] + (float)(int)(d[i] > 1e300); }   // F2F.F64.F32 (+ cheap dependency)
        if (kMode == 4) { f[i] = (float)d[i]; d[i] = d[i] + (double)(int)(f[i] > 1e30f); }   // F2F.F32.F64 + DADD... see mode 1
        if (kMode == 5) f[i] = f[i] * fa;                        // FMUL (reference)
      }
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 4; ++i) s += (float)d[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float *out; long long *clk, h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 2000;
  const char *names[] = {"DMUL", "DADD", "DFMA", "cvt f32->f64 (+FADD, ISETP, I2F)", "cvt f64->f32 (+DADD, ...)", "FMUL"};
  for (int m = 0; m < 6; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      if (m == 0) k<0><<<148, 1024>>>(out, clk, iters, 1.0000001, 1.0001f);
      if (m == 1) k<1><<<148, 1024>>>(out, clk, iters, 1.0000001, 1.0001f);
      if (m == 2) k<2><<<148, 1024>>>(out, clk, iters, 1.0000001, 1.0001f);
      if (m == 3) k<3><<<148, 1024>>>(out, clk, iters, 1.0000001, 1.0001f);
      if (m == 4) k<4><<<148, 1024>>>(out, clk, iters, 1.0000001, 1.0001f);
      if (m == 5) k<5><<<148, 1024>>>(out, clk, iters, 1.0000001, 1.0001f);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double lane_ops = 1024.0 * iters * 16;
    printf("%-36s %.1f lane-operations per clock per SM\n", names[m], lane_ops / (double)h);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
