// Developer microbenchmark: how fast can ONE warp issue independent FP32 instructions (FMUL / FADD / FFMA, register operands)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o /tmp/fp_issue tools/microbench/fp_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int kMode>
__global__ void k(float *out, long long *clk, int iters, float a, float b) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (kMode == 0) v[i] = v[i] * a;              // FMUL x8 independent
        if (kMode == 1) v[i] = v[i] + b;              // FADD
        if (kMode == 2) v[i] = fmaf(v[i], a, b);      // FFMA
        if (kMode == 3) v[i] = (i & 1) ? v[i] * a : v[i] + b;   // mixed
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

// the cascade step itself: dependent chain + independent work, one warp
__global__ void casc(float *out, long long *clk, int iters, float b0, float b1, float b2, float a1, float a2, int chain_only) {
  float x1 = threadIdx.x, x2 = 1, y1 = 0.5f, y2 = 0.25f, pre = 0.1f, x = 0.3f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float acc = pre + a1 * y1;
      acc = acc + a2 * y2;
      y2 = y1; y1 = acc;
      if (!chain_only) {
        x2 = x1; x1 = x; x = x + 1.0f;
        pre = b0 * x;
        pre = pre + b1 * x1;
        pre = pre + b2 * x2;
      }
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = y1 + pre;
  if (threadIdx.x == 0) *clk = t1 - t0;
}

int main() {
  float *out; long long *clk, h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&clk, 8);
  const int iters = 10000;
  const char *names[] = {"FMUL", "FADD", "FFMA", "FMUL/FADD mixed"};
  for (int warps = 1; warps <= 8; warps *= 2) {
    for (int m = 0; m < 4; ++m) {
      for (int rep = 0; rep < 2; ++rep) {
        if (m == 0) k<0><<<1, 32 * warps>>>(out, clk, iters, 1.0001f, 0.5f);
        if (m == 1) k<1><<<1, 32 * warps>>>(out, clk, iters, 1.0001f, 0.5f);
        if (m == 2) k<2><<<1, 32 * warps>>>(out, clk, iters, 1.0001f, 0.5f);
        if (m == 3) k<3><<<1, 32 * warps>>>(out, clk, iters, 1.0001f, 0.5f);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
      printf("%d warp(s) %-16s: %.2f clk per instruction per warp\n", warps, names[m], (double)h / (iters * 32.0));
    }
  }
  for (int co = 1; co >= 0; --co) {
    for (int rep = 0; rep < 2; ++rep) { casc<<<1, 32>>>(out, clk, iters, 0.9f, 0.8f, 0.7f, 0.5f, -0.3f, co); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("cascade step (%s): %.2f clk per step\n", co ? "chain only: 2 FMUL + 2 FADD" : "chain + pre: 5 FMUL + 5 FADD", (double)h / (iters * 8.0));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
