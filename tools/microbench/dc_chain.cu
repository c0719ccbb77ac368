// Developer microbenchmark: the DC-block recurrence (y = b0 x + d1; d1 = b1 x + a1 y) of one warp in stages: registers only,
// with shared-memory loads, with the loads batched ahead.  Clocks per step.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/microbench/_bin/dc_chain tools/microbench/dc_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "rx_phases.cuh"

__device__ __forceinline__ float Step(float x, float &d1) {
  const float b0 = 0.927176191943378969f, b1 = -0.927176191943378969f, a1 = 0.854352383886757938f;
  const float y = b0 * x + d1;
  d1 = b1 * x + a1 * y;
  return y;
}

template <int kMode>
__global__ void k(long long *clk, float *sink, float rfg, int n) {
  __shared__ float sm[32 * 66 + 256];
  for (int i = threadIdx.x; i < 32 * 66 + 256; i += 32) sm[i] = 0.001f * (float)((i * 37) % 101) - 0.05f;
  __syncwarp();
  const float *x = sm + 65 * threadIdx.x;      /* lane L: 192 samples from 65 L (overlapping ranges, as the warm-up's) */
  float d1 = 0.0f, acc = 0.0f;
  long long t0 = clock64();
  if (kMode == 0) {                  /* registers only */
    float v = rfg;
#pragma unroll 4
    for (int i = 0; i < n; ++i) { Step(v * rfg, d1); v += 1.0f; }
  } else if (kMode == 1) {           /* one load per step, as written */
#pragma unroll 4
    for (int i = 0; i < n; ++i) Step(x[i] * rfg, d1);
  } else if (kMode == 2) {           /* batches of 4 fetched one batch ahead */
    float cur[4], nxt[4];
    for (int j = 0; j < 4; ++j) cur[j] = x[j];
    for (int i = 0; i < n; i += 4) {
      if (i + 4 < n) {
#pragma unroll
        for (int j = 0; j < 4; ++j) nxt[j] = x[i + 4 + j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) Step(cur[j] * rfg, d1);
#pragma unroll
      for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
    }
  } else if (kMode == 3) {           /* batches of 8, unconditional fetch (reads past the end are harmless) */
    float cur[8], nxt[8];
    for (int j = 0; j < 8; ++j) cur[j] = x[j];
#pragma unroll 1
    for (int i = 0; i < n; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) nxt[j] = x[i + 8 + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) Step(cur[j] * rfg, d1);
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
    }
  } else if (kMode == 4) {           /* the input products formed a batch ahead: only the chain's three operations stay in the step */
    float cur[8], nxt[8];
    for (int j = 0; j < 8; ++j) cur[j] = 0.927176191943378969f * (x[j] * rfg);
#pragma unroll 1
    for (int i = 0; i < n; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) nxt[j] = 0.927176191943378969f * (x[i + 8 + j] * rfg);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = cur[j] + d1;
        d1 = 0.854352383886757938f * y - cur[j];
        acc += y;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
    }
  }
  else if (kMode == 5) {           /* the product's DcSegment<false> */
    const t41rx::DcCoef kc = t41rx::DcCoefs();
    const t41rx::DcPost p{rfg, 3.0f, -1.01f, false};
    float lx = 0, ly = 0;
    t41rx::DcSegment<false>(const_cast<float *>(x), n, kc, p, false, d1, lx, ly);
    acc = lx + ly;
  } else if (kMode == 6) {           /* the product's DcSegment<true> (65 steps per lane in the product; here n) */
    const t41rx::DcCoef kc = t41rx::DcCoefs();
    const t41rx::DcPost p{rfg, 3.0f, -1.01f, false};
    float lx = 0, ly = 0;
    t41rx::DcSegment<true>(const_cast<float *>(sm + 66 * threadIdx.x), 64, kc, p, false, d1, lx, ly);
    acc = lx + ly;
  }
  long long t1 = clock64();
  sink[threadIdx.x] = d1 + acc;
  if (threadIdx.x == 0) *clk = t1 - t0;
}

int main() {
  long long *clk, h;
  float *sink;
  cudaMalloc(&clk, 8);
  cudaMalloc(&sink, 1024);
  const char *names[] = {"registers only", "one load per step", "batches of 4 (the product's form)", "batches of 8, unconditional fetch",
                         "batches of 8, products ahead", "DcSegment<false> of rx_phases.cuh", "DcSegment<true>, 64 steps"};
  const int n = 192;
  for (int m = 0; m < 7; ++m) {
    for (int rep = 0; rep < 3; ++rep) {
      if (m == 0) k<0><<<1, 32>>>(clk, sink, 1.5f, n);
      if (m == 1) k<1><<<1, 32>>>(clk, sink, 1.5f, n);
      if (m == 2) k<2><<<1, 32>>>(clk, sink, 1.5f, n);
      if (m == 3) k<3><<<1, 32>>>(clk, sink, 1.5f, n);
      if (m == 4) k<4><<<1, 32>>>(clk, sink, 1.5f, n);
      if (m == 5) k<5><<<1, 32>>>(clk, sink, 1.5f, n);
      if (m == 6) k<6><<<1, 32>>>(clk, sink, 1.5f, n);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %.2f clk per step (%s)\n", names[m], (double)h / (m == 6 ? 64 : n), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
