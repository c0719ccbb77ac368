// Developer microbenchmark: the ZoomFFT cascade loop of PhZoomIirPipe (one warp, shared-memory hand-over between the
// stages) in variants, clocks per step.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/microbench/_bin/casc_loop tools/microbench/casc_loop.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kSlot = 7208, kRawLen = 2076, kBlock = 2048;

template <int kSkew, int kAhead, int kGroup, bool kPredStore, bool kNoLoad, bool kNoStore, int kDelay = 0>
__global__ void pipe(float *out, long long *clk, const float *coef) {
  extern __shared__ float smem[];
  for (int i = threadIdx.x; i < 4 * kSlot; i += 32) smem[i] = (float)(i % 97) * 0.01f;
  __syncwarp();
  const int lane = threadIdx.x, g = lane >> 3, chn = (lane >> 2) & 1, sg = lane & 3;
  const bool active = coef[40] == 0.0f;   // runtime true
  float *x = smem + g * kSlot + chn * kRawLen + 27;
  float b0 = coef[5 * sg], b1 = coef[5 * sg + 1], b2 = coef[5 * sg + 2], a1 = coef[5 * sg + 3], a2 = coef[5 * sg + 4];
  float x1 = 0.1f, x2 = 0.2f, y1 = 0.3f, y2 = 0.4f, yold = 0.0f;
  float *xm = x - kSkew * sg;
  const int lo = kSkew * sg - 27;
  float xn0 = xm[max(0, lo)], xn1 = xm[max(1, lo)], xn2 = xm[max(2, lo)], xn3 = xm[max(3, lo)];
  float pre = b0 * xn0;
  pre = pre + b1 * x1;
  pre = pre + b2 * x2;
  long long t0 = clock64();
  int k = 40;
#pragma unroll 1
  for (; k < kBlock;) {
#pragma unroll
    for (int u_ = 0; u_ < kGroup; ++u_, ++k) {
      const float xin = xn0;
      float acc = pre + a1 * y1;
      acc = acc + a2 * y2;
      x2 = x1; x1 = xin; y2 = y1; y1 = acc;
      if (!kNoStore) {
        if (kDelay == 0) { if (!kPredStore || active) xm[k] = acc; }
        if (kDelay == 1) { if (!kPredStore || active) xm[k - 1] = y2; }     /* the previous step's output (y2 after the shift) */
        if (kDelay == 2) { if (!kPredStore || active) xm[k - 2] = yold; }
      }
      yold = y2;
      pre = b0 * xn1;
      pre = pre + b1 * x1;
      pre = pre + b2 * x2;
      xn0 = xn1; xn1 = xn2; xn2 = xn3;
      if (!kNoLoad) xn3 = xm[k + kAhead]; else xn3 = xn3 + 1.0f;
    }
    __syncwarp();
  }
  long long t1 = clock64();
  out[threadIdx.x] = y1 + pre + xn3;
  if (threadIdx.x == 0) *clk = t1 - t0;
}

template <typename K>
void run(const char *name, K kern, float *out, long long *clk, const float *coef) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kSlot * 4);
  long long h;
  for (int rep = 0; rep < 3; ++rep) { kern<<<1, 32, 4 * kSlot * 4>>>(out, clk, coef); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-58s %.2f clk per step  (%s)\n", name, (double)h / (kBlock - 40), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float *out, *coef; long long *clk;
  cudaMalloc(&out, 1 << 12); cudaMalloc(&clk, 8); cudaMalloc(&coef, 256);
  float hc[64] = {0};
  for (int i = 0; i < 20; ++i) hc[i] = 0.05f * (i % 5) - 0.1f;
  cudaMemcpy(coef, hc, 256, cudaMemcpyHostToDevice);
  run("skew 12 ahead 4 group 8 predicated store (round-2 form)", pipe<12, 4, 8, true, false, false>, out, clk, coef);
  run("skew 13 ahead 4 group 8 predicated store", pipe<13, 4, 8, true, false, false>, out, clk, coef);
  run("skew 13 ahead 4 group 8 plain store", pipe<13, 4, 8, false, false, false>, out, clk, coef);
  run("skew 13 no load", pipe<13, 4, 8, true, true, false>, out, clk, coef);
  run("skew 13 no store", pipe<13, 4, 8, true, false, true>, out, clk, coef);
  run("skew 13 no load no store", pipe<13, 4, 8, true, true, true>, out, clk, coef);
  run("skew 13 plain store delayed 1", pipe<13, 4, 8, false, false, false, 1>, out, clk, coef);
  run("skew 13 plain store delayed 2", pipe<13, 4, 8, false, false, false, 2>, out, clk, coef);
  run("skew 13 predicated store delayed 1", pipe<13, 4, 8, true, false, false, 1>, out, clk, coef);
  run("skew 13 predicated store delayed 2", pipe<13, 4, 8, true, false, false, 2>, out, clk, coef);
  run("skew 21 ahead 4 group 16", pipe<21, 4, 16, true, false, false>, out, clk, coef);
  run("skew 13 ahead 2 group 8", pipe<13, 2, 8, true, false, false>, out, clk, coef);
  run("skew 21 ahead 8 group 8", pipe<21, 8, 8, true, false, false>, out, clk, coef);
  return 0;
}
