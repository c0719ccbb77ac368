// Developer microbenchmark: the DC-block speculation phases (PhDcWarm / PhDcMain of rx_phases.cuh) alone, one CTA of the
// product's shape (4 receivers x 64 chunk lanes), clocks per warp.  Variants by -DT41RX_DC_BATCH=..., -DWARPS=... (how many of
// the CTA's 8 warps run), -DCTAS=... (CTAs on the SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -I t41_sdr_b200/csrc -I include -o tools/microbench/_bin/dc_loop tools/microbench/dc_loop.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "rx_phases.cuh"
using namespace t41rx;

#ifndef WARPS
#define WARPS 8
#endif

__global__ void __launch_bounds__(kNT, 2) k(long long *clk, float *sink, int main_pass, float rfg) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  for (int i = tid; i < kSmemFloats; i += kNT) smem[i] = 0.001f * (float)((i * 37) % 101) - 0.05f;
  __syncthreads();
  const int g = tid / kDcChunks, ch = tid % kDcChunks;
  float *s = smem + g * kSlot;
  const DcPost p{rfg, 3.0f, -1.01f, false};
  float d1 = 0.0f, d2 = 0.0f;
  long long t0 = clock64();
  if ((tid >> 5) < WARPS && ch > 0) {
    if (!main_pass) {
      int start = ch * kDcChunkLen - kDcSpecWarm;
      if (start < 0) start = 0;
#ifdef UNIFORM
      start = (ch % 28) * kDcChunkLen;              /* every lane the same 192 steps, all in the I block */
      DcRun<false>(s, p, start, start + kDcSpecWarm, d1, d2);
#else
      DcRun<false>(s, p, start, ch * kDcChunkLen, d1, d2);
#endif
    } else {
      const int begin = ch * kDcChunkLen;
      const int end = (ch == kDcChunks - 1) ? 2 * kBlock : begin + kDcChunkLen;
      DcRun<true>(s, p, begin, end, d1, d2);
    }
  }
  long long t1 = clock64();
  sink[blockIdx.x * kNT + tid] = d1 + d2;
  if ((tid & 31) == 0) clk[blockIdx.x * 8 + (tid >> 5)] = t1 - t0;
}

int main(int argc, char **argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 1;
  long long *clk, h[64];
  float *sink;
  cudaMalloc(&clk, 64 * 8);
  cudaMalloc(&sink, 1 << 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemFloats * sizeof(float)));
  for (int main_pass = 0; main_pass < 2; ++main_pass) {
    for (int rep = 0; rep < 3; ++rep) {
      k<<<ctas * 148, kNT, kSmemFloats * sizeof(float)>>>(clk, sink, main_pass, 1.5f);   /* ctas per SM */
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
    printf("%s, batch %d, %d warps, %d CTA(s) per SM: clocks per warp of CTA 0:", main_pass ? "DcMain" : "DcWarm", T41RX_DC_BATCH, WARPS, ctas);
    for (int w = 0; w < 8; ++w) printf(" %lld", h[w]);
    printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
