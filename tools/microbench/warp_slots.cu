// Developer microbenchmark: which hardware warp slots (%warpid) do the warps of two co-resident 256-thread CTAs get?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/_bin/warp_slots tools/microbench/warp_slots.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) k(unsigned *out) {
  extern __shared__ float smem[];
  unsigned wid, sm;
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  smem[threadIdx.x] = wid;
  long long t0 = clock64();
  while (clock64() - t0 < 200000) {}
  if ((threadIdx.x & 31) == 0) {
    out[(blockIdx.x * 8 + (threadIdx.x >> 5)) * 2] = sm;
    out[(blockIdx.x * 8 + (threadIdx.x >> 5)) * 2 + 1] = wid;
  }
}
int main() {
  unsigned *out, h[296 * 16];
  cudaMalloc(&out, sizeof(h));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 115328);
  k<<<296, 256, 115328>>>(out);
  cudaDeviceSynchronize();
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  for (int b = 0; b < 296; ++b) {
    if (h[b * 16] > 2) continue;
    printf("CTA %3d on SM %u: warp slots", b, h[b * 16]);
    for (int w = 0; w < 8; ++w) printf(" %u", h[(b * 8 + w) * 2 + 1]);
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
