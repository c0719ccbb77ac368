// Developer microbenchmark: issue throughput of the instruction mixes the fused RX kernel is made of
// (sm_100a).  Prints warp-instructions per cycle per SM for each mix at a given number of warps/SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  float2 v = make_float2(a, b);
  return *reinterpret_cast<unsigned long long *>(&v);
}

template <int MODE>
__global__ void thr(float *out, long long *cyc, float a, float b) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (float)(i & 15) * 1e-3f;
  __syncthreads();
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = a * (j + 1) + threadIdx.x;
  unsigned long long acc2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc2[j] = pk(a * j, b * j + threadIdx.x);
  const unsigned long long ab = pk(a, b);
  const float4 *s4 = reinterpret_cast<const float4 *>(sm);
  int base = threadIdx.x & 31;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
    if (MODE == 0) {  // 16 independent FFMA (3 distinct regs)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = fmaf(acc[j], a, b);
    }
    if (MODE == 1) {  // 8 independent FFMA2 (= 16 FMAs)
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(acc2[j]) : "l"(ab));
    }
    if (MODE == 2) {  // FIR-like: 1 LDS.128 + 16 FFMA (4 samples x 4 accumulators)
      const float4 v = s4[(base + i) & 1023];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[4 * j + 0] = fmaf(v.x, acc[(4 * j + 5) & 15], acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(v.y, acc[(4 * j + 6) & 15], acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(v.z, acc[(4 * j + 7) & 15], acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(v.w, acc[(4 * j + 8) & 15], acc[4 * j + 3]);
      }
    }
    if (MODE == 3) {  // FIR-like with FFMA2: 1 LDS.128 (2 complex samples) + 8 FFMA2
      const float4 v = s4[(base + i) & 1023];
      const unsigned long long s0 = pk(v.x, v.y), s1 = pk(v.z, v.w);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[2 * j]) : "l"(s0), "l"(ab));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[2 * j + 1]) : "l"(s1), "l"(ab));
      }
    }
    if (MODE == 4) {  // 16 FADD
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = acc[j] + b;
    }
    if (MODE == 5) {  // 8 FFMA + 8 FMNMX (alu pipe) interleaved
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] = fmaf(acc[j], a, b);
        acc[8 + j] = fmaxf(acc[8 + j], acc[j]);
      }
    }
    if (MODE == 6) {  // 4 LDS.128 only
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = s4[(base + 32 * j + i) & 1023];
        acc[4 * j] += v.x;
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float r = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) r += acc[j];
#pragma unroll
  for (int j = 0; j < 8; ++j) r += (float)(acc2[j] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, int warps, double instr_per_iter) {
  float *out;
  long long *cyc;
  const int blocks = 148;
  cudaMalloc(&out, sizeof(float) * warps * 32 * blocks);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  for (int r = 0; r < 2; ++r) thr<MODE><<<blocks, warps * 32>>>(out, cyc, 1.0001f, 0.9999f);
  cudaDeviceSynchronize();
  long long h[1];
  cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
  const double c = (double)h[0] / ITERS;
  printf("%-44s warps/SM %2d : %7.2f cyc/iter, %5.2f counted warp-instr/cyc/SM\n", name, warps, c, instr_per_iter * warps / c);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    if (w == 4) { run<0>("16 FFMA", 4, 16); run<1>("8 FFMA2", 4, 8); run<2>("LDS.128+16 FFMA", 4, 17); run<3>("LDS.128+8 FFMA2", 4, 9); run<4>("16 FADD", 4, 16); run<5>("8 FFMA+8 FMNMX", 4, 16); run<6>("4 LDS.128+4 FADD", 4, 8); }
    if (w == 8) { run<0>("16 FFMA", 8, 16); run<1>("8 FFMA2", 8, 8); run<2>("LDS.128+16 FFMA", 8, 17); run<3>("LDS.128+8 FFMA2", 8, 9); run<4>("16 FADD", 8, 16); run<5>("8 FFMA+8 FMNMX", 8, 16); run<6>("4 LDS.128+4 FADD", 8, 8); }
    if (w == 16) { run<0>("16 FFMA", 16, 16); run<1>("8 FFMA2", 16, 8); run<2>("LDS.128+16 FFMA", 16, 17); run<3>("LDS.128+8 FFMA2", 16, 9); run<4>("16 FADD", 16, 16); run<5>("8 FFMA+8 FMNMX", 16, 16); run<6>("4 LDS.128+4 FADD", 16, 8); }
    if (w == 32) { run<0>("16 FFMA", 32, 16); run<1>("8 FFMA2", 32, 8); run<2>("LDS.128+16 FFMA", 32, 17); run<3>("LDS.128+8 FFMA2", 32, 9); run<4>("16 FADD", 32, 16); run<5>("8 FFMA+8 FMNMX", 32, 16); run<6>("4 LDS.128+4 FADD", 32, 8); }
  }
  return 0;
}
