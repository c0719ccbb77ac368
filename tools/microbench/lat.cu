// Developer microbenchmark: dependent-chain latencies and throughputs on sm_100a (cycles per op).
#include <cstdio>
#include <cuda_runtime.h>

#define N 4096
template <int MODE>
__global__ void chain(float *out, long long *cyc, float a, float b, double da, double db) {
  float x = a;
  double xd = da;
  __shared__ float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (float)((i * 7 + 1) & 1023);
  __syncthreads();
  int idx = threadIdx.x & 1023;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (MODE == 0) x = __fadd_rn(x, b);
    if (MODE == 1) x = __fmul_rn(x, b);
    if (MODE == 2) x = fmaf(x, b, a);
    if (MODE == 3) xd = __dadd_rn(xd, db);
    if (MODE == 4) xd = __dmul_rn(xd, db);
    if (MODE == 5) xd = fma(xd, db, da);
    if (MODE == 6) { x = (float)((double)x + db); }                 // F2D, DADD, D2F
    if (MODE == 7) { idx = (int)sm[idx]; }                           // LDS + F2I dependent
    if (MODE == 8) { x = (x > b) ? x * a : x + a; }                  // compare + select + op
    if (MODE == 9) { x = __fsqrt_rn(x + b); }
    if (MODE == 10) { x = __fdiv_rn(a, x + b); }
    if (MODE == 11) { xd = (double)(float)xd + db; }                 // D2F, F2D, DADD
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + (float)xd + (float)idx;
}

template <int MODE>
void run(const char *name, int threads, int blocks) {
  float *out; long long *cyc;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  chain<MODE><<<blocks, threads>>>(out, cyc, 1.0001f, 0.9999f, 1.0000001, 0.99999999);
  chain<MODE><<<blocks, threads>>>(out, cyc, 1.0001f, 0.9999f, 1.0000001, 0.99999999);
  cudaDeviceSynchronize();
  long long h[1];
  cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
  printf("%-28s threads/blk %4d blocks %4d : %7.2f cycles per iteration per warp-step\n", name, threads, blocks, (double)h[0] / N);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  // latency: one warp
  run<0>("FADD dependent", 32, 1);
  run<1>("FMUL dependent", 32, 1);
  run<2>("FFMA dependent", 32, 1);
  run<3>("DADD dependent", 32, 1);
  run<4>("DMUL dependent", 32, 1);
  run<5>("DFMA dependent", 32, 1);
  run<6>("F2D+DADD+D2F dependent", 32, 1);
  run<11>("D2F+F2D+DADD dependent", 32, 1);
  run<7>("LDS+F2I dependent", 32, 1);
  run<8>("cmp+sel+op dependent", 32, 1);
  run<9>("FADD+sqrt dependent", 32, 1);
  run<10>("FADD+fdiv dependent", 32, 1);
  // throughput: full SM occupancy (32 warps/SM x 148*2 blocks): cycles per iteration seen by one warp
  run<3>("DADD 1024 thr (thrpt)", 1024, 296);
  run<5>("DFMA 1024 thr (thrpt)", 1024, 296);
  run<6>("F2D+DADD+D2F 1024 thr", 1024, 296);
  run<2>("FFMA 1024 thr (thrpt)", 1024, 296);
  run<0>("FADD 1024 thr (thrpt)", 1024, 296);
  return 0;
}
