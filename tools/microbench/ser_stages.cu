// Developer microbenchmark: the three serial stages of the bit-exact chain (SerAgc / SerGain / SerDemod of rx_phases.cuh) alone,
// one warp, lane = receiver, clocks per sample; rolled loop with the kernel's unroll factor.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -I t41_sdr_b200/csrc -I include -o tools/microbench/_bin/ser_stages tools/microbench/ser_stages.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "rx_phases.cuh"
using namespace t41rx;

template <int kStage, int kUnroll>
__global__ void k(LaunchArgs a, const float *sin_table, long long *clk, float *sink, int n) {
  __shared__ float sin_tab[513];
  __shared__ float2 zin[64][32];
  for (int i = threadIdx.x; i < 513; i += 32) sin_tab[i] = sin_table[i];
  for (int i = 0; i < 64; ++i) zin[i][threadIdx.x] = float2{0.3f * __sinf(0.37f * i + threadIdx.x), 0.3f * __cosf(0.37f * i + 0.1f * threadIdx.x)};
  __syncwarp();
  const StreamCfg &cf = a.cfg[0];
  StreamState &st = a.st[0];
  float acc = 0.0f;
  long long t0, t1;
  if (kStage == 0) {
    SerAgc agc; agc.Load(cf, st);
    agc.state = threadIdx.x % 5;            /* receivers in different states, as in a bank */
    t0 = clock64();
#pragma unroll kUnroll
    for (int i = 0; i < n; ++i) { const float2 z = zin[i & 63][threadIdx.x]; acc += agc.Step(fabsf(z.x), fabsf(z.y) + 0.1f); }
    t1 = clock64();
  } else if (kStage == 1) {
    SerGain g; g.Load(cf);
    t0 = clock64();
#pragma unroll kUnroll
    for (int i = 0; i < n; ++i) { const float2 z = zin[i & 63][threadIdx.x]; acc += g.Mult(fabsf(z.x) + 0.01f); }
    t1 = clock64();
  } else {
    SerDemod d; d.Load(a, cf, st, sin_tab); d.mode = kModeSam; d.BlockStart();
    t0 = clock64();
#pragma unroll kUnroll
    for (int i = 0; i < n; ++i) { const float2 z = zin[i & 63][threadIdx.x]; acc += d.Step(z.x, z.y); }
    t1 = clock64();
  }
  sink[threadIdx.x] = acc;
  if (threadIdx.x == 0) *clk = t1 - t0;
}

int main() {
  StreamCfg hc; memset(&hc, 0, sizeof(hc));
  StreamState hs; memset(&hs, 0, sizeof(hs));
  hc.mode = kModeSam; hc.agc_mode = 3;
  hc.agc.fast_backmult = 0.1f; hc.agc.onemfast_backmult = 0.9f; hc.agc.hang_backmult = 0.05f; hc.agc.onemhang_backmult = 0.95f;
  hc.agc.attack_mult = 0.3f; hc.agc.decay_mult = 0.001f; hc.agc.fast_decay_mult = 0.01f; hc.agc.hang_decay_mult = 0.002f;
  hc.agc.pop_ratio = 5.0f; hc.agc.hang_level = 0.3f; hc.agc.min_volts = 1e-4f; hc.agc.hang_counter_load = 100; hc.agc.hang_enable = 1;
  hc.agc.inv_max_input = 1.0f; hc.agc.out_target = 1.0f; hc.agc.slope_constant = 0.1f;
  hs.agc_volts = 0.2f;
  float hsam[4] = {-0.1f, 0.1f, 0.02f, 0.0005f};
  std::vector<float> tab(513);
  for (int i = 0; i < 513; ++i) tab[i] = (float)sin(2.0 * 3.14159265358979 * i / 512.0);
  StreamCfg *dc; StreamState *ds; float *dsam, *dtab, *sink; long long *clk, h;
  cudaMalloc(&dc, sizeof(hc)); cudaMalloc(&ds, sizeof(hs)); cudaMalloc(&dsam, 16); cudaMalloc(&dtab, 513 * 4); cudaMalloc(&sink, 256); cudaMalloc(&clk, 8);
  cudaMemcpy(dc, &hc, sizeof(hc), cudaMemcpyHostToDevice); cudaMemcpy(ds, &hs, sizeof(hs), cudaMemcpyHostToDevice);
  cudaMemcpy(dsam, hsam, 16, cudaMemcpyHostToDevice); cudaMemcpy(dtab, tab.data(), 513 * 4, cudaMemcpyHostToDevice);
  LaunchArgs a; memset(&a, 0, sizeof(a));
  a.cfg = dc; a.st = ds; a.sam_consts = dsam; a.n_streams = 1;
  const int n = 4096;
  const char *names[] = {"AGC envelope (lanes in different states)", "gain from volts", "SAM PLL"};
  for (int s = 0; s < 3; ++s) {
    for (int u = 0; u < 2; ++u) {
      for (int rep = 0; rep < 2; ++rep) {
        if (s == 0 && u == 0) k<0, 4><<<1, 32>>>(a, dtab, clk, sink, n);
        if (s == 0 && u == 1) k<0, 1><<<1, 32>>>(a, dtab, clk, sink, n);
        if (s == 1 && u == 0) k<1, 4><<<1, 32>>>(a, dtab, clk, sink, n);
        if (s == 1 && u == 1) k<1, 1><<<1, 32>>>(a, dtab, clk, sink, n);
        if (s == 2 && u == 0) k<2, 4><<<1, 32>>>(a, dtab, clk, sink, n);
        if (s == 2 && u == 1) k<2, 1><<<1, 32>>>(a, dtab, clk, sink, n);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
      printf("%-44s unroll %d: %.1f clk per sample (%s)\n", names[s], u ? 1 : 4, (double)h / n, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
