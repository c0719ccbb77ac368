/* rx_rows.cu — the rows-only kernel (display spectrum + waterfall rows beside the throughput kernel) as a unit of its own:
 * the phases of rx_phases.cuh compiled against the ROWS-ONLY SLOT LAYOUT (T41RX_ROWS_LAYOUT: raw tile + DC scratch + scalars,
 * 19.1 KB per receiver instead of the audio chain's 28.8) and a register budget that lets three CTAs share an SM.
 * Reference: ZoomFFTExe / CalcZoom1Magn / ShowSpectrum, FFT.cpp:67-251, Display.cpp:170-330 (cited phase by phase in
 * rx_phases.cuh). */
#define T41RX_ROWS_LAYOUT 1
#define T41RX_DC_BATCH 2          /* at this unit's 80 registers a batch of 4 spills inside the DC loops (C4: 38 instead of 46 Gsamples/s) */
#include <cuda_runtime.h>

#include "rx_phases.cuh"
#include "rx_launch.h"

namespace t41rx {

#ifndef T41RX_ROWS_CTAS
#define T41RX_ROWS_CTAS 3
#endif
constexpr int kRowsCtasPerSm = T41RX_ROWS_CTAS;
static_assert((size_t)kRowsCtasPerSm * (kSmemFloats * sizeof(float) + 1024) <= 233472, "three CTAs' slots fit an SM's 228 KB");

#ifdef T41RX_PHASE_TIMING
__device__ unsigned long long g_rows_phase_cycles[64];
#endif

/* display spectrum + waterfall rows of the row-producing blocks, for the receivers the throughput kernel
   serves; launched before it on the same stream (reads the launch-start state, writes only the zoom /
   spectrum state and the row outputs) */
__device__ __forceinline__ void RowsKernelBody(const LaunchArgs &a, float *smem) {
  Cta c;
  c.a = a;
  c.smem = smem;
  c.s0 = blockIdx.x * kG;
  c.ng = min(kG, a.n_streams - c.s0);
  c.row = 1;
  c.rows_only = 1;
  const int tid = threadIdx.x;
  /* The cascade phase is ONE warp per CTA issuing about an instruction per clock for 2093 steps: the cascade warps of
     the CTAs that share an SM must sit on different warp schedulers (hardware warp slot mod 4), or they take turns.
     The slots of co-resident 8-warp CTAs start at 0, 9, 16 (tools/microbench/warp_slots.cu): warp 0 of the first and of
     the third CTA would share scheduler 0.  Each CTA gives the cascade to its warp on scheduler (first slot / 8). */
  __shared__ unsigned warp_slot[kNT / 32];
  if ((tid & 31) == 0) asm("mov.u32 %0, %%warpid;" : "=r"(warp_slot[tid >> 5]));
  PhCtaInit(c, tid);
  __syncthreads();
  c.casc_warp = 0;
  {
    const unsigned want = (warp_slot[0] >> 3) & 3u;
    for (int w = 0; w < 4; ++w)
      if ((warp_slot[w] & 3u) == want) c.casc_warp = w;
  }
  /* the row-producing blocks of this launch: absolute index a multiple of row_every */
  c.dc_carried = 0;
  for (int t = (a.row_every - a.t0 % a.row_every) % a.row_every; t < a.n_blocks; t += a.row_every) {
    c.t = t;
    c.row_idx = (a.t0 + t) / a.row_every;
#ifdef T41RX_PHASE_TIMING
    /* developer build only: cycles per phase of CTA 0, g_rows_phase_cycles (slots 32.. of t41rx_debug_phase_cycles) */
    int phase_no = 0;
#define T41RX_KPHASE(stmt)                                                                   \
  do {                                                                                       \
    const long long t0_ = clock64();                                                         \
    stmt;                                                                                    \
    const long long t1_ = clock64();                                                         \
    __syncthreads();                                                                         \
    const long long t2_ = clock64();                                                         \
    if (blockIdx.x == 0 && tid == 0) {                                                       \
      g_rows_phase_cycles[2 * phase_no] += (unsigned long long)(t2_ - t0_);                       \
      g_rows_phase_cycles[2 * phase_no + 1] += (unsigned long long)(t1_ - t0_);                   \
    }                                                                                        \
    ++phase_no;                                                                              \
  } while (0)
#else
#define T41RX_KPHASE(stmt) \
  do {                     \
    stmt;                  \
    __syncthreads();       \
  } while (0)
#endif
    T41RX_ROWS_SCHEDULE_FAST(T41RX_KPHASE)
#undef T41RX_KPHASE
    c.dc_carried = a.row_every == 1;       /* the next block continues where this one ended */
  }
}

__global__ void __launch_bounds__(kNT, kRowsCtasPerSm) t41rx_rows_kernel(const LaunchArgs a) {
  extern __shared__ __align__(16) float smem[];
  RowsKernelBody(a, smem);
}
/* the same kernel with the registers of two CTAs per SM (no spills): for launches whose CTAs are all resident at two per SM
   anyway (a bank's one row per step: C2's 256 CTAs) */
__global__ void __launch_bounds__(kNT, 2) t41rx_rows_wide_kernel(const LaunchArgs a) {
  extern __shared__ __align__(16) float smem[];
  RowsKernelBody(a, smem);
}

cudaError_t ConfigureRowsKernel() {
  cudaError_t e = cudaFuncSetAttribute(t41rx_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemFloats * sizeof(float)));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(t41rx_rows_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemFloats * sizeof(float)));
}

/* rows of a.n_streams receivers (a.stream_ids / a.stream_base) on stream st */
cudaError_t LaunchRowsKernel(const LaunchArgs &a, int n_sms, cudaStream_t st) {
  const int grid = (a.n_streams + kG - 1) / kG;
  if (grid <= 2 * n_sms) t41rx_rows_wide_kernel<<<grid, kNT, kSmemFloats * sizeof(float), st>>>(a);
  else t41rx_rows_kernel<<<grid, kNT, kSmemFloats * sizeof(float), st>>>(a);
  return cudaGetLastError();
}

/* developer instrumentation of this unit (each compilation unit has its own copy of the header's counters) */
long long RowsDcRefilterCount() {
  unsigned long long v = 0;
  if (cudaMemcpyFromSymbol(&v, g_dc_refilter_count, sizeof(v)) != cudaSuccess) return -1;
  return (long long)v;
}

cudaError_t RowsPhaseCycles(unsigned long long *out64, int reset) {
#ifdef T41RX_PHASE_TIMING
  cudaError_t e = cudaMemcpyFromSymbol(out64, g_rows_phase_cycles, sizeof(unsigned long long) * 64);
  if (e != cudaSuccess) return e;
  if (reset) {
    unsigned long long z[64] = {0};
    e = cudaMemcpyToSymbol(g_rows_phase_cycles, z, sizeof(z));
  }
  return e;
#else
  (void)out64;
  (void)reset;
  return cudaSuccess;
#endif
}

}  // namespace t41rx
