/*
 * rx_fast.cu — t41rx_stream_rx_kernel (rx_fast.cuh) and its launcher.  Built WITHOUT --fmad=false:
 * this kernel trades bit-exactness for throughput (see rx_fast.cuh); the bit-exact kernel lives in
 * rx_api.cu.
 */
#include <cuda_runtime.h>
#include <stdlib.h>

#include "rx_fast.cuh"
#include "rx_launch.h"

namespace t41rx {

__global__ void __launch_bounds__(64 * fast::kFastMaxG + 32, 1) t41rx_stream_rx_kernel(const LaunchArgs a, const int G) {
  extern __shared__ __align__(16) float smem[];
  fast::StreamKernelBody<false>(a, G, smem);
}
/* the same kernel on the firmware's q15 block format (a.iq16 in, a.audio16 out) */
__global__ void __launch_bounds__(64 * fast::kFastMaxG + 32, 1) t41rx_stream_rx_q15_kernel(const LaunchArgs a, const int G) {
  extern __shared__ __align__(16) float smem[];
  fast::StreamKernelBody<true>(a, G, smem);
}

int StreamKernelMaxReceiversPerCta() { return fast::kFastMaxG; }

#ifdef T41RX_FAST_TIMING
extern "C" int t41rx_debug_fast_cycles(unsigned long long *out32, int reset) {
  if (cudaMemcpyFromSymbol(out32, fast::g_fast_cycles, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(fast::g_fast_cycles, z, sizeof(z));
  }
  return 0;
}
#endif

cudaError_t ConfigureStreamKernel() {
  const cudaError_t e = cudaFuncSetAttribute(t41rx_stream_rx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(fast::kFastMaxG * fast::kSlotF * sizeof(float) + 32));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(t41rx_stream_rx_q15_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(fast::kFastMaxG * fast::kSlotF * sizeof(float) + 32));
}

cudaError_t LaunchStreamKernel(const LaunchArgs &a, int n_sms, cudaStream_t st) {
  /* receivers per CTA: as few as fill every SM once (one CTA per SM: 1024 receivers on 148 SMs -> 7),
     capped by shared memory; larger banks run in waves */
  int G = (a.n_streams + n_sms - 1) / n_sms;
  if (G < 1) G = 1;
  if (G > fast::kFastMaxG) G = fast::kFastMaxG;
#ifdef T41RX_DEV_KNOBS
  if (const char *e = getenv("T41RX_FAST_G")) {          /* developer builds only: receivers per CTA */
    const int g = atoi(e);
    if (g >= 1 && g <= fast::kFastMaxG) G = g;
  }
#endif
  const int grid = (a.n_streams + G - 1) / G;
  if (a.iq16) t41rx_stream_rx_q15_kernel<<<grid, 64 * G + 32, (size_t)G * fast::kSlotF * sizeof(float) + 32, st>>>(a, G);
  else t41rx_stream_rx_kernel<<<grid, 64 * G + 32, (size_t)G * fast::kSlotF * sizeof(float) + 32, st>>>(a, G);
  return cudaGetLastError();
}

}  // namespace t41rx
