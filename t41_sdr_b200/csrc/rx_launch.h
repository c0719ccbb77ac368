/* rx_launch.h — launcher of the throughput kernel (rx_fast.cu), called from the C-ABI (rx_api.cu). */
#ifndef T41RX_LAUNCH_H
#define T41RX_LAUNCH_H

#include <cuda_runtime.h>

namespace t41rx {

struct LaunchArgs;

cudaError_t ConfigureStreamKernel();
cudaError_t LaunchStreamKernel(const LaunchArgs &a, int n_sms, cudaStream_t st);
int StreamKernelMaxReceiversPerCta();

}  // namespace t41rx
#endif
