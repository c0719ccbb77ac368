#!/usr/bin/env python3
"""Developer tool: all-SAM bank through the bit-exact chain, then how many blocks had to re-run their input conditioning
serially (t41rx_dc_refilter_count): the miss rate of the time-parallel DC-block speculation for the library selected by
T41RX_LIB (built with -DT41RX_DC_SPEC_WARM=<n>)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from t41_sdr_b200 import rx
S, T = 4096, 32
os.environ["T41RX_BENCH_MODE"] = "8"
params, sigs = bench.workload(T)
iq = torch.from_numpy(np.stack(sigs)).cuda().repeat(S // len(sigs), 1, 1, 1).contiguous()
audio = torch.empty((S, T, 2048), dtype=torch.float32, device="cuda")
with rx.Receiver(S) as eng:
    eng.set_params_each([params[s % len(params)] for s in range(S)])
    for _ in range(2):
        eng.process_device(iq.data_ptr(), audio.data_ptr(), T)
    eng.synchronize()
    c0 = rx.lib().t41rx_dc_refilter_count()
    t0 = time.perf_counter()
    n = 4
    for _ in range(n):
        eng.process_device(iq.data_ptr(), audio.data_ptr(), T)
    eng.synchronize()
    dt = (time.perf_counter() - t0) / n
    c1 = rx.lib().t41rx_dc_refilter_count()
print("%s: %.2f ms/step, %.0f Msamples/s, refiltered blocks %d of %d (%.2e)" % (
    os.environ.get("T41RX_LIB", "default"), dt * 1e3, S * T * 2048 / dt / 1e6, c1 - c0, n * S * T, (c1 - c0) / (n * S * T)))
