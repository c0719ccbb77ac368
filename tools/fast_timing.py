#!/usr/bin/env python3
"""Developer tool: per-section cycle breakdown of t41rx_stream_rx_kernel (CTA 0: receiver warp 0 and the
AGC warp).  Needs a library built with -DT41RX_FAST_TIMING (tools/build_fast_timing.sh) selected via T41RX_LIB."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
import rx_driver  # noqa: E402
from t41_sdr_b200 import rx  # noqa: E402

NAMES = ["fe:setup+tail", "fe:wait cp.async", "fe:dc+mix", "fe:dec1", "fe:save hist", "fe:dec2", "fe:fft filter",
         "fe:|z|+winmax", "be:gain+demod", "be:interp1", "be:interp2", "codec+arrive", "wait AGC done", "", "", "",
         "agc:work", "agc:wait FE done"]


def main():
    S = int(os.environ.get("STREAMS", "1024"))
    T = int(os.environ.get("BLOCKS", "32"))
    rows = int(os.environ.get("ROWS", "0"))
    params, sigs = bench.workload(T)
    eng = rx.Receiver(S)
    eng.set_params_each([rx_driver.to_rx_params(params[s % 16]) for s in range(S)])
    dev = torch.device("cuda", 0)
    base = torch.from_numpy(np.stack(sigs)).to(dev)
    iq = base.index_select(0, torch.arange(S, device=dev) % 16).contiguous()
    audio = torch.empty((S, T, 2048), dtype=torch.float32, device=dev)
    spec = torch.empty((S, T, 512), dtype=torch.int16, device=dev)
    wf = torch.empty((S, T, 512), dtype=torch.int16, device=dev)
    L = rx.lib()
    buf = (C.c_ulonglong * 32)()
    for it in range(3):
        eng.process_device(iq.data_ptr(), audio.data_ptr(), T, rows, spec.data_ptr(), wf.data_ptr())
        eng.synchronize()
        L.t41rx_debug_fast_cycles(buf, 1)
    v = np.array(buf[:], dtype=np.float64) / T
    tot = v[:13].sum()
    print("cycles per block (CTA 0, receiver warp 0): total %.0f; kernel %.3f ms = %.0f cycles per block at 1.965 GHz" % (
        tot, eng.last_kernel_ms(), eng.last_kernel_ms() * 1e-3 * 1.965e9 / (T + 2)))
    for i, n in enumerate(NAMES):
        if n:
            print("  %-18s %8.0f  %5.1f%%" % (n, v[i], 100 * v[i] / (tot if i < 16 else v[16:18].sum())))
    print("  agc fallback chunks per block (of 32): %.2f" % v[20])


if __name__ == "__main__":
    main()
