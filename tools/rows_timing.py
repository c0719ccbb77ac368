#!/usr/bin/env python3
"""Developer tool: cycles per phase of t41rx_rows_kernel (CTA 0) on the bench workload; needs tools/build_phase_timing.sh.
usage: T41RX_LIB=t41_sdr_b200/libt41rx_ptiming.so tools/rows_timing.py [zoom]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import rx_driver  # noqa: E402
from t41_sdr_b200 import rx, synth  # noqa: E402

NAMES = ["Load+TailLoad", "RowDcSeedFast", "DcWarm", "DcMain", "DcVerify", "DcFix", "ZoomShift", "ZoomIir", "ZoomDecimate",
         "ZoomDecimateEnd", "SpecWindow", "SpecFft0", "SpecFft1", "SpecFft2", "SpecRow"]


def main():
    zoom = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    S, T = int(os.environ.get("ROWS_TIMING_S", "1024")), 8
    p = cases.P(mode=cases.USB, spectrum_zoom=zoom)
    iq1 = synth.tone(5, T, 1000.0)
    iq = torch.from_numpy(np.broadcast_to(iq1, (S,) + iq1.shape).copy()).cuda()
    audio = torch.empty((S, T, 2048), dtype=torch.float32, device="cuda")
    spec = torch.empty((S, T, 512), dtype=torch.int16, device="cuda")
    wf = torch.empty((S, T, 512), dtype=torch.int16, device="cuda")
    L = rx.lib()
    buf = (C.c_ulonglong * 128)()
    with rx.Receiver(S) as eng:
        eng.set_params(rx_driver.to_rx_params(p))
        for it in range(3):
            if it == 2:
                L.t41rx_debug_phase_cycles(buf, 1)
            eng.process_device(iq.data_ptr(), audio.data_ptr(), T, 1, spec.data_ptr(), wf.data_ptr())
            eng.synchronize()
        L.t41rx_debug_phase_cycles(buf, 0)
    tot = sum(buf[2 * (32 + i)] for i in range(len(NAMES)))
    for i, n in enumerate(NAMES):
        print("%-16s %9.0f clk/row (work %9.0f)  %5.1f %%" % (n, buf[2 * (32 + i)] / T, buf[2 * (32 + i) + 1] / T,
                                                             100.0 * buf[2 * (32 + i)] / max(tot, 1)))
    print("total %.0f clk/row" % (tot / T))


if __name__ == "__main__":
    main()
