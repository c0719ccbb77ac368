#!/usr/bin/env python3
"""Developer tool: per-phase cycle breakdown of the fused RX kernel (CTA 0, thread 0).
Needs a library built with -DT41RX_PHASE_TIMING (tools/build_phase_timing.sh) selected via T41RX_LIB;\nenv: ROWS (row period, 0 = none), FLAGS (t41rx_process flags, default 2), bench.py's T41RX_BENCH_* workload knobs."""
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

NAMES = ["Load", "DcWarm", "DcMain", "DcVerify", "DcFix", "NcoPrep", "Mix", "Dec1", "Dec2", "PostDec2", "NfmAsm", "NfmAsm2",
         "FftA0", "FftA1", "FftA2", "Mask", "FftB0", "FftB1", "FftB2", "AgcPre", "MaxA", "MaxB", "MaxC", "AgcSerial", "AgcPost", "DemodPar", "DemodSer", "EqBands", "EqSum", "NrStageIn", "NrNotch", "NrStageOut",
         "CwFilter", "Interp1b", "Interp2", "BlockEnd"]
FRONT_TAIL = ["SerialStore", "SerialRing+CodecGain"]      # the split chain's front kernel (FLAGS without 32) after MaxC
ROW_NAMES = ["ZoomIir", "SpecWin", "SpecFft0", "SpecFft1", "SpecFft2", "SpecRow"]


def main():
    S, T = int(os.environ.get("PHASE_TIMING_S", "1024")), 16
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
    buf = (C.c_ulonglong * 128)()
    for it in range(3):
        eng.process_device(iq.data_ptr(), audio.data_ptr(), T, rows, spec.data_ptr(), wf.data_ptr(), None, None,
                           int(os.environ.get("FLAGS", "2")))      # 2: the bit-exact kernel with the closed-form oscillator
        eng.synchronize()
        L.t41rx_debug_phase_cycles(buf, 1)
    v = np.array(buf[:], dtype=np.float64).reshape(64, 2) / T
    flags = int(os.environ.get("FLAGS", "2"))
    body = NAMES[5:] if (flags & 32) else NAMES[5:NAMES.index("MaxC") + 1] + FRONT_TAIL
    names = NAMES[:5] + (ROW_NAMES if rows else []) + body
    tot = v[:, 0].sum()
    print("cycles per block-group (CTA 0), total %.0f, kernel %.3f ms" % (tot, eng.last_kernel_ms()))
    order = sorted(range(len(names)), key=lambda i: -v[i, 0])
    for i in order:
        n = names[i]
        if v[i, 0] < 0.004 * tot:
            continue
        print("%-10s total %8.0f  work(thread0) %8.0f  %5.1f%%" % (n, v[i, 0], v[i, 1], 100 * v[i, 0] / tot))


if __name__ == "__main__":
    main()
