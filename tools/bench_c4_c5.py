#!/usr/bin/env python3
"""Developer tool: device-resident throughput of workloads shaped like BASELINE.json configs[3] (C4: 16384 receivers,
zoom x1..x16, every block producing a spectrum + waterfall row) and configs[4] (C5: 32768 receivers, PSK31 front end with
the DBPSK + varicode tap).  Wall-clock around synchronised launches; bench.py stays the C2 metric."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, rx_driver
from t41_sdr_b200 import rx, synth

def run(name, S, T, params, sigs, row_every, psk):
    D = len(params)
    iq1 = np.stack(sigs)
    iq = torch.from_numpy(iq1).cuda().repeat((S + D - 1) // D, 1, 1, 1)[:S].contiguous()
    R = (T + row_every - 1) // row_every if row_every else 0
    audio = torch.empty((S, T, 2048), dtype=torch.float32, device="cuda")
    spec = torch.empty((S, max(R, 1), 512), dtype=torch.int16, device="cuda")
    wf = torch.empty((S, max(R, 1), 512), dtype=torch.int16, device="cuda")
    bits = torch.empty((S, T), dtype=torch.int8, device="cuda")
    chars = torch.empty((S, T), dtype=torch.uint8, device="cuda")
    with rx.Receiver(S) as eng:
        eng.set_params_each([rx_driver.to_rx_params(params[s % D]) for s in range(S)])
        def step():
            eng.process_device(iq.data_ptr(), audio.data_ptr(), T, row_every, spec.data_ptr() if R else None, wf.data_ptr() if R else None,
                               bits.data_ptr() if psk else None, chars.data_ptr() if psk else None)
        for _ in range(3): step()
        eng.synchronize()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n): step()
        eng.synchronize()
        dt = (time.perf_counter() - t0) / n
    print("%s: %d receivers x %d blocks: %.2f ms/step, %.0f Msamples/s, %.2f M rows/s" % (name, S, T, dt * 1e3, S * T * 2048 / dt / 1e6, S * R / dt / 1e6))

T = 32
p4 = [cases.P(mode=cases.USB, spectrum_zoom=z, current_scale=1) for z in range(5)]
s4 = [synth.two_tone(40 + z, T, 46500.0, 50500.0) for z in range(5)]
run("C4 (zoom x1..x16, every block a row)", 16384, T, p4, s4, 1, False)
T = 24
p5 = [cases.P(mode=cases.USB, f_lo_cut=-100, f_hi_cut=100, agc_mode=0, psk31_enable=1) for _ in range(4)]
s5 = [synth.tone(50 + k, T, 0.0) for k in range(4)]
run("C5 (PSK31 front end + DBPSK + varicode)", 32768, T, p5, s5, 0, True)
