#!/usr/bin/env python3
"""Developer tool: run parity cases through the throughput kernel (flags = 0) and print, per receiver,
the audio SNR against the CPU oracle (whole run and worst block), spectrum-row differences and debug-state
mismatches.  usage: tools/try_fast.py [case_name ...] [--norows]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle_py as O  # noqa: E402
import rx_driver  # noqa: E402
from t41_sdr_b200 import rx  # noqa: E402


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")]
    norows = "--norows" in sys.argv
    flags = 2 if "--phased" in sys.argv else 0
    for a_ in sys.argv[1:]:
        if a_.startswith("--flags="):
            flags = int(a_.split("=")[1])
    for make in cases.ALL_CASES:
        if names and make.__name__ not in names:
            continue
        case = make()
        if norows:
            case.row_every = 0
        want = cases.run_case_on(case, lambda p: O.OracleStream(p))
        with rx.Receiver(case.n_streams) as eng:
            got = rx_driver.run_case_batched(case, eng, flags=flags)
        print("== %s (%d receivers, %d blocks, row_every %d)" % (case.name, case.n_streams, case.n_blocks, case.row_every))
        for s, (g, w) in enumerate(zip(got, want)):
            ga, wa = g["audio"], w["audio"]
            ok = ~(np.isnan(ga) | np.isnan(wa))
            snr = O.snr_db(np.where(ok, wa, 0), np.where(ok, ga, 0))
            blk = [O.snr_db(np.where(ok[b], wa[b], 0), np.where(ok[b], ga[b], 0)) for b in range(ga.shape[0])]
            worst = int(np.argmin(blk))
            msg = "  rx %2d mode %d agc %d: SNR %6.1f dB, worst block %d (%.1f dB), first blocks %s" % (
                s, case.segments[0][0][s].mode, case.segments[0][0][s].agc_mode, snr, worst, blk[worst],
                " ".join("%.0f" % b for b in blk[:40]))
            if w["spec"].size:
                d = np.abs(g["spec"].astype(int) - w["spec"].astype(int))
                msg += " | spec maxdiff %d same %.4f" % (d.max(), np.mean(d == 0))
            if case.psk:
                msg += " | psk bits %s chars %s" % (np.array_equal(g["psk_bits"], w["psk_bits"]),
                                                   np.array_equal(g["psk_chars"], w["psk_chars"]))
            bad = [f for f in cases.DEBUG_INT_FIELDS if getattr(g["debug"], f) != getattr(w["debug"], f)]
            if bad:
                msg += " | STATE MISMATCH " + ",".join("%s %s!=%s" % (f, getattr(g["debug"], f), getattr(w["debug"], f)) for f in bad)
            print(msg)


if __name__ == "__main__":
    main()
