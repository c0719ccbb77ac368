#!/usr/bin/env python3
"""Generate tests/golden/ref_vectors.npz from the REFERENCE ITSELF (Tier-A).

Runs every case of tests/cases.py through oracle/_ref/libt41ref.so — the reference's own
hot-path translation units compiled in place from /root/reference (oracle/Makefile `ref`) —
and stores, per receiver:
  * sha256 of the full float32 audio output (the exact check),
  * every 61st audio sample (for a readable SNR when a digest mismatches),
  * all spectrum / waterfall rows, audio-spectrum rows (audioYPixel) and audioMaxSquaredAve values, PSK31 bit
    decisions and characters,
  * the debug scalars the reference exposes with external linkage.
Inputs are not stored: they are regenerated from the seeded generators in
t41_sdr_b200/synth.py.  Only runnable where /root/reference is mounted (this container):
    make -C oracle ref && python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import oracle_py as O  # noqa: E402

STRIDE = 61
REF_DEBUG_FIELDS = ("agc_hang_counter", "agc_action", "rf_gain", "zoom_sample_ptr", "first_block",
                    "sam_phzerror", "sam_omega2", "sam_fil_out", "osc_vect_q", "osc_vect_i")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    if not O.tier_a_available():
        sys.exit("oracle/_ref/libt41ref.so missing: run `make -C oracle ref` where /root/reference is mounted")
    out = {}
    for make in cases.ALL_CASES:
        case = make()
        res = cases.run_case_on(case, lambda p: O.RefStream(p))
        for s, r in enumerate(res):
            key = "%s/%d/" % (case.name, s)
            out[key + "audio_sha256"] = np.frombuffer(bytes.fromhex(digest(r["audio"])), np.uint8)
            out[key + "audio_sub"] = r["audio"].ravel()[::STRIDE].copy()
            out[key + "spec"] = r["spec"]
            out[key + "wf"] = r["wf"]
            out[key + "audio_ypixel"] = r["audio_ypixel"].astype(np.int16)      # 0..~200
            out[key + "audio_max_sq_ave"] = r["audio_max_sq_ave"]
            out[key + "spec_frames"] = r["spec_frames"]                         # control-port writes per row
            out[key + "audio_frames"] = r["audio_frames"]
            if case.psk:
                out[key + "psk_bits"] = r["psk_bits"]
                out[key + "psk_chars"] = r["psk_chars"]
            d = r["debug"]
            out[key + "debug"] = np.array([float(getattr(d, f)) for f in REF_DEBUG_FIELDS], np.float64)
        print("%-26s receivers=%d blocks=%d" % (case.name, case.n_streams, case.n_blocks))
    # the firmware's WAV reader (Utility.cpp:773-888) on the files of tests/wav_cases.py
    import tempfile
    import wav_cases
    with tempfile.TemporaryDirectory() as tmp:
        for name, (data, limit, chunk) in wav_cases.cases().items():
            p = os.path.join(tmp, name + ".wav")
            if data is not None:
                open(p, "wb").write(data)
            ref = O.RefStream()
            rc, n, full, tail = wav_cases.run(ref.load_wav, ref.read_wave, p, data, limit, chunk)
            out["wav/%s/rc_n" % name] = np.array([rc, n], np.int64)
            out["wav/%s/full" % name] = full
            out["wav/%s/tail" % name] = tail
            print("wav %-24s rc=%d reads=%d" % (name, rc, n))
    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KiB" % (os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
