"""Parity cases shared by the golden generator, the CPU tier and the GPU tier.

Each case is a list of receivers (params + seeded synthetic I/Q) plus how it is driven
(row cadence, PSK tap, optional mid-run parameter changes).  They are scaled-down versions
of BASELINE.json's configs C1-C5 plus the edge cases a streaming
DSP chain has: silence, full scale, one block per call, ragged receiver counts, retunes,
mode / AGC / zoom changes between calls.
"""
import numpy as np

import oracle_py as O
from t41_sdr_b200 import synth

USB, LSB, AM, NFM, PSK31, SAM = O.DEMOD_USB, O.DEMOD_LSB, O.DEMOD_AM, O.DEMOD_NFM, O.DEMOD_PSK31, O.DEMOD_SAM


def P(**kw):
    p = O.default_params()
    if "mode" in kw and "f_lo_cut" not in kw and "f_hi_cut" not in kw:
        kw["f_lo_cut"], kw["f_hi_cut"] = O.mode_default_cuts(kw["mode"])
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class Case:
    """segments: list of (params_list, n_blocks).  Parameters change between segments
    (t41rx_set_params_each), I/Q of a receiver is one continuous seeded signal."""

    def __init__(self, name, segments, iq, row_every=0, psk=False, note=""):
        self.name = name
        self.segments = segments
        self.iq = iq                      # list of [T_total, 2048, 2] arrays, one per receiver
        self.row_every = row_every
        self.psk = psk
        self.note = note

    @property
    def n_streams(self):
        return len(self.iq)

    @property
    def n_blocks(self):
        return sum(n for _, n in self.segments)


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def c1_single_usb():
    """C1: one USB receiver, 2.7 kHz filter (+300..+3000), AGC Long, tone at +1 kHz audio."""
    p = P(mode=USB, f_lo_cut=300, f_hi_cut=3000, agc_mode=1, spectrum_zoom=1)
    return Case("c1_usb_agc_long", [([p], 48)], [synth.tone(1, 48, 1000.0)], row_every=16)


def c1_single_usb_agc_off():
    p = P(mode=USB, f_lo_cut=300, f_hi_cut=3000, agc_mode=0)
    return Case("c1_usb_agc_off", [([p], 24)], [synth.tone(2, 24, 1000.0)], row_every=24)


def c2_ssb_am_mix(n=6, T=20):
    """C2: even receivers USB, odd AM, per-receiver tone offset and NCO frequency."""
    r = _rng(202)
    ps, iqs = [], []
    for s in range(n):
        nco = int(r.integers(-20000, 20001))
        if s % 2 == 0:
            f = float(r.uniform(300, 2700))
            ps.append(P(mode=USB, f_lo_cut=300, f_hi_cut=3000, nco_freq=nco))
            iqs.append(synth.tone(100 + s, T, f, mode=USB, nco_freq=nco))
        else:
            ps.append(P(mode=AM, nco_freq=nco))
            iqs.append(synth.am(100 + s, T, mode=AM, nco_freq=nco, depth=0.5, f_mod=400.0))
    return Case("c2_ssb_am_mix", [(ps, T)], iqs, row_every=T)


def c3_nfm_sam_agc(n=8, T=40):
    """C3: even NFM, odd SAM, AGC mode cycling 1..4, 20 dB level step to force AGC transitions."""
    r = _rng(303)
    ps, iqs = [], []
    for s in range(n):
        agc = 1 + (s % 4)
        if s % 2 == 0:
            ps.append(P(mode=NFM, agc_mode=agc, nfm_filter_bw=12000))
            iqs.append(synth.nfm(300 + s, T, level_step_block=T // 2, level_step_db=-20.0))
        else:
            off = float(r.uniform(-200, 200))
            ps.append(P(mode=SAM, agc_mode=agc))
            iqs.append(synth.am(300 + s, T, mode=SAM, carrier_offset=off, depth=0.5,
                                level_step_block=T // 2, level_step_db=-20.0))
    return Case("c3_nfm_sam_agc", [(ps, T)], iqs, row_every=13)


def c4_zoom_rows(n=10, T=6):
    """C4: zoom index = receiver % 5, every block produces a spectrum + waterfall row."""
    ps = [P(spectrum_zoom=s % 5, current_scale=1 + (s // 5)) for s in range(n)]
    # after the I mirror and the +Fs/4 shift these tones sit at +1.5 kHz.. and -2.5 kHz: inside every zoom window
    iqs = [synth.two_tone(400 + s, T, f1=46500.0 - 300 * s, f2=50500.0) for s in range(n)]
    return Case("c4_zoom_rows", [(ps, T)], iqs, row_every=1)


PSK_TEXT = "CQ de T41 B200"


def c5_psk31(n=3):
    """C5: BPSK31 of a fixed string, tone tuned to DC by NCOFreq, +-100 Hz mask, AGC off."""
    ps, iqs = [], []
    T = None
    for s in range(n):
        nco = 1000 * (s + 1)
        iq, _bits = synth.psk31(500 + s, PSK_TEXT, nco_freq=nco, symbol_offset=5014, ebn0_db=20.0)
        T = iq.shape[0]
        ps.append(P(mode=USB, f_lo_cut=-100, f_hi_cut=100, agc_mode=0, psk31_enable=1, nco_freq=nco))
        iqs.append(iq)
    return Case("c5_psk31", [(ps, T)], iqs, psk=True)


def edge_silence_fullscale():
    """All-zero input (sign-of-zero handling) and a full-scale square-ish input (q15 rails)."""
    T = 6
    z = np.zeros((T, 2048, 2), np.float32)
    r = _rng(606)
    rails = np.where(r.random((T, 2048, 2)) < 0.5, -1.0, 32767.0 / 32768.0).astype(np.float32)
    ps = [P(mode=USB), P(mode=AM), P(mode=SAM), P(mode=NFM), P(mode=USB, agc_mode=0), P(mode=LSB)]
    iqs = [z, z, rails, z.copy(), rails.copy(), rails.copy()]
    return Case("edge_silence_fullscale", [(ps, T)], iqs, row_every=2)


def edge_param_changes():
    """Mode, cut-off, AGC, zoom, volume and NCO changes between calls; one block per call at
    the end; 5 receivers (ragged against the 4-receiver CTA)."""
    T1, T2, T3 = 7, 6, 1
    n = 5
    iqs = [synth.tone(700 + s, T1 + T2 + T3 + T3, 900.0 + 100 * s, nco_freq=0, amp=0.2) for s in range(n)]
    seg1 = [P(mode=USB), P(mode=LSB), P(mode=AM), P(mode=NFM, agc_mode=3), P(mode=SAM, agc_mode=4)]
    seg2 = [P(mode=LSB, nco_freq=2500), P(mode=USB, f_lo_cut=100, f_hi_cut=2400, agc_mode=2),
            P(mode=SAM, agc_mode=1, spectrum_zoom=3), P(mode=USB, agc_mode=1, audio_volume=55),
            P(mode=NFM, agc_mode=2, spectrum_zoom=0, rf_gain=4)]
    seg3 = [P(mode=LSB, nco_freq=-12345), P(mode=USB, f_lo_cut=100, f_hi_cut=2400, agc_mode=0),
            P(mode=AM, agc_mode=1, spectrum_zoom=4), P(mode=PSK31), P(mode=NFM, agc_mode=2, spectrum_zoom=2)]
    return Case("edge_param_changes", [(seg1, T1), (seg2, T2), (seg3, T3), (seg3, T3)], iqs, row_every=3)


def edge_rf_gain_ramp():
    """Codec_gain raises RFgain every 50 blocks (B13): cross two increments."""
    T = 104
    return Case("edge_rf_gain_ramp", [([P(mode=USB)], T)], [synth.tone(800, T, 1200.0, amp=0.02)], row_every=0)


def _eq(p, levels):
    p.receive_eq_flag = 1
    for i, v in enumerate(levels):
        p.equalizer_rec[i] = v
    return p


def c6_receive_eq():
    """Receive equaliser (Filter.cpp:117-165) in every mode, levels from flat to a deep notch,
    switched and re-levelled between the segments (the band states carry over, also while it is off)."""
    T1, T2, T3 = 7, 6, 5
    T = T1 + T2 + T3
    flat = [100] * 14
    tilt = [100, 80, 0, 120, 55, 100, 30, 90, 100, 10, 70, 100, 45, 100]
    bass = [0, 0, 0, 10, 20, 40, 60, 80, 100, 100, 100, 100, 100, 100]
    iqs = [synth.tone(900, T, 1000.0, amp=0.1), synth.tone(901, T, -1200.0, mode=LSB), synth.am(902, T),
           synth.nfm(903, T), synth.am(904, T, mode=SAM, carrier_offset=50.0), synth.tone(905, T, 500.0),
           synth.two_tone(906, T, 46500.0, 47900.0), synth.tone(907, T, 700.0)]
    seg1 = [_eq(P(mode=USB), tilt), _eq(P(mode=LSB), bass), _eq(P(mode=AM), tilt), _eq(P(mode=NFM, agc_mode=3), bass),
            _eq(P(mode=SAM, agc_mode=4), flat), _eq(P(mode=PSK31), tilt), _eq(P(mode=USB, agc_mode=0), flat), P(mode=USB)]
    seg2 = [P(mode=USB), _eq(P(mode=LSB), tilt), _eq(P(mode=AM), bass), _eq(P(mode=NFM, agc_mode=3), bass),
            P(mode=SAM, agc_mode=4), _eq(P(mode=PSK31), flat), _eq(P(mode=USB, agc_mode=0), bass), _eq(P(mode=USB), tilt)]
    seg3 = [_eq(P(mode=USB), bass), P(mode=LSB), _eq(P(mode=AM), bass), _eq(P(mode=USB, agc_mode=3), tilt),
            _eq(P(mode=SAM, agc_mode=4), tilt), P(mode=PSK31), _eq(P(mode=USB, agc_mode=0), bass), _eq(P(mode=USB), tilt)]
    return Case("c6_receive_eq", [(seg1, T1), (seg2, T2), (seg3, T3)], iqs, row_every=4)


def c7_lms_notch():
    """Automatic notch and LMS noise reduction (Xanr, Noise.cpp:322-369): each alone, both (two
    passes over one shared state), with the equaliser in front, switched between the segments."""
    T1, T2 = 8, 6
    T = T1 + T2
    tilt = [100, 80, 0, 120, 55, 100, 30, 90, 100, 10, 70, 100, 45, 100]
    iqs = [synth.two_tone(910, T, 46500.0, 47900.0), synth.tone(911, T, -1200.0, mode=LSB), synth.am(912, T),
           synth.nfm(913, T), synth.am(914, T, mode=SAM, carrier_offset=50.0), synth.tone(915, T, 500.0),
           synth.two_tone(916, T, 46800.0, 48100.0)]
    seg1 = [P(mode=USB, anr_notch_on=1), P(mode=LSB, nr_option=3), P(mode=AM, nr_option=3, anr_notch_on=1),
            P(mode=NFM, agc_mode=3, anr_notch_on=1), P(mode=SAM, agc_mode=4, anr_notch_on=1), P(mode=PSK31, anr_notch_on=1),
            _eq(P(mode=USB, nr_option=3, anr_notch_on=1), tilt)]
    seg2 = [P(mode=USB, nr_option=3), P(mode=LSB, anr_notch_on=1), P(mode=AM), P(mode=NFM, agc_mode=3, nr_option=3, anr_notch_on=1),
            P(mode=SAM, agc_mode=4), P(mode=PSK31, nr_option=3), P(mode=USB, anr_notch_on=1)]
    return Case("c7_lms_notch", [(seg1, T1), (seg2, T2)], iqs, row_every=4)


def c8_cw_filters():
    """The CW receive state's audio low-passes (Process.cpp:878-914): every filter index, off,
    outside the CW state, switched between the segments (each filter keeps its own state), behind equaliser + notch."""
    T1, T2 = 7, 6
    T = T1 + T2
    tilt = [100, 80, 0, 120, 55, 100, 30, 90, 100, 10, 70, 100, 45, 100]
    iqs = [synth.two_tone(920 + k, T, 46300.0 + 150.0 * k, 48200.0 - 90.0 * k) for k in range(8)]
    cw = lambda i, **kw: P(mode=kw.pop("mode", USB), cw_receive=1, cw_filter_index=i, **kw)
    seg1 = [cw(0), cw(1, mode=LSB), cw(2), cw(3, agc_mode=0), cw(4), cw(5), P(mode=USB, cw_filter_index=2),
            _eq(cw(1, anr_notch_on=1), tilt)]
    seg2 = [cw(3), cw(1, mode=LSB), P(mode=USB), cw(0, agc_mode=0), cw(5), cw(2), cw(2), _eq(cw(4, anr_notch_on=1), tilt)]
    return Case("c8_cw_filters", [(seg1, T1), (seg2, T2)], iqs, row_every=5)


def _with_clicks(iq, every=4, first=3, amp=0.5):
    """add a three-sample impulse to every `every`-th block (something for the noise blanker to find)"""
    iq = iq.copy()
    for t in range(first, iq.shape[0], every):
        iq[t, 700:703, 0] += amp
    return synth.to_q15_grid(iq)


def c9_kim_spectral_nb():
    """The 256-point spectral noise-reduction stages and the noise blanker (rest of SURVEY 8(f) rank 2): Kim1_NR
    (Noise.cpp:108-311, then x 30), SpectralNoiseReduction (Noise.cpp:379-655: 10 blocks of training that pass the audio
    through, then - the reference never sets its long-tone gain - signed zeros), NoiseBlanker (DSP_Fn.cpp:105-362) on
    clicks; every mode, narrow and wide filters (the stages' bin ranges follow the cut-offs), combined with the notch,
    the equaliser and each other, switched between the segments (Kim and spectral NR share their arrays)."""
    T1, T2 = 14, 8
    T = T1 + T2
    tilt = [100, 80, 0, 120, 55, 100, 30, 90, 100, 10, 70, 100, 45, 100]
    iqs = [synth.tone(930, T, 1000.0), synth.tone(931, T, -1200.0, mode=LSB), _with_clicks(synth.am(932, T)),
           synth.nfm(933, T), _with_clicks(synth.am(934, T, mode=SAM, carrier_offset=50.0)), synth.tone(935, T, 500.0),
           _with_clicks(synth.two_tone(936, T, 46800.0, 48100.0)), synth.tone(937, T, 700.0),
           _with_clicks(synth.tone(938, T, 900.0), every=1, first=0)]
    seg1 = [P(mode=USB, nr_option=1), P(mode=LSB, nr_option=2), P(mode=AM, nb_on=1), P(mode=NFM, agc_mode=3, nr_option=1),
            P(mode=SAM, agc_mode=4, nr_option=1, nb_on=1), P(mode=PSK31, nr_option=1, nb_on=1),
            _eq(P(mode=USB, nr_option=2, anr_notch_on=1, nb_on=1), tilt), P(mode=USB, f_lo_cut=-100, f_hi_cut=100, nr_option=1),
            P(mode=USB, nb_on=1, agc_mode=0)]
    seg2 = [P(mode=USB, nr_option=2), P(mode=LSB, nr_option=1), P(mode=AM, nr_option=1, nb_on=1), P(mode=NFM, agc_mode=3),
            P(mode=SAM, agc_mode=4, nr_option=2), P(mode=PSK31), P(mode=USB, nr_option=1, anr_notch_on=1),
            P(mode=USB, f_lo_cut=300, f_hi_cut=3000, nr_option=1), P(mode=USB, nb_on=1, nr_option=3)]
    return Case("c9_kim_spectral_nb", [(seg1, T1), (seg2, T2)], iqs, row_every=5)


ALL_CASES = [c1_single_usb, c1_single_usb_agc_off, c2_ssb_am_mix, c3_nfm_sam_agc, c4_zoom_rows, c5_psk31,
             edge_silence_fullscale, edge_param_changes, edge_rf_gain_ramp, c6_receive_eq, c7_lms_notch, c8_cw_filters,
             c9_kim_spectral_nb]


def run_case_on(case, make_stream):
    """Run a case through per-receiver CPU implementations (Tier-A or Tier-B).
    make_stream(params) -> object with set_params/process/debug.  Returns per-receiver dicts."""
    out = []
    for s in range(case.n_streams):
        st = make_stream(case.segments[0][0][s])
        parts = []
        b0 = 0
        for seg_i, (plist, n) in enumerate(case.segments):
            if seg_i > 0:
                st.set_params(plist[s])
            parts.append(st.process(case.iq[s][b0:b0 + n], case.row_every, case.psk))
            b0 += n
        res = dict(audio=np.concatenate([p["audio"] for p in parts]),
                   spec=np.concatenate([p["spec"] for p in parts]),
                   wf=np.concatenate([p["wf"] for p in parts]),
                   audio_ypixel=np.concatenate([p["audio_ypixel"] for p in parts]),
                   audio_max_sq_ave=np.concatenate([p["audio_max_sq_ave"] for p in parts]),
                   spec_frames=np.concatenate([p["spec_frames"] for p in parts]),
                   audio_frames=np.concatenate([p["audio_frames"] for p in parts]))
        if case.psk:
            res["psk_bits"] = np.concatenate([p["psk_bits"] for p in parts])
            res["psk_chars"] = np.concatenate([p["psk_chars"] for p in parts])
        res["debug"] = st.debug()
        out.append(res)
    return out


DEBUG_INT_FIELDS = ("agc_state", "agc_decay_type", "agc_hang_counter", "agc_action", "rf_gain", "codec_timer",
                    "zoom_sample_ptr", "first_block")
DEBUG_FLOAT_FIELDS = ("agc_volts", "agc_ring_max", "agc_save_volts", "agc_fast_backaverage",
                      "agc_hang_backaverage", "sam_phzerror", "sam_omega2", "sam_fil_out", "am_wold")
