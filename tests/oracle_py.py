"""ctypes access to the TEST-ONLY CPU oracles (tests, smoke() and bench.py's CPU legs only).

Tier-B: oracle/libt41oracle.so   our portable restatement, one handle per receiver.
Tier-A: oracle/_ref/libt41ref.so the reference's own translation units compiled in place
        (only buildable where /root/reference is mounted; the built .so ships to the GPU box).
        The reference keeps DSP state in function statics, so every RefStream loads a
        private copy of the library.
"""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
TIER_B_PATH = os.path.join(ORACLE_DIR, "libt41oracle.so")
TIER_A_PATH = os.path.join(ORACLE_DIR, "_ref", "libt41ref.so")

DEMOD_USB, DEMOD_LSB, DEMOD_AM, DEMOD_NFM, DEMOD_PSK31, DEMOD_SAM = 0, 1, 2, 3, 5, 8
BLOCK = 2048
SPEC_FRAME_BYTES = 518    # specData[], t41Control.cpp:23
AUDIO_SPEC_PIXELS = 270   # AUDIO_SPEC_BOX_W - 2 (Display.h:45, Process.cpp:555)


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "mode", "f_lo_cut", "f_hi_cut", "nco_freq", "agc_mode", "agc_thresh", "audio_volume",
        "rf_gain_all_bands", "rf_gain", "spectrum_zoom", "current_scale", "pixel_offset",
        "current_nf", "spectrum_noise_floor", "nfm_filter_bw", "psk31_enable")] + [
        ("iq_amp_correction", C.c_float), ("iq_phase_correction", C.c_float),
        ("receive_eq_flag", C.c_int32), ("equalizer_rec", C.c_int32 * 14),
        ("nr_option", C.c_int32), ("anr_notch_on", C.c_int32), ("cw_receive", C.c_int32),
        ("cw_filter_index", C.c_int32), ("nb_on", C.c_int32)]

    def copy(self):
        p = Params()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(Params))
        return p


class Debug(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "agc_state", "agc_decay_type", "agc_hang_counter", "agc_action", "rf_gain", "codec_timer",
        "zoom_sample_ptr", "first_block")] + [(n, C.c_float) for n in (
        "agc_volts", "agc_ring_max", "agc_save_volts", "agc_fast_backaverage", "agc_hang_backaverage",
        "sam_phzerror", "sam_omega2", "sam_fil_out")] + [
        ("dc_state", C.c_float * 2), ("am_wold", C.c_float),
        ("osc_vect_q", C.c_double), ("osc_vect_i", C.c_double)]


class Tables(C.Structure):
    _fields_ = [("dec1", C.c_float * 28), ("dec2", C.c_float * 46), ("int1", C.c_float * 48),
                ("int2", C.c_float * 32), ("mask", C.c_float * 1024), ("am_lp", C.c_float * 5),
                ("zoom_fir", C.c_float * 4), ("agc", C.c_float * 16),
                ("attack_buffsize", C.c_int32), ("hang_counter_load", C.c_int32)]

    def as_dict(self):
        d = {n: np.array(getattr(self, n), dtype=np.float32) for n in
             ("dec1", "dec2", "int1", "int2", "mask", "am_lp", "zoom_fir", "agc")}
        d["attack_buffsize"] = int(self.attack_buffsize)
        d["hang_counter_load"] = int(self.hang_counter_load)
        return d


def build_oracle(ref=False):
    """Compile the oracle libraries (used by build() and by the test session)."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])
    if ref and os.path.isdir("/root/reference/software/T41_SDR"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


_tier_b = None


def tier_b():
    global _tier_b
    if _tier_b is None:
        if not os.path.exists(TIER_B_PATH):
            build_oracle()
        lib = C.CDLL(TIER_B_PATH)
        lib.t41o_create.restype = C.c_void_p
        lib.t41o_destroy.argtypes = [C.c_void_p]
        lib.t41o_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        lib.t41o_get_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        lib.t41o_get_tables.argtypes = [C.c_void_p, C.POINTER(Tables)]
        lib.t41o_get_debug.argtypes = [C.c_void_p, C.POINTER(Debug)]
        lib.t41o_process.argtypes = [C.c_void_p] + [C.c_void_p] * 2 + [C.c_int, C.c_int] + [C.c_void_p] * 4
        lib.t41o_default_params.argtypes = [C.POINTER(Params)]
        lib.t41o_capture_audio_spectrum.argtypes = [C.c_void_p] * 3
        lib.t41o_capture_audio_spectrum.restype = None
        lib.t41o_capture_control_frames.argtypes = [C.c_void_p] * 3
        lib.t41o_capture_control_frames.restype = None
        lib.t41o_smeter_dbm.restype = C.c_float
        lib.t41o_smeter_dbm.argtypes = [C.c_float, C.c_float, C.c_int32, C.c_int32]
        lib.t41o_smeter_bar.argtypes = [C.c_float]
        lib.t41o_smeter_bar.restype = C.c_int32
        lib.t41o_log10f_fast.restype = C.c_float
        lib.t41o_log10f_fast.argtypes = [C.c_float]
        lib.t41o_approx_atan2.restype = C.c_float
        lib.t41o_approx_atan2.argtypes = [C.c_float, C.c_float]
        _tier_b = lib
    return _tier_b


def default_params():
    p = Params()
    tier_b().t41o_default_params(C.byref(p))
    return p


def mode_default_cuts(mode):
    lo, hi = C.c_int32(), C.c_int32()
    tier_b().t41o_mode_default_cuts(C.c_int32(mode), C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _StreamBase:
    def process(self, iq, row_every=0, want_psk=False):
        """iq: float32 [n_blocks, 2048, 2].  Returns dict(audio, spec, wf, psk_bits, psk_chars, audio_ypixel,
        audio_max_sq_ave)."""
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        n = iq.shape[0]
        assert iq.shape == (n, BLOCK, 2)
        audio = np.empty((n, BLOCK), np.float32)
        n_rows = 0 if row_every <= 0 else (n + row_every - 1) // row_every
        spec = np.zeros((n_rows, 512), np.int16)
        wf = np.zeros((n_rows, 512), np.uint16)
        bits = np.full(n, -1, np.int8) if want_psk else None
        chars = np.zeros(n, np.uint8) if want_psk else None
        ypix = np.zeros((n_rows, AUDIO_SPEC_PIXELS), np.int32)
        mxave = np.zeros(n_rows, np.float32)
        frames = np.zeros((n_rows, SPEC_FRAME_BYTES), np.uint8)
        aframes = np.zeros((n_rows, AUDIO_SPEC_PIXELS), np.uint8)
        self._capture(_ptr(ypix) if n_rows else None, _ptr(mxave) if n_rows else None)
        self._capture_frames(_ptr(frames) if n_rows else None, _ptr(aframes) if n_rows else None)
        try:
            rc = self._process(_ptr(iq), _ptr(audio), n, row_every, _ptr(spec) if n_rows else None,
                               _ptr(wf) if n_rows else None, _ptr(bits), _ptr(chars))
        finally:
            self._capture(None, None)
            self._capture_frames(None, None)
        if rc < 0:
            raise RuntimeError("oracle process failed rc=%d" % rc)
        assert rc == n_rows
        return dict(audio=audio, spec=spec, wf=wf, psk_bits=bits, psk_chars=chars, audio_ypixel=ypix,
                    audio_max_sq_ave=mxave, spec_frames=frames, audio_frames=aframes)


class OracleStream(_StreamBase):
    """One Tier-B receiver."""

    def __init__(self, params=None):
        self.lib = tier_b()
        self.h = self.lib.t41o_create()
        if params is not None:
            self.set_params(params)

    def close(self):
        if self.h:
            self.lib.t41o_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, p):
        rc = self.lib.t41o_set_params(self.h, C.byref(p))
        if rc:
            raise ValueError("t41o_set_params rc=%d" % rc)

    def _process(self, *a):
        return self.lib.t41o_process(self.h, *a)

    def _capture(self, ypix, mx):
        self.lib.t41o_capture_audio_spectrum(self.h, ypix, mx)

    def _capture_frames(self, fr, afr):
        self.lib.t41o_capture_control_frames(self.h, fr, afr)

    def tables(self):
        t = Tables()
        self.lib.t41o_get_tables(self.h, C.byref(t))
        return t.as_dict()

    def debug(self):
        d = Debug()
        self.lib.t41o_get_debug(self.h, C.byref(d))
        return d


def tier_a_available():
    return os.path.exists(TIER_A_PATH)


class RefStream(_StreamBase):
    """One Tier-A receiver = one private copy of the reference-built library."""

    def __init__(self, params=None):
        if not tier_a_available():
            raise RuntimeError("oracle/_ref/libt41ref.so not built")
        fd, self._tmp = tempfile.mkstemp(suffix=".so", prefix="t41ref_")
        os.close(fd)
        shutil.copyfile(TIER_A_PATH, self._tmp)
        self.lib = C.CDLL(self._tmp)
        os.unlink(self._tmp)  # mapping stays valid
        lib = self.lib
        lib.t41ref_set_params.argtypes = [C.POINTER(Params)]
        lib.t41ref_get_params.argtypes = [C.POINTER(Params)]
        lib.t41ref_get_tables.argtypes = [C.POINTER(Tables)]
        lib.t41ref_get_debug.argtypes = [C.POINTER(Debug)]
        lib.t41ref_process.argtypes = [C.c_void_p] * 2 + [C.c_int, C.c_int] + [C.c_void_p] * 4
        lib.t41ref_capture_audio_spectrum.argtypes = [C.c_void_p] * 2
        lib.t41ref_capture_audio_spectrum.restype = None
        lib.t41ref_capture_control_frames.argtypes = [C.c_void_p] * 2
        lib.t41ref_capture_control_frames.restype = None
        lib.t41ref_load_wav.argtypes = [C.c_char_p, C.c_uint32]
        lib.t41ref_read_wave.argtypes = [C.c_void_p, C.c_int]
        lib.t41ref_log10f_fast.restype = C.c_float
        lib.t41ref_log10f_fast.argtypes = [C.c_float]
        lib.t41ref_approx_atan2.restype = C.c_float
        lib.t41ref_approx_atan2.argtypes = [C.c_float, C.c_float]
        rc = lib.t41ref_init()
        if rc:
            raise RuntimeError("t41ref_init rc=%d" % rc)
        if params is not None:
            self.set_params(params)

    def set_params(self, p):
        rc = self.lib.t41ref_set_params(C.byref(p))
        if rc:
            raise ValueError("t41ref_set_params rc=%d" % rc)

    def _process(self, *a):
        return self.lib.t41ref_process(*a)

    def _capture(self, ypix, mx):
        self.lib.t41ref_capture_audio_spectrum(ypix, mx)

    def _capture_frames(self, fr, afr):
        self.lib.t41ref_capture_control_frames(fr, afr)

    # the reference's WAV reader (Utility.cpp:773-888) on a host file
    def load_wav(self, path, num_samples):
        return self.lib.t41ref_load_wav(os.fsencode(path), int(num_samples))

    def read_wave(self, size_buf):
        buf = np.empty(size_buf, np.float32)
        return buf if self.lib.t41ref_read_wave(_ptr(buf), size_buf) else None

    def tables(self):
        t = Tables()
        self.lib.t41ref_get_tables(C.byref(t))
        return t.as_dict()

    def debug(self):
        d = Debug()
        self.lib.t41ref_get_debug(C.byref(d))
        return d


def snr_db(ref, got):
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    err = np.sum((ref - got) ** 2)
    sig = np.sum(ref ** 2)
    if err == 0:
        return float("inf")
    if sig == 0:
        return float("-inf")
    return 10.0 * np.log10(sig / err)
