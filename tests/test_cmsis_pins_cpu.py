"""CPU tier: the CMSIS-DSP restatement (oracle/cmsis_port.c) pinned against independent mathematics.

cmsis_port.c sits under BOTH oracle tiers (Tier-B links it, and so does the Tier-A build of the reference's own
translation units), so the golden vectors cannot catch an error in it.  CMSIS-DSP itself is not in the reference
tree (it ships prebuilt inside Teensyduino, no version pinned): these tests check every primitive the receive path
calls against scipy.signal / numpy evaluated in float64 from the published CMSIS semantics:

  arm_fir_decimate_f32      time-reversed tap convention, window ending at input m * M   Process.cpp:474-479
  arm_fir_interpolate_f32   polyphase up-sampler, no x L gain                            Process.cpp:917-919
  arm_fir_f32               plain FIR                                                    FFT.cpp (zoom), Filter.cpp
  arm_biquad_cascade_df1    coefficients {b0, b1, b2, a1, a2} with the a's pre-negated   Process.cpp:705
  arm_biquad_cascade_df2T   same coefficient convention, transposed form                 Process.cpp:127-128
  arm_sin_f32 / arm_cos_f32 513-entry table + linear interpolation (peak error ~1.9e-5)  Demod.cpp:73-74
  arm_q15_to_float / arm_float_to_q15, arm_cmplx_mult_cmplx_f32, arm_max_f32, arm_var_f32, arm_power_f32,
  arm_dot_prod_f32, arm_scale/add/mult/negate/copy

The tap-order checks use deliberately ASYMMETRIC taps (the reference's own low-pass design is asymmetric by one tap,
SURVEY.md B20), so a reversed convention fails by orders of magnitude, not by rounding.
"""
import ctypes as C

import numpy as np
import pytest
from scipy import signal

import oracle_py as O

F32P = C.POINTER(C.c_float)


class DecInst(C.Structure):
    _fields_ = [("M", C.c_uint8), ("numTaps", C.c_uint16), ("pCoeffs", F32P), ("pState", F32P)]


class IntInst(C.Structure):
    _fields_ = [("L", C.c_uint8), ("phaseLength", C.c_uint16), ("pCoeffs", F32P), ("pState", F32P)]


class FirInst(C.Structure):
    _fields_ = [("numTaps", C.c_uint16), ("pState", F32P), ("pCoeffs", F32P)]


class Df1Inst(C.Structure):
    _fields_ = [("numStages", C.c_uint32), ("pState", F32P), ("pCoeffs", F32P)]


class Df2TInst(C.Structure):
    _fields_ = [("numStages", C.c_uint8), ("pState", F32P), ("pCoeffs", F32P)]


def _p(a):
    return a.ctypes.data_as(F32P)


@pytest.fixture(scope="module")
def lib():
    lib = O.tier_b()
    lib.arm_sin_f32.restype = C.c_float
    lib.arm_sin_f32.argtypes = [C.c_float]
    lib.arm_cos_f32.restype = C.c_float
    lib.arm_cos_f32.argtypes = [C.c_float]
    return lib


def _asym_taps(n, seed):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal(n) * np.exp(-np.arange(n) / (n / 3.0))     # strongly asymmetric
    return h.astype(np.float32)


@pytest.mark.parametrize("M,ntaps,block", [(4, 28, 2048), (2, 46, 512), (2, 4, 64)])
def test_fir_decimate_tap_order_and_block_continuity(lib, M, ntaps, block):
    h = _asym_taps(ntaps, 1)
    rng = np.random.default_rng(2)
    x = rng.standard_normal(3 * block).astype(np.float32)
    state = np.zeros(ntaps + block - 1, np.float32)
    inst = DecInst()
    assert lib.arm_fir_decimate_init_f32(C.byref(inst), ntaps, M, _p(h), _p(state), block) == 0
    out = np.empty(3 * block // M, np.float32)
    for b in range(3):                                   # three calls: the history carried in pState is part of the pin
        lib.arm_fir_decimate_f32(C.byref(inst), _p(x[b * block:]), _p(out[b * block // M:]), block)
    # CMSIS: pCoeffs holds the taps time-reversed, {b[N-1] .. b[0]}, and the window of output m ENDS AT INPUT m*M (the
    # first of the M samples appended for it; the other M-1 serve the next output): y[m] = sum_i pCoeffs[i] * x[m*M - (N-1) + i]
    b = h[::-1].astype(np.float64)
    full = signal.lfilter(b, [1.0], x.astype(np.float64))
    want = full[0::M]
    assert np.abs(out - want).max() < 2e-5 * max(1.0, np.abs(want).max())
    # the opposite convention is far away
    wrong = signal.lfilter(h.astype(np.float64), [1.0], x.astype(np.float64))[0::M]
    assert np.abs(out - wrong).max() > 0.1
    assert np.abs(out - full[M - 1::M]).max() > 0.1      # nor does the window end at the LAST appended sample


def test_fir_decimate_init_rejects_ragged_block(lib):
    h = _asym_taps(28, 1)
    state = np.zeros(28 + 30, np.float32)
    inst = DecInst()
    assert lib.arm_fir_decimate_init_f32(C.byref(inst), 28, 4, _p(h), _p(state), 30) != 0   # ARM_MATH_LENGTH_ERROR


@pytest.mark.parametrize("L,ntaps,block", [(2, 48, 256), (4, 32, 512)])
def test_fir_interpolate_matches_upfirdn(lib, L, ntaps, block):
    h = _asym_taps(ntaps, 3)
    rng = np.random.default_rng(4)
    x = rng.standard_normal(3 * block).astype(np.float32)
    state = np.zeros(ntaps // L + block - 1, np.float32)
    inst = IntInst()
    assert lib.arm_fir_interpolate_init_f32(C.byref(inst), L, ntaps, _p(h), _p(state), block) == 0
    out = np.empty(3 * block * L, np.float32)
    for b in range(3):
        lib.arm_fir_interpolate_f32(C.byref(inst), _p(x[b * block:]), _p(out[b * block * L:]), block)
    # zero-stuff by L, filter with the taps time-reversed (CMSIS stores {b[N-1] .. b[0]}), no gain of L
    want = signal.upfirdn(h[::-1].astype(np.float64), x.astype(np.float64), up=L)[:out.size]
    assert np.abs(out - want).max() < 2e-5 * max(1.0, np.abs(want).max())
    wrong = signal.upfirdn(h.astype(np.float64), x.astype(np.float64), up=L)[:out.size]
    assert np.abs(out - wrong).max() > 0.1


def test_fir_matches_lfilter(lib):
    h = _asym_taps(4, 5)
    rng = np.random.default_rng(6)
    x = rng.standard_normal(3 * 64).astype(np.float32)
    state = np.zeros(4 + 64 - 1, np.float32)
    inst = FirInst()
    lib.arm_fir_init_f32(C.byref(inst), 4, _p(h), _p(state), 64)
    out = np.empty_like(x)
    for b in range(3):
        lib.arm_fir_f32(C.byref(inst), _p(x[b * 64:]), _p(out[b * 64:]), 64)
    want = signal.lfilter(h[::-1].astype(np.float64), [1.0], x.astype(np.float64))
    assert np.abs(out - want).max() < 2e-5


def _sos_as_cmsis(sos):
    """scipy sos rows [b0 b1 b2 1 a1 a2] -> CMSIS {b0, b1, b2, -a1, -a2} per stage"""
    return np.concatenate([[r[0], r[1], r[2], -r[4], -r[5]] for r in sos]).astype(np.float32)


@pytest.mark.parametrize("form", ["df1", "df2T"])
def test_biquad_cascades_match_sosfilt(lib, form):
    sos = signal.ellip(8, 0.5, 60, 0.2, output="sos")            # 4 stages, like the ZoomFFT and equaliser cascades
    coef = _sos_as_cmsis(sos)
    sos32 = np.array([[c[0], c[1], c[2], 1.0, -c[3], -c[4]] for c in coef.reshape(-1, 5)], np.float64)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(3 * 256).astype(np.float32)
    out = np.empty_like(x)
    if form == "df1":
        state = np.zeros(4 * 4, np.float32)
        inst = Df1Inst()
        lib.arm_biquad_cascade_df1_init_f32(C.byref(inst), 4, _p(coef), _p(state))
        for b in range(3):
            lib.arm_biquad_cascade_df1_f32(C.byref(inst), _p(x[b * 256:]), _p(out[b * 256:]), 256)
    else:
        state = np.zeros(2 * 4, np.float32)
        inst = Df2TInst()
        lib.arm_biquad_cascade_df2T_init_f32(C.byref(inst), 4, _p(coef), _p(state))
        for b in range(3):
            lib.arm_biquad_cascade_df2T_f32(C.byref(inst), _p(x[b * 256:]), _p(out[b * 256:]), 256)
    want = signal.sosfilt(sos32, x.astype(np.float64))
    assert np.abs(out - want).max() < 2e-5 * max(1.0, np.abs(want).max())


def test_sin_cos_table_interpolation(lib):
    xs = np.concatenate([np.linspace(-20.0, 20.0, 4001), [0.0, np.pi / 2, np.pi, 2 * np.pi, -np.pi / 2, 1e-8, -1e-8]])
    es, ec = 0.0, 0.0
    for x in xs.astype(np.float32):
        es = max(es, abs(lib.arm_sin_f32(float(x)) - np.sin(np.float64(x))))
        ec = max(ec, abs(lib.arm_cos_f32(float(x)) - np.cos(np.float64(x))))
    # linear interpolation in a 512-step table: error <= (2 pi / 512)^2 / 8 = 1.88e-5 (+ float rounding of the argument)
    assert es < 2.3e-5 and ec < 2.3e-5
    assert es > 5e-6            # it IS the table, not libm's sinf (the SAM PLL's parity depends on this, Appendix D)
    assert lib.arm_sin_f32(0.0) == 0.0 and lib.arm_cos_f32(0.0) == 1.0


def test_q15_conversions(lib):
    q = np.array([-32768, -32767, -1, 0, 1, 12345, 32767], np.int16)
    f = np.empty(q.size, np.float32)
    lib.arm_q15_to_float(q.ctypes.data_as(C.c_void_p), _p(f), q.size)
    assert np.array_equal(f, q.astype(np.float32) / np.float32(32768.0))
    x = np.array([0.0, 0.5, -0.5, 0.99999, 1.0, 1.5, -1.0, -1.5, 1e-6, -1e-6, 3e5, -3e5, np.inf, -np.inf, np.nan,
                  12345.7 / 32768.0, -12345.7 / 32768.0], np.float32)
    out = np.empty(x.size, np.int16)
    lib.arm_float_to_q15(_p(x), out.ctypes.data_as(C.c_void_p), x.size)
    # saturate(trunc(x * 32768)): truncation toward zero, no rounding (ARM_MATH_ROUNDING is not defined); NaN -> 0
    want = np.array([0, 16384, -16384, 32767, 32767, 32767, -32768, -32768, 0, 0, 32767, -32768, 32767, -32768, 0,
                     12345, -12345], np.int16)
    assert np.array_equal(out, want)


def test_elementwise_and_reductions(lib):
    rng = np.random.default_rng(8)
    a = rng.standard_normal(256).astype(np.float32)
    b = rng.standard_normal(256).astype(np.float32)
    o = np.empty(256, np.float32)
    lib.arm_scale_f32.argtypes = [F32P, C.c_float, F32P, C.c_uint32]
    lib.arm_scale_f32(_p(a), 1.5, _p(o), 256)
    assert np.array_equal(o, a * np.float32(1.5))
    lib.arm_add_f32(_p(a), _p(b), _p(o), 256)
    assert np.array_equal(o, a + b)
    lib.arm_mult_f32(_p(a), _p(b), _p(o), 256)
    assert np.array_equal(o, a * b)
    lib.arm_negate_f32(_p(a), _p(o), 256)
    assert np.array_equal(o, -a)
    lib.arm_copy_f32(_p(a), _p(o), 256)
    assert np.array_equal(o, a)
    oc = np.empty(256, np.float32)
    lib.arm_cmplx_mult_cmplx_f32(_p(a), _p(b), _p(oc), 128)
    want = (a[0::2].astype(np.float64) + 1j * a[1::2]) * (b[0::2].astype(np.float64) + 1j * b[1::2])
    assert np.abs(oc[0::2] - want.real).max() < 1e-6 and np.abs(oc[1::2] - want.imag).max() < 1e-6
    r = C.c_float()
    lib.arm_dot_prod_f32(_p(a), _p(b), 256, C.byref(r))
    assert abs(r.value - float(np.dot(a.astype(np.float64), b.astype(np.float64)))) < 1e-4
    lib.arm_power_f32(_p(a), 256, C.byref(r))
    assert abs(r.value - float(np.sum(a.astype(np.float64) ** 2))) < 1e-3
    lib.arm_var_f32(_p(a), 256, C.byref(r))
    assert abs(r.value - float(np.var(a.astype(np.float64), ddof=1))) < 1e-5
    idx = C.c_uint32()
    lib.arm_max_f32(_p(a), 256, C.byref(r), C.byref(idx))
    assert r.value == a.max() and idx.value == int(np.argmax(a))


def test_cfft256_matches_numpy(lib):
    """the 256-point instance the noise-reduction stages use (Noise.cpp:206,279,454,636); the 512-point one is pinned in
    test_oracle_cpu.py"""
    class CfftInst(C.Structure):
        _fields_ = [("fftLen", C.c_uint16), ("pTwiddle", F32P), ("pBitRevTable", C.POINTER(C.c_uint16)),
                    ("bitRevLength", C.c_uint16)]
    inst = CfftInst.in_dll(lib, "arm_cfft_sR_f32_len256")
    assert inst.fftLen == 256
    rng = np.random.default_rng(12)
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    buf = np.empty(512, np.float32)
    buf[0::2], buf[1::2] = x.real, x.imag
    orig = buf.copy()
    lib.arm_cfft_f32(C.byref(inst), _p(buf), 0, 1)
    want = np.fft.fft(orig[0::2].astype(np.float64) + 1j * orig[1::2].astype(np.float64))
    assert np.abs((buf[0::2] + 1j * buf[1::2]) - want).max() / np.abs(want).max() < 1e-6
    lib.arm_cfft_f32(C.byref(inst), _p(buf), 1, 1)
    assert np.abs(buf - orig).max() < 2e-6
