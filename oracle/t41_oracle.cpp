/*
 * oracle/t41_oracle.cpp — TEST INFRASTRUCTURE ("Tier-B oracle"), not product code.
 *
 * Portable restatement of the T41 receive chain.  Every function cites the
 * reference lines it follows (T41/ = /root/reference/software/T41_SDR/).  The
 * firmware's single receiver built from globals becomes one `t41o_stream` object.
 * Build: g++ -O2 -ffp-contract=off (see oracle/Makefile).  Mixed float/double
 * expressions keep the promotions of the reference source (Teensy 4.1 has a
 * double-precision FPU, so `double` there is real binary64).
 */
#include "t41_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "cmsis_port.h"
#include "t41_tables_data.h"

namespace {

/* float constants of T41/FIR.h:6-16 (they shadow Arduino's double PI in the DSP files) */
const float kPi = 3.1415926535897932384626433832795f;
const float kHalfPi = 1.5707963267948966192313216916398f;
const float kTwoPi = 6.283185307179586476925286766559f;
const float kFourPi = 2.0f * kTwoPi;
const float kSixPi = 3.0f * kTwoPi;

const int kSampleRate = 192000;      /* T41/T41_SDR.ino:129 */
const float kDF1 = 4.0f;             /* T41/T41_SDR.ino:333-335 */
const float kDF = 8.0f;
const float kNAtt = 90.0f;           /* T41/T41_SDR.ino:336 */
const int kBlock = 2048;             /* BUFFER_SIZE * N_BLOCKS */
const int kFFT = 512;                /* FFT_LENGTH, T41/SDT.h:39 */
const int kDec = 256;                /* samples per block at 24 kS/s */
const int kDec1Taps = 28;            /* T41/T41_SDR.ino:344 (probed) */
const int kDec2Taps = 46;            /* T41/T41_SDR.ino:345 (probed) */
const int kMaskTaps = 257;           /* T41/Filter.cpp:18 */
const int kRing = 1921;              /* RB_SIZE, T41/DSP_Fn.cpp:470 */
const int kSpecRes = 512;            /* SPECTRUM_RES */
const int kSpectrumTopY = 100;       /* T41/Display.h:16,21 */
const int kSpectrumBottom = 249;     /* T41/Display.h:22 */
const double kFmDemodK = 0.340447550238101026565118445432744920253753662109375; /* T41/Demod.h:7 */

/* ------------------------------------------------------------------ */
/* math helpers, T41/Utility.cpp                                       */
/* ------------------------------------------------------------------ */

/* T41/Utility.cpp:245-258 */
float Log10Fast(float X) {
  int E;
  float F = frexpf(fabsf(X), &E);
  float Y = 1.23149591368684f;
  Y *= F;
  Y += -4.11852516267426f;
  Y *= F;
  Y += 6.02197014179219f;
  Y *= F;
  Y += -3.13396450166353f;
  Y += E;
  return Y * 0.3010299956639812f;
}

/* T41/Utility.cpp:269-285 */
float AlphaBetaMagnitude(float inphase, float quadrature) {
  const float alpha = 0.960433870103;
  const float beta = 0.397824734759;
  float ai = (float)fabs(inphase);
  float aq = (float)fabs(quadrature);
  if (ai > aq) return alpha * ai + beta * aq;
  return alpha * aq + beta * ai;
}

/* T41/Utility.cpp:298-302 */
float AtanPoly(float z) {
  const float n1 = 0.97239411f;
  const float n2 = -0.19194795f;
  return (n1 + n2 * z * z) * z;
}

/* T41/Demod.cpp:148-197, including the +-TPI (not pi/2) quirk of the |y|>=|x| branch */
float Atan2Approx(float y, float x) {
  if (x != 0.0f) {
    if (fabsf(x) > fabsf(y)) {
      const float z = y / x;
      if (x > 0.0f) return AtanPoly(z);
      if (y >= 0.0f) return AtanPoly(z) + kPi;
      return AtanPoly(z) - kPi;
    }
    const float z = x / y;
    if (y > 0.0f) return -AtanPoly(z) + kTwoPi;
    return -AtanPoly(z) - kTwoPi;
  }
  if (y > 0.0f) return kTwoPi;
  if (y < 0.0f) return -kTwoPi;
  return 0.0f;
}

/* T41/Utility.cpp:197-203 */
float SincM(int m, float fc) {
  float x = m * kHalfPi;
  if (m == 0) return 1.0f;
  return sinf(x * fc) / (fc * x);
}

/* T41/Utility.cpp:213-230 — modified Bessel I0 by series */
float BesselI0(float x) {
  float x2 = x / 2.0;
  float summe = 1.0;
  float ds = 1.0;
  float di = 1.0;
  float errorlimit = 1e-9;
  float tmp;
  do {
    tmp = x2 / di;
    tmp *= tmp;
    ds *= tmp;
    summe += ds;
    di += 1.0;
  } while (ds >= errorlimit * summe);
  return summe;
}

/* ------------------------------------------------------------------ */
/* filter design, T41/FIR.cpp                                          */
/* ------------------------------------------------------------------ */

/* T41/FIR.cpp:908-980, Kaiser-windowed sinc; only type 0 (low-pass) is on the RX path */
void DesignKaiserLowpass(float *coeffs, int numCoeffs, float fc, float Astop, float Fsamprate) {
  int nc = numCoeffs;
  float Beta;
  float izb;
  float fcf;
  float x, w;
  fc = fc / Fsamprate;
  if (Astop < 20.96) {
    Beta = 0.0;
  } else if (Astop >= 50.0) {
    Beta = 0.1102 * (Astop - 8.71);
  } else {
    Beta = 0.5842 * powf((Astop - 20.96), 0.4) + 0.07886 * (Astop - 20.96);
  }
  izb = BesselI0(Beta);
  fcf = fc * 2.0;
  /* nc taps for ii = -nc, -nc+2, ..., nc-2: centre at index nc/2 (B20) */
  int jj = 0;
  for (int ii = -nc; ii < nc; ii += 2, jj++) {
    x = (float)ii / (float)nc;
    w = BesselI0(Beta * sqrtf(1.0f - x * x)) / izb;
    coeffs[jj] = fcf * SincM(ii, fcf) * w;
  }
}

/* T41/FIR.cpp:1008-1065, window 1 (4-term Blackman-Harris, FIR_filter_window = 1, FIR.cpp:10) */
void DesignComplexBandpass(float *coeffs_I, float *coeffs_Q, int numCoeffs, float FLoCut, float FHiCut,
                           float SampleRate) {
  float nFL = FLoCut / SampleRate;
  float nFH = FHiCut / SampleRate;
  float nFc = (nFH - nFL) / 2.0;
  float nFs = kPi * (nFH + nFL);
  float fCenter = 0.5 * (float)(numCoeffs - 1);
  float x;
  float z;
  for (int i = 0; i < numCoeffs; i++) {
    x = (float)i - fCenter;
    float d = (float)i - fCenter;
    float ad = d > 0 ? d : -d; /* Arduino abs() macro */
    if (ad < 0.01) {
      z = 2.0 * nFc;
    } else {
      z = (float)sinf(kTwoPi * x * nFc) / (kPi * x) *
          (0.35875 - 0.48829 * cosf((kTwoPi * i) / (numCoeffs - 1)) +
           0.14128 * cosf((kFourPi * i) / (numCoeffs - 1)) -
           0.01168 * cosf((kSixPi * i) / (numCoeffs - 1)));
    }
    coeffs_I[i] = z * cosf(nFs * x);
    coeffs_Q[i] = z * sinf(nFs * x);
  }
}

/* T41/FIR.cpp:1076-1106, low-pass branch */
void DesignBiquadLowpass(float *set, float f0, float Q, float sample_rate) {
  if (f0 > sample_rate / 2.0) f0 = sample_rate / 2.0;
  float w0 = f0 * (kTwoPi / sample_rate);
  float sinW0 = sinf(w0);
  float alpha = sinW0 / (Q * 2.0);
  float cosW0 = cosf(w0);
  float scale = 1.0 / (1.0 + alpha);
  set[0] = ((1.0 - cosW0) / 2.0) * scale;
  set[1] = (1.0 - cosW0) * scale;
  set[2] = set[0];
  set[3] = (2.0 * cosW0) * scale;
  set[4] = (-1.0 + alpha) * scale;
}

/* ------------------------------------------------------------------ */
/* AGC parameters, T41/DSP_Fn.cpp:368-468                              */
/* ------------------------------------------------------------------ */
struct AgcSetup {
  /* AGCPrep() values */
  float tau_attack, tau_decay, max_input, out_targ, var_gain, tau_fast_backaverage, tau_fast_decay;
  float pop_ratio, tau_hang_backmult, hangtime, hang_thresh, tau_hang_decay, fixed_gain;
  int n_tau;
  int hang_enable;
  /* AGCLoadValues() results */
  float max_gain, attack_mult, decay_mult, fast_decay_mult, fast_backmult, onemfast_backmult;
  float out_target, min_volts, slope_constant, inv_max_input, hang_level, hang_backmult;
  float onemhang_backmult, hang_decay_mult;
  int attack_buffsize;
};

/* T41/DSP_Fn.cpp:444-468 (without the trailing AGCLoadValues call) */
void AgcPrepDefaults(AgcSetup &a) {
  a.tau_attack = 0.001;
  a.tau_decay = 0.250;
  a.n_tau = 4;
  a.max_gain = 10000.0;
  a.fixed_gain = 20.0;
  a.max_input = 1.0;
  a.out_targ = 1.0;
  a.var_gain = 1.5;
  a.tau_fast_backaverage = 0.250;
  a.tau_fast_decay = 0.005;
  a.pop_ratio = 5.0;
  a.hang_enable = 1;
  a.tau_hang_backmult = 0.500;
  a.hangtime = 0.250;
  a.hang_thresh = 0.250;
  a.tau_hang_decay = 0.100;
}

/* T41/DSP_Fn.cpp:368-435; hang_thresh is sticky across mode changes (B12) */
void AgcLoad(AgcSetup &a, int AGCMode, int AGC_thresh) {
  float tmp;
  float sample_rate = (float)kSampleRate / kDF;
  switch (AGCMode) {
    case 1: a.hangtime = 2.000; a.tau_decay = 2.000; break;
    case 2: a.hangtime = 1.000; a.tau_decay = 0.5; break;
    case 3: a.hang_thresh = 1.0; a.hangtime = 0.000; a.tau_decay = 0.250; break;
    case 4: a.hang_thresh = 1.0; a.hangtime = 0.0; a.tau_decay = 0.050; break;
    default: break;
  }
  a.max_gain = powf(10.0, (float)AGC_thresh / 20.0);
  a.attack_buffsize = (int)ceil(sample_rate * a.n_tau * a.tau_attack);
  a.attack_mult = 1.0 - expf(-1.0 / (sample_rate * a.tau_attack));
  a.decay_mult = 1.0 - expf(-1.0 / (sample_rate * a.tau_decay));
  a.fast_decay_mult = 1.0 - expf(-1.0 / (sample_rate * a.tau_fast_decay));
  a.fast_backmult = 1.0 - expf(-1.0 / (sample_rate * a.tau_fast_backaverage));
  a.onemfast_backmult = 1.0 - a.fast_backmult;
  a.out_target = a.out_targ * (1.0 - expf(-(float)a.n_tau)) * 0.9999;
  a.min_volts = a.out_target / (a.var_gain * a.max_gain);
  tmp = log10f(a.out_target / (a.max_input * a.var_gain * a.max_gain));
  if (tmp == 0.0) tmp = 1e-16;
  a.slope_constant = (a.out_target * (1.0 - 1.0 / a.var_gain)) / tmp;
  a.inv_max_input = 1.0 / a.max_input;
  tmp = powf(10.0, (a.hang_thresh - 1.0) / 0.125);
  a.hang_level = (a.max_input * tmp + (a.out_target / (a.var_gain * a.max_gain)) * (1.0 - tmp)) * 0.637;
  a.hang_backmult = 1.0 - expf(-1.0 / (sample_rate * a.tau_hang_backmult));
  a.onemhang_backmult = 1.0 - a.hang_backmult;
  a.hang_decay_mult = 1.0 - expf(-1.0 / (sample_rate * a.tau_hang_decay));
}

}  // namespace

/* ------------------------------------------------------------------ */
/* one receiver                                                        */
/* ------------------------------------------------------------------ */
struct t41o_stream {
  t41o_params prm;
  int last_set_rf_gain;

  /* tables */
  float dec1_coeffs[kDec1Taps];
  float dec2_coeffs[kDec2Taps];
  float int1_coeffs[48];
  float int2_coeffs[32];
  float mask[2 * kFFT];
  float am_lp_coeffs[5];
  float zoom_fir_coeffs[4];
  const float *zoom_iir_coeffs;
  AgcSetup agc;

  /* working buffers (names follow the firmware's) */
  float bufL[kBlock], bufR[kBlock];
  float bufL_ex[kBlock], bufR_ex[kBlock];
  float fft_buf[2 * kFFT];
  float ifft_buf[2 * kFFT + 1];

  /* persistent DSP state */
  float dc_state[2];                                     /* T41/Process.cpp:42 */
  arm_biquad_cascade_df2T_instance_f32 dc_block;
  double osc_vect_q, osc_vect_i;                         /* T41/Freq_Shift.cpp:13-14 */
  float dec1_state_i[kDec1Taps + kBlock - 1], dec1_state_q[kDec1Taps + kBlock - 1];
  float dec2_state_i[kDec2Taps + kBlock / 4 - 1], dec2_state_q[kDec2Taps + kBlock / 4 - 1];
  arm_fir_decimate_instance_f32 dec1_i, dec1_q, dec2_i, dec2_q;
  float last_sample_l[kDec], last_sample_r[kDec];        /* T41/T41_SDR.ino:403-404 */
  int first_block;                                       /* T41/Process.cpp:47 */
  /* AGC, T41/DSP_Fn.cpp:28-64,482-492 */
  float agc_ring[2 * kRing];
  float agc_abs_ring[kRing];
  int agc_out_index;
  unsigned agc_in_index;
  int agc_hang_counter;
  int agc_state, agc_decay_type, agc_action;
  float agc_fast_backaverage, agc_hang_backaverage, agc_ring_max, agc_save_volts, agc_volts;
  /* AM, T41/Process.cpp:73, T41_SDR.ino:373 */
  float am_wold;
  float am_lp_state[4];
  arm_biquad_casd_df1_inst_f32 am_lp;
  /* SAM, T41/Demod.cpp:13-23 */
  float sam_omega_min, sam_omega_max, sam_g1, sam_g2;
  float sam_phzerror, sam_fil_out, sam_del_out, sam_omega2;
  /* NFM, T41/Demod.cpp:221-222 */
  float nfm_last_i, nfm_last_q;
  /* interpolators, T41/T41_SDR.ino:384-394 */
  float int1_state[24 + kDec - 1];
  float int2_state[8 + 2 * kDec - 1];
  arm_fir_interpolate_instance_f32 int1, int2;
  /* Codec_gain, T41/Process.cpp:980 */
  unsigned codec_timer;
  int rf_gain;
  /* display spectrum, T41/FFT.cpp:13-18,72-73 */
  float zoom_iir_state_i[16], zoom_iir_state_q[16];
  arm_biquad_casd_df1_inst_f32 zoom_iir_i, zoom_iir_q;
  float zoom_fir_state_i[4 + kBlock - 1], zoom_fir_state_q[4 + kBlock - 1];
  arm_fir_decimate_instance_f32 zoom_fir_i, zoom_fir_q;
  float zoom_ring_x[kSpecRes], zoom_ring_y[kSpecRes];
  int zoom_sample_ptr;
  /* variable-leak LMS of the automatic notch / LMS noise reduction, T41/Noise.cpp:33-53 */
  float anr_d[512], anr_w[512];
  int anr_in_idx;
  float anr_lidx, anr_ngamma;
  /* spectral noise reduction stages, T41/Noise.cpp:17-37: Kim1_NR and SpectralNoiseReduction work on the SAME arrays
     (they are globals there); zero at start like every static of the host build */
  float nr_last_sample[128], nr_last_ifft[128];
  float nr_X[128][3], nr_E[128][15], nr_M[128], nr_Nest[128][2], nr_lambda[128], nr_Gts[128][2], nr_G[128];
  float nr_SNR_prio[128], nr_SNR_post[128], nr_Hk_old[128], nr_long_tone_gain[128];
  uint32_t nr_X_pointer, nr_E_pointer;               /* Kim1_NR's statics, Noise.cpp:109-110 */
  uint8_t nr_init_counter;                           /* SpectralNoiseReduction's statics, Noise.cpp:389-427 */
  int nr_first_time_2;
  float nr_pslp[128], nr_xt[128];
  int nr_consts_ready;
  float nr_xih1r, nr_pfac;
  /* noise blanker, T41/DSP_Fn.cpp:143 */
  float nb_last_frame_end[80];
  /* CW audio low-passes, T41/CWProcessing.cpp:38-49 */
  float cw_state[5][12];
  arm_biquad_cascade_df2T_instance_f32 cw[5];
  /* receive equaliser, T41/Filter.cpp:43-72 */
  float eq_state[14][8];
  arm_biquad_cascade_df2T_instance_f32 eq[14];
  /* audio-spectrum by-product, T41/Process.cpp:32-34 */
  float audio_max_sq_ave;
  int audio_ypixel[T41O_AUDIO_SPEC_PIXELS];
  int32_t *cap_ypixel;
  float *cap_max_ave;
  /* spectrum serial frame for the PC control app, T41/t41Control.cpp:21-23 */
  uint8_t spec_data[T41O_SPEC_FRAME_BYTES];
  int control_data_flag;
  uint8_t sent_spec[T41O_SPEC_FRAME_BYTES], sent_audio[T41O_AUDIO_SPEC_PIXELS];   /* serial-port writes of this block */
  uint8_t *cap_frames, *cap_audio_frames;
  float spec_buf[2 * kSpecRes];
  float fft_spec[kSpecRes];
  float fft_spec_old[kSpecRes];
  int16_t pixelnew[kSpecRes];
  uint16_t waterfall[kSpecRes];
  /* PSK31 tap */
  t41o_psk31 psk;
  unsigned psk_block_count;
};

namespace {

/* T41/Filter.cpp:260-284: taps 0..256 interleaved, zero from float index 513 (B7), FFT */
void BuildFilterMask(t41o_stream *s, const float *coef_i, const float *coef_q) {
  for (int i = 0; i < kMaskTaps; i++) {
    s->mask[i * 2] = coef_i[i];
    s->mask[i * 2 + 1] = coef_q[i];
  }
  for (int i = kFFT + 1; i < kFFT * 2; i++) s->mask[i] = 0.0;
  arm_cfft_f32(&arm_cfft_sR_f32_len512, s->mask, 0, 1);
}

/* T41/Filter.cpp:396-418 */
void DesignDecIntFilters(t41o_stream *s) {
  int filter_BW_highest = s->prm.f_hi_cut;
  if (filter_BW_highest < -s->prm.f_lo_cut) filter_BW_highest = -s->prm.f_lo_cut;
  int LP_F_help = filter_BW_highest;
  if (LP_F_help > 10000) LP_F_help = 10000;
  DesignKaiserLowpass(s->dec1_coeffs, kDec1Taps, (float)LP_F_help, kNAtt, (float)kSampleRate);
  DesignKaiserLowpass(s->dec2_coeffs, kDec2Taps, (float)LP_F_help, kNAtt, (float)(kSampleRate / kDF1));
  DesignKaiserLowpass(s->int1_coeffs, 48, (float)LP_F_help, kNAtt, (float)(kSampleRate / kDF1));
  DesignKaiserLowpass(s->int2_coeffs, 32, (float)LP_F_help, kNAtt, (float)kSampleRate);
}

/* T41/Filter.cpp:429-438 — run every NFM block (Process.cpp:259) */
void DesignDecFiltersFor(t41o_stream *s, int filter_BW) {
  int LP_F_help = filter_BW;
  DesignKaiserLowpass(s->dec1_coeffs, kDec1Taps, (float)LP_F_help, kNAtt, (float)kSampleRate);
  DesignKaiserLowpass(s->dec2_coeffs, kDec2Taps, (float)LP_F_help, kNAtt, (float)(kSampleRate / kDF1));
}

/* T41/Filter.cpp:235-249 (nfmBWFilterActive is false by default, ButtonProc.cpp:29) */
void RecalcFilters(t41o_stream *s) {
  float coef_i[kMaskTaps], coef_q[kMaskTaps];
  DesignComplexBandpass(coef_i, coef_q, kMaskTaps, (float)s->prm.f_lo_cut, (float)s->prm.f_hi_cut,
                        (float)kSampleRate / kDF);
  BuildFilterMask(s, coef_i, coef_q);
  DesignDecIntFilters(s);
}

/* T41/FFT.cpp:35-55 */
void ZoomPrep(t41o_stream *s) {
  int z = s->prm.spectrum_zoom;
  float Fstop_Zoom = 0.5 * (float)kSampleRate / (1 << z);
  DesignKaiserLowpass(s->zoom_fir_coeffs, 4, Fstop_Zoom, 60, (float)kSampleRate);
  s->zoom_fir_i.M = (uint8_t)(1 << z);
  s->zoom_fir_q.M = (uint8_t)(1 << z);
  s->zoom_iir_coeffs = (z >= 1 && z <= 4) ? t41o_zoom_iir[z - 1] : 0;
  s->zoom_iir_i.pCoeffs = s->zoom_iir_coeffs;
  s->zoom_iir_q.pCoeffs = s->zoom_iir_coeffs;
  s->zoom_sample_ptr = 0;
}

int ValidParams(const t41o_params *p) {
  switch (p->mode) {
    case T41O_DEMOD_USB: case T41O_DEMOD_LSB: case T41O_DEMOD_AM:
    case T41O_DEMOD_NFM: case T41O_DEMOD_PSK31: case T41O_DEMOD_SAM: break;
    default: return 0;
  }
  if (p->agc_mode < 0 || p->agc_mode > 4) return 0;
  if (p->spectrum_zoom < 0 || p->spectrum_zoom > 4) return 0;
  if (p->current_scale < 0 || p->current_scale > 4) return 0;
  if (p->f_hi_cut <= p->f_lo_cut) return 0;
  if (p->audio_volume < 0 || p->audio_volume > 100) return 0;
  if (p->nr_option < 0 || p->nr_option > 3) return 0;
  if (p->cw_filter_index < 0 || p->cw_filter_index > 5) return 0;
  return 1;
}

/* T41/Freq_Shift.cpp:42-65 */
void ShiftQuarterRate(t41o_stream *s) {
  float *L = s->bufL, *R = s->bufR;
  for (int i = 0; i < kBlock; i += 4) {
    float a, b;
    a = -R[i + 1]; b = L[i + 1]; L[i + 1] = a; R[i + 1] = b;
    a = -L[i + 2]; b = -R[i + 2]; L[i + 2] = a; R[i + 2] = b;
    a = R[i + 3]; b = -L[i + 3]; L[i + 3] = a; R[i + 3] = b;
  }
  memcpy(s->bufL_ex, L, sizeof(s->bufL_ex));
  memcpy(s->bufR_ex, R, sizeof(s->bufR_ex));
}

/* T41/Freq_Shift.cpp:94-141 (xmtMode == SSB_MODE: no side-tone shift) */
void ShiftNco(t41o_stream *s) {
  float NCO_INC = 2.0 * kPi * (long)(s->prm.nco_freq) / 192000.0;
  double OSC_COS = cos(NCO_INC);
  double OSC_SIN = sin(NCO_INC);
  double vq = s->osc_vect_q, vi = s->osc_vect_i;
  for (int i = 0; i < kBlock; i++) {
    double Osc_Q = (vq * OSC_COS) - (vi * OSC_SIN);
    double Osc_I = (vi * OSC_COS) + (vq * OSC_SIN);
    double Osc_Gain = 1.95 - ((vq * vq) + (vi * vi));
    vq = Osc_Gain * Osc_Q;
    vi = Osc_Gain * Osc_I;
    float freqAdjFactor = 1.1;
    s->bufL[i] = (s->bufL_ex[i] * freqAdjFactor * Osc_Q) + (s->bufR_ex[i] * freqAdjFactor * Osc_I);
    s->bufR[i] = (s->bufR_ex[i] * freqAdjFactor * Osc_Q) - (s->bufL_ex[i] * freqAdjFactor * Osc_I);
  }
  s->osc_vect_q = vq;
  s->osc_vect_i = vi;
}

/* shared tail of T41/FFT.cpp:136-157 / :234-246: power spectrum with half swap */
void PowerSpectrumSwapped(t41o_stream *s) {
  const float *b = s->spec_buf;
  for (int i = 0; i < kSpecRes / 2; i++) {
    s->fft_spec[i + kSpecRes / 2] = (b[i * 2] * b[i * 2] + b[i * 2 + 1] * b[i * 2 + 1]);
    int k = i + kSpecRes / 2;
    s->fft_spec[i] = (b[k * 2] * b[k * 2] + b[k * 2 + 1] * b[k * 2 + 1]);
  }
}

int16_t PixelFromPower(const t41o_stream *s, float power) {
  int scale = s->prm.current_scale;
  int v = t41o_base_offset[scale] + (int16_t)s->prm.pixel_offset +
          (int16_t)(t41o_db_scale[scale] * Log10Fast(power));
  return (int16_t)v;
}

/* T41/FFT.cpp:208-251 — zoom x1: first 512 conditioned samples, before FreqShift1 */
void SpectrumZoom1(t41o_stream *s) {
  float LPFcoeff = 0.7;
  for (int i = 0; i < kSpecRes; i++) {
    s->spec_buf[i * 2] = s->bufL[i] * (0.5 - 0.5 * cos(6.28 * i / kSpecRes));
    s->spec_buf[i * 2 + 1] = s->bufR[i] * (0.5 - 0.5 * cos(6.28 * i / kSpecRes));
  }
  arm_cfft_f32(&arm_cfft_sR_f32_len512, s->spec_buf, 0, 1);
  PowerSpectrumSwapped(s);
  for (int x = 0; x < kSpecRes; x++) {
    float spec_help = LPFcoeff * s->fft_spec[x] + (1.0 - LPFcoeff) * s->fft_spec_old[x];
    s->fft_spec_old[x] = spec_help;
    s->pixelnew[x] = PixelFromPower(s, s->fft_spec[x]);   /* unsmoothed value (B8) */
  }
}

/* T41/FFT.cpp:67-196 — zoom x2..x16 on the Fs/4-shifted block */
void SpectrumZoomN(t41o_stream *s) {
  const int z = s->prm.spectrum_zoom;
  float x_buffer[kBlock], y_buffer[kBlock];
  int sample_no = kBlock / (1 << z);
  if (sample_no > kSpecRes) sample_no = kSpecRes;
  arm_biquad_cascade_df1_f32(&s->zoom_iir_i, s->bufL, x_buffer, kBlock);
  arm_biquad_cascade_df1_f32(&s->zoom_iir_q, s->bufR, y_buffer, kBlock);
  arm_fir_decimate_f32(&s->zoom_fir_i, x_buffer, x_buffer, kBlock);
  arm_fir_decimate_f32(&s->zoom_fir_q, y_buffer, y_buffer, kBlock);
  for (int i = 0; i < sample_no; i++) {
    s->zoom_ring_x[s->zoom_sample_ptr] = x_buffer[i];
    s->zoom_ring_y[s->zoom_sample_ptr] = y_buffer[i];
    if (++s->zoom_sample_ptr >= kSpecRes) s->zoom_sample_ptr = 0;
  }
  float multiplier = (float)z;
  if (z > 3) multiplier = (float)(1 << z);
  for (int idx = 0; idx < kSpecRes; idx++) {
    s->spec_buf[idx * 2 + 0] = multiplier * s->zoom_ring_x[s->zoom_sample_ptr] * (0.5 - 0.5 * cos(6.28 * idx / kSpecRes));
    s->spec_buf[idx * 2 + 1] = multiplier * s->zoom_ring_y[s->zoom_sample_ptr] * (0.5 - 0.5 * cos(6.28 * idx / kSpecRes));
    if (++s->zoom_sample_ptr >= kSpecRes) s->zoom_sample_ptr = 0;
  }
  float LPFcoeff = 0.7;
  float onem_LPFcoeff = 1.0 - LPFcoeff;
  arm_cfft_f32(&arm_cfft_sR_f32_len512, s->spec_buf, 0, 1);
  PowerSpectrumSwapped(s);
  for (int i = 0; i < kSpecRes; i++) {
    s->fft_spec[i] = LPFcoeff * s->fft_spec[i] + onem_LPFcoeff * s->fft_spec_old[i];
    s->fft_spec_old[i] = s->fft_spec[i];
    s->pixelnew[i] = PixelFromPower(s, s->fft_spec[i]);   /* smoothed value */
  }
  /* T41/FFT.cpp:142-194: the "FDxxx<512 bytes>;" frame T41ControlSendData() writes to the control serial port:
     data = pixelnew + currentNF, shifted so that its maximum (at least 0) becomes 255, negatives clamped to 0;
     xxx = 255 - max.  sprintf's terminating NUL lands on byte 5 and is overwritten by the first data byte. */
  int16_t min = 0, max = 0;
  int16_t data[kSpecRes];
  for (int i = 0; i < kSpecRes; i++) {
    if (s->control_data_flag) {
      data[i] = s->pixelnew[i] + s->prm.current_nf;
      if (data[i] < min) min = data[i];
      if (data[i] > max) max = data[i];
    }
  }
  (void)min;
  char head[16];
  snprintf(head, sizeof(head), "FD%03d", 255 - max);
  memcpy(s->spec_data, head, strlen(head) + 1 < 8 ? strlen(head) + 1 : 8);
  s->spec_data[517] = ';';
  if (s->control_data_flag) {
    for (int i = 0; i < kSpecRes; i++) {
      int tmp = data[i] + 255 - max;
      if (tmp < 0) tmp = 0;
      s->spec_data[i + 5] = (uint8_t)tmp;
    }
    memcpy(s->sent_spec, s->spec_data, T41O_SPEC_FRAME_BYTES);     /* T41ControlSendData(specData, SPECTRUM_RES + 6) */
  }
}

/* T41/Display.cpp:343-358,459-466 for x1 = 0..510 (B18) */
void WaterfallRow(t41o_stream *s) {
  for (int x1 = 0; x1 < kSpecRes - 1; x1++) {
    int y = s->prm.spectrum_noise_floor - s->pixelnew[x1] - s->prm.current_nf;
    if (y > kSpectrumBottom) y = kSpectrumBottom;
    if (y < kSpectrumTopY) y = kSpectrumTopY;
    int idx = -y + 230;
    if (idx < 0) idx = 0;
    if (idx > 116) idx = 116;
    s->waterfall[x1] = t41o_gradient[idx];
  }
}

/* T41/DSP_Fn.cpp:479-632 on ifft_buf[512..1023] */
void AgcBlock(t41o_stream *s) {
  const AgcSetup &a = s->agc;
  float *io = s->ifft_buf + kFFT;
  if (s->prm.agc_mode == 0) {
    for (int i = 0; i < kDec; i++) {
      io[2 * i + 0] = a.fixed_gain * io[2 * i + 0];
      io[2 * i + 1] = a.fixed_gain * io[2 * i + 1];
    }
    return;
  }
  const unsigned ring_n = kRing;
  for (int i = 0; i < kDec; i++) {
    if (++s->agc_out_index >= (int)ring_n) s->agc_out_index -= ring_n;
    if (++s->agc_in_index >= ring_n) s->agc_in_index -= ring_n;
    const int oi = s->agc_out_index;
    const unsigned ii = s->agc_in_index;
    const float out_re = s->agc_ring[2 * oi + 0];
    const float out_im = s->agc_ring[2 * oi + 1];
    const float abs_out = s->agc_abs_ring[oi];
    const float re = io[2 * i + 0], im = io[2 * i + 1];
    s->agc_ring[2 * ii + 0] = re;
    s->agc_ring[2 * ii + 1] = im;
    s->agc_abs_ring[ii] = sqrtf(re * re + im * im);   /* pmode == 1 */

    s->agc_fast_backaverage = a.fast_backmult * abs_out + a.onemfast_backmult * s->agc_fast_backaverage;
    s->agc_hang_backaverage = a.hang_backmult * abs_out + a.onemhang_backmult * s->agc_hang_backaverage;

    /* outgoing sample was the window maximum: rescan the attack_buffsize newest entries */
    if ((abs_out >= s->agc_ring_max) && (abs_out > 0.0)) {
      s->agc_ring_max = 0.0;
      int k = oi;
      for (int j = 0; j < a.attack_buffsize; j++) {
        if (++k == (int)ring_n) k = 0;
        if (s->agc_abs_ring[k] > s->agc_ring_max) s->agc_ring_max = s->agc_abs_ring[k];
      }
    }
    if (s->agc_abs_ring[ii] > s->agc_ring_max) s->agc_ring_max = s->agc_abs_ring[ii];

    if (s->agc_hang_counter > 0) --s->agc_hang_counter;

    const float rm = s->agc_ring_max;
    float &v = s->agc_volts;
    const bool rising = rm >= v;
    if (rising) {
      /* every state: attack; states 2,3,4 also remember the level they left */
      if (s->agc_state >= 2) s->agc_save_volts = v;
      s->agc_state = 0;
      v += (rm - v) * a.attack_mult;
    } else {
      switch (s->agc_state) {
        case 0:
          if (v > a.pop_ratio * s->agc_fast_backaverage) {
            s->agc_state = 1;
            v += (rm - v) * a.fast_decay_mult;
          } else if (a.hang_enable && (s->agc_hang_backaverage > a.hang_level)) {
            s->agc_state = 2;
            s->agc_hang_counter = (int)(a.hangtime * kSampleRate / kDF);
            s->agc_decay_type = 1;
          } else {
            s->agc_state = 3;
            v += (rm - v) * a.decay_mult;
            s->agc_decay_type = 0;
          }
          break;
        case 1:
          if (v > s->agc_save_volts) {
            v += (rm - v) * a.fast_decay_mult;
          } else if (s->agc_hang_counter > 0) {
            s->agc_state = 2;
          } else if (s->agc_decay_type == 0) {
            s->agc_state = 3;
            v += (rm - v) * a.decay_mult;
          } else {
            s->agc_state = 4;
            v += (rm - v) * a.hang_decay_mult;
          }
          break;
        case 2:
          if (s->agc_hang_counter == 0) {
            s->agc_state = 4;
            v += (rm - v) * a.hang_decay_mult;
          }
          break;
        case 3:
          v += (rm - v) * a.decay_mult * .05;   /* double product (DSP_Fn.cpp:607) */
          break;
        case 4:
          v += (rm - v) * a.hang_decay_mult;
          break;
      }
    }
    if (v < a.min_volts) {
      v = a.min_volts;
      s->agc_action = 0;
    } else {
      s->agc_action = 1;
    }
    /* Arduino min() macro on (double 0.0, float): double arithmetic (DSP_Fn.cpp:628) */
    double lg = Log10Fast(a.inv_max_input * v);
    double clipped = (0.0 < lg) ? 0.0 : lg;
    float mult = (a.out_target - a.slope_constant * clipped) / v;
    io[2 * i + 0] = out_re * mult;
    io[2 * i + 1] = out_im * mult;
  }
}

/* T41/Demod.cpp:40-139 minus the TFT calls; the fade leveller reduces to the identity
   because exp(-1 / 24000 * tau) is exp(0) (integer division, B3), but is kept as written */
void DemodSam(t41o_stream *s) {
  float tauR = 0.02;
  float tauI = 1.4;
  float dc = 0.0;
  float dc_insert = 0.0;
  float mtauR = exp(-1 / 24000 * tauR);
  float onem_mtauR = 1.0 - mtauR;
  float mtauI = exp(-1 / 24000 * tauI);
  float onem_mtauI = 1.0 - mtauI;
  const float *in = s->ifft_buf + kFFT;
  for (int i = 0; i < kDec; i++) {
    float Sin = arm_sin_f32(s->sam_phzerror);
    float Cos = arm_cos_f32(s->sam_phzerror);
    float ai = Cos * in[i * 2];
    float bi = Sin * in[i * 2];
    float aq = Cos * in[i * 2 + 1];
    float bq = Sin * in[i * 2 + 1];
    float corr0 = +ai + bq;
    float corr1 = -bi + aq;
    float audio = (ai - bi) + (aq + bq);
    dc = mtauR * dc + onem_mtauR * audio;
    dc_insert = mtauI * dc_insert + onem_mtauI * corr0;
    audio = audio + dc_insert - dc;
    s->bufL[i] = audio;
    float det = Atan2Approx(corr1, corr0);
    s->sam_del_out = s->sam_fil_out;
    s->sam_omega2 = s->sam_omega2 + s->sam_g2 * det;
    if (s->sam_omega2 < s->sam_omega_min) s->sam_omega2 = s->sam_omega_min;
    else if (s->sam_omega2 > s->sam_omega_max) s->sam_omega2 = s->sam_omega_max;
    s->sam_fil_out = s->sam_g1 * det + s->sam_omega2;
    s->sam_phzerror = s->sam_phzerror + s->sam_del_out;
    while (s->sam_phzerror >= kTwoPi) s->sam_phzerror -= kTwoPi;
    while (s->sam_phzerror < 0.0) s->sam_phzerror += kTwoPi;
  }
}

/* T41/Demod.cpp:220-235, including the "sample 127" save (B4) */
void DemodNfm(t41o_stream *s, const float *input, float *output, int input_size) {
  output[0] = kFmDemodK * (input[0] * (input[1] - s->nfm_last_q) - input[1] * (input[0] - s->nfm_last_i)) /
              (input[0] * input[0] + input[1] * input[1]);
  for (int i = 1; i < input_size; i++) {
    float qnow = input[i * 2 + 1];
    float qlast = input[(i - 1) * 2 + 1];
    float inow = input[i * 2];
    float ilast = input[(i - 1) * 2];
    output[i] = kFmDemodK * (qnow * ilast - inow * qlast) / (inow * inow + qnow * qnow);
  }
  s->nfm_last_i = input[input_size - 2];
  s->nfm_last_q = input[input_size - 1];
}

/* T41/Process.cpp:955-967 */
float VolumeGain(int volume) {
  float x = volume / 100.0f;
  return 5 * x * x * x * x * x;
}

/* T41/Process.cpp:979-1016 with half_clip == quarter_clip == 0 (never set) */
void CodecGainStep(t41o_stream *s) {
  s->codec_timer++;
  if (s->codec_timer > 10000) s->codec_timer = 10000;
  if (s->codec_timer >= 50) {
    s->rf_gain += 1;
    s->codec_timer = 0;
    if (s->rf_gain > 15) s->rf_gain = 15;
  }
}

void InitStream(t41o_stream *s) {
  memset(s, 0, sizeof(*s));
  t41o_default_params(&s->prm);
  s->last_set_rf_gain = s->prm.rf_gain;
  s->rf_gain = s->prm.rf_gain;
  s->osc_vect_q = 1.0;
  for (int i = 0; i < 5; i++) arm_biquad_cascade_df2T_init_f32(&s->cw[i], 6, t41o_cw_coeffs[i], s->cw_state[i]);
  s->anr_lidx = 120.0;       /* T41/Noise.cpp:47,52 */
  s->nr_first_time_2 = 1;    /* T41/Noise.cpp:427 */
  s->anr_ngamma = 0.001;
  for (int i = 0; i < 14; i++) arm_biquad_cascade_df2T_init_f32(&s->eq[i], 4, t41o_eq_coeffs[i], s->eq_state[i]);
  s->osc_vect_i = 0.0;
  s->first_block = 1;
  s->agc_out_index = -1;

  arm_biquad_cascade_df2T_init_f32(&s->dc_block, 1, t41o_dc_block, s->dc_state);
  arm_fir_decimate_init_f32(&s->dec1_i, kDec1Taps, 4, s->dec1_coeffs, s->dec1_state_i, kBlock);
  arm_fir_decimate_init_f32(&s->dec1_q, kDec1Taps, 4, s->dec1_coeffs, s->dec1_state_q, kBlock);
  arm_fir_decimate_init_f32(&s->dec2_i, kDec2Taps, 2, s->dec2_coeffs, s->dec2_state_i, kBlock / 4);
  arm_fir_decimate_init_f32(&s->dec2_q, kDec2Taps, 2, s->dec2_coeffs, s->dec2_state_q, kBlock / 4);
  arm_fir_interpolate_init_f32(&s->int1, 2, 48, s->int1_coeffs, s->int1_state, kDec);
  arm_fir_interpolate_init_f32(&s->int2, 4, 32, s->int2_coeffs, s->int2_state, 2 * kDec);
  arm_biquad_cascade_df1_init_f32(&s->am_lp, 1, s->am_lp_coeffs, s->am_lp_state);
  arm_biquad_cascade_df1_init_f32(&s->zoom_iir_i, 4, 0, s->zoom_iir_state_i);
  arm_biquad_cascade_df1_init_f32(&s->zoom_iir_q, 4, 0, s->zoom_iir_state_q);
  arm_fir_decimate_init_f32(&s->zoom_fir_i, 4, 128, s->zoom_fir_coeffs, s->zoom_fir_state_i, kBlock);
  arm_fir_decimate_init_f32(&s->zoom_fir_q, 4, 128, s->zoom_fir_coeffs, s->zoom_fir_state_q, kBlock);

  /* AM low-pass designed once for 3 kHz, Q 1.3 (T41/T41_SDR.ino:560-566, B16): the
     default band is USB +200..+3000 so max(FHiCut, -FLoCut) = 3000 */
  DesignBiquadLowpass(s->am_lp_coeffs, (float)3000, 1.3, (float)kSampleRate / kDF);

  /* SAM PLL constants, T41/Demod.cpp:13-18 with omegaN = 200, pll_fmax = 4000 (gwv.cpp:64-65) */
  {
    const float omegaN = 200.0;
    const float pll_fmax = +4000.0;
    int zeta_help = 65;
    float zeta = (float)zeta_help / 100.0;
    s->sam_omega_min = kTwoPi * -pll_fmax * 1 / 24000;
    s->sam_omega_max = kTwoPi * pll_fmax * 1 / 24000;
    s->sam_g1 = 1.0 - exp(-2.0 * omegaN * zeta * 1 / 24000);
    s->sam_g2 = -s->sam_g1 + 2.0 * (1 - exp(-omegaN * zeta * 1 / 24000) * cosf(omegaN * 1 / 24000 * sqrtf(1.0 - zeta * zeta)));
  }

  AgcPrepDefaults(s->agc);
  AgcLoad(s->agc, s->prm.agc_mode, s->prm.agc_thresh);
  s->agc_in_index = s->agc.attack_buffsize + s->agc_out_index;
  RecalcFilters(s);
  ZoomPrep(s);
  t41o_psk31_reset(&s->psk);
}

}  // namespace

/* ------------------------------------------------------------------ */
/* C API                                                               */
/* ------------------------------------------------------------------ */
extern "C" {

void t41o_default_params(t41o_params *p) {
  memset(p, 0, sizeof(*p));
  p->mode = T41O_DEMOD_USB;       /* bands[] 20M..10M, T41/T41_SDR.ino:163-167 */
  p->f_lo_cut = 200;
  p->f_hi_cut = 3000;
  p->nco_freq = 0;                /* T41/T41_SDR.ino:793 */
  p->agc_mode = 1;                /* T41/gwv.cpp:15 */
  p->agc_thresh = 20;
  p->audio_volume = 30;           /* T41/gwv.cpp:16 */
  p->rf_gain_all_bands = 1;       /* T41/gwv.cpp:17 */
  p->rf_gain = 1;
  p->spectrum_zoom = 1;           /* T41/gwv.cpp:25 */
  p->current_scale = 1;           /* T41/gwv.cpp:24 */
  p->pixel_offset = 20;
  p->current_nf = 0;
  p->spectrum_noise_floor = 247;  /* T41/gwv.cpp:18 */
  p->nfm_filter_bw = 12000;       /* T41/Filter.cpp:16 */
  p->psk31_enable = 0;
  p->iq_amp_correction = 1.0f;    /* T41/gwv.cpp:70-71 */
  p->iq_phase_correction = 0.0f;
  p->receive_eq_flag = 0;
  for (int i = 0; i < 14; i++) p->equalizer_rec[i] = 100;   /* T41/EEPROM.cpp:59,698 */
  p->nr_option = 0;
  p->anr_notch_on = 0;
  p->cw_receive = 0;
  p->cw_filter_index = 5;
  p->nb_on = 0;
}

void t41o_mode_default_cuts(int32_t mode, int32_t *f_lo_cut, int32_t *f_hi_cut) {
  switch (mode) {
    case T41O_DEMOD_LSB: *f_hi_cut = -200; *f_lo_cut = -3000; break;
    case T41O_DEMOD_AM:
    case T41O_DEMOD_SAM: *f_hi_cut = 3000; *f_lo_cut = -3000; break;
    default: *f_hi_cut = 3000; *f_lo_cut = 200; break;
  }
}

t41o_stream *t41o_create(void) {
  t41o_stream *s = (t41o_stream *)malloc(sizeof(t41o_stream));
  if (!s) return 0;
  InitStream(s);
  return s;
}

void t41o_destroy(t41o_stream *s) { free(s); }

int t41o_set_params(t41o_stream *s, const t41o_params *p) {
  if (!s || !p || !ValidParams(p)) return -1;
  const t41o_params old = s->prm;
  s->prm = *p;
  if (p->rf_gain != s->last_set_rf_gain) {
    s->last_set_rf_gain = p->rf_gain;
    s->rf_gain = p->rf_gain;
  }
  if (p->mode != old.mode || p->f_lo_cut != old.f_lo_cut || p->f_hi_cut != old.f_hi_cut) RecalcFilters(s);
  if (p->agc_mode != old.agc_mode || p->agc_thresh != old.agc_thresh) {
    AgcLoad(s->agc, p->agc_mode, p->agc_thresh);           /* AGCOptions, T41/MenuProc.cpp:275-284 */
    s->agc_in_index = s->agc.attack_buffsize + s->agc_out_index;
  }
  if (p->spectrum_zoom != old.spectrum_zoom) ZoomPrep(s);  /* SetZoom, T41/Display.cpp:1402-1417 */
  return 0;
}

void t41o_get_params(const t41o_stream *s, t41o_params *p) { *p = s->prm; }

void t41o_get_tables(const t41o_stream *s, t41o_tables *t) {
  memset(t, 0, sizeof(*t));
  memcpy(t->dec1, s->dec1_coeffs, sizeof(t->dec1));
  memcpy(t->dec2, s->dec2_coeffs, sizeof(t->dec2));
  memcpy(t->int1, s->int1_coeffs, sizeof(t->int1));
  memcpy(t->int2, s->int2_coeffs, sizeof(t->int2));
  memcpy(t->mask, s->mask, sizeof(t->mask));
  memcpy(t->am_lp, s->am_lp_coeffs, sizeof(t->am_lp));
  memcpy(t->zoom_fir, s->zoom_fir_coeffs, sizeof(t->zoom_fir));
  const AgcSetup &a = s->agc;
  const float v[16] = {a.max_gain, a.attack_mult, a.decay_mult, a.fast_decay_mult, a.fast_backmult,
                       a.onemfast_backmult, a.out_target, a.min_volts, a.slope_constant, a.inv_max_input,
                       a.hang_level, a.hang_backmult, a.onemhang_backmult, a.hang_decay_mult, a.hangtime,
                       a.fixed_gain};
  memcpy(t->agc, v, sizeof(v));
  t->attack_buffsize = a.attack_buffsize;
  t->hang_counter_load = (int)(a.hangtime * kSampleRate / kDF);
}

void t41o_get_debug(const t41o_stream *s, t41o_debug *d) {
  memset(d, 0, sizeof(*d));
  d->agc_state = s->agc_state;
  d->agc_decay_type = s->agc_decay_type;
  d->agc_hang_counter = s->agc_hang_counter;
  d->agc_action = s->agc_action;
  d->rf_gain = s->rf_gain;
  d->codec_timer = (int32_t)s->codec_timer;
  d->zoom_sample_ptr = s->zoom_sample_ptr;
  d->first_block = s->first_block;
  d->agc_volts = s->agc_volts;
  d->agc_ring_max = s->agc_ring_max;
  d->agc_save_volts = s->agc_save_volts;
  d->agc_fast_backaverage = s->agc_fast_backaverage;
  d->agc_hang_backaverage = s->agc_hang_backaverage;
  d->sam_phzerror = s->sam_phzerror;
  d->sam_omega2 = s->sam_omega2;
  d->sam_fil_out = s->sam_fil_out;
  d->dc_state[0] = s->dc_state[0];
  d->dc_state[1] = s->dc_state[1];
  d->am_wold = s->am_wold;
  d->osc_vect_q = s->osc_vect_q;
  d->osc_vect_i = s->osc_vect_i;
}

/* T41/Filter.cpp:117-165 DoReceiveEQ: the 256 demodulated samples through 14 third-octave band-pass cascades
   (4 DF2T biquads each, T41/FIR.cpp:279-371), bands scaled by -/+ equalizerRec / 100 alternately and added in band
   order */
static void ReceiveEq(t41o_stream *s) {
  float *L = s->bufL;
  float band[14][kDec];
  float scale[14];
  for (int i = 0; i < 14; i++) scale[i] = (float)s->prm.equalizer_rec[i] / 100.0;
  for (int i = 0; i < 14; i++) arm_biquad_cascade_df2T_f32(&s->eq[i], L, band[i], kDec);
  for (int i = 0; i < 14; i++) arm_scale_f32(band[i], (i & 1) ? scale[i] : -scale[i], band[i], kDec);
  arm_add_f32(band[0], band[1], L, kDec);
  for (int i = 2; i < 14; i++) arm_add_f32(L, band[i], L, kDec);
}

/* T41/Noise.cpp:322-369 Xanr: variable-leak LMS (wdsp), 64 taps, delay 16, 512-deep delay line; constants
   T41/Noise.cpp:39-53.  Unsuffixed literals make several expressions FP64, as written there. */
static void Xanr(t41o_stream *s, int notch, const float *in, float *out) {
  const int ANR_delay = 16, ANR_mask = 511, ANR_taps = 64;
  const float ANR_den_mult = 6.25e-10, ANR_gamma = 0.1, ANR_lidx_min = 120.0, ANR_lidx_max = 200.0, ANR_lincr = 1.0,
              ANR_ldecr = 3.0, ANR_two_mu = 0.0001;
  int idx;
  float c0, c1;
  float y, error, sigma, inv_sigp;
  float nel, nev;
  for (int i = 0; i < kDec; i++) {
    s->anr_d[s->anr_in_idx] = in[i];
    y = 0;
    sigma = 0;
    for (int j = 0; j < ANR_taps; j++) {
      idx = (s->anr_in_idx + j + ANR_delay) & ANR_mask;
      y += s->anr_w[j] * s->anr_d[idx];
      sigma += s->anr_d[idx] * s->anr_d[idx];
    }
    inv_sigp = 1.0 / (sigma + 1e-10);
    error = s->anr_d[s->anr_in_idx] - y;
    if (notch) out[i] = error;
    else out[i] = y;
    if ((nel = error * (1.0 - ANR_two_mu * sigma * inv_sigp)) < 0.0) nel = -nel;
    if ((nev = s->anr_d[s->anr_in_idx] - (1.0 - ANR_two_mu * s->anr_ngamma) * y - ANR_two_mu * error * sigma * inv_sigp) < 0.0)
      nev = -nev;
    if (nev < nel) {
      if ((s->anr_lidx += ANR_lincr) > ANR_lidx_max) s->anr_lidx = ANR_lidx_max;
      else if ((s->anr_lidx -= ANR_ldecr) < ANR_lidx_min) s->anr_lidx = ANR_lidx_min;
    }
    s->anr_ngamma = ANR_gamma * (s->anr_lidx * s->anr_lidx) * (s->anr_lidx * s->anr_lidx) * ANR_den_mult;
    c0 = 1.0 - ANR_two_mu * s->anr_ngamma;
    c1 = ANR_two_mu * error * inv_sigp;
    for (int j = 0; j < ANR_taps; j++) {
      idx = (s->anr_in_idx + j + ANR_delay) & ANR_mask;
      s->anr_w[j] = c0 * s->anr_w[j] + c1 * s->anr_d[idx];
    }
    s->anr_in_idx = (s->anr_in_idx + ANR_mask) & ANR_mask;
  }
}

/* ---- spectral noise-reduction stages (T41/Noise.cpp:108-311, 379-655), 256-point frames with 128 new samples each ---- */
static const float kNrPsi = 0.0, kNrAlpha = 0.95, kNrBeta = 0.85;     /* EEPROM.cpp:71-73 (gwv.cpp defaults) */
enum { kNrL = 256, kNrHalf = 128, kNrLFrames = 3, kNrNFrames = 15 };

/* the voice-activity bin range both stages derive from the filter cut-offs (Noise.cpp:141-172, 429-443, 515-529) */
static void NrVadRange(const t41o_stream *s, uint8_t *lo, uint8_t *hi) {
  float lf_freq, uf_freq;
  const int flo = s->prm.f_lo_cut, fhi = s->prm.f_hi_cut;
  if (flo <= 0 && fhi >= 0) {
    lf_freq = 0.0;
    uf_freq = fmax(-(float)flo, (float)fhi);
  } else if (flo > 0) {
    lf_freq = (float)flo;
    uf_freq = (float)fhi;
  } else {
    uf_freq = -(float)flo;
    lf_freq = -(float)fhi;
  }
  lf_freq /= (((float)kSampleRate / kDF) / kNrL);
  uf_freq /= (((float)kSampleRate / kDF) / kNrL);
  uint8_t VAD_low = (int)lf_freq, VAD_high = (int)uf_freq;
  if (VAD_low == VAD_high) VAD_high++;
  if (VAD_low < 1) VAD_low = 1;
  else if (VAD_low > kNrL / 2 - 2) VAD_low = kNrL / 2 - 2;
  if (VAD_high < 1) VAD_high = 1;
  else if (VAD_high > kNrL / 2) VAD_high = kNrL / 2;
  *lo = VAD_low;
  *hi = VAD_high;
}

/* frame k of a block: previous 128 samples | the block's samples 128 k .. 128 k + 127, imaginary parts zero */
static void NrLoadFrame(t41o_stream *s, const float *L, int k, float *buf) {
  for (int i = 0; i < kNrHalf; i++) {
    buf[i * 2] = s->nr_last_sample[i];
    buf[i * 2 + 1] = 0.0;
  }
  for (int i = 0; i < kNrHalf; i++) s->nr_last_sample[i] = L[i + k * kNrHalf];
  for (int i = 0; i < kNrHalf; i++) {
    buf[kNrL + i * 2] = L[i + k * kNrHalf];
    buf[kNrL + i * 2 + 1] = 0.0;
  }
}

/* Kim & Ruwisch 2002 as the reference runs it (power instead of magnitude, gains clamped at 0), T41/Noise.cpp:108-311 */
static void Kim1Nr(t41o_stream *s, float *L, float *R) {
  float buf[2 * kNrL], out[kNrL];
  const float NR_KIM_K = 1.0;
  const float NR_onemalpha = (1.0 - kNrAlpha);
  const float NR_onemtwobeta = (1.0 - (2.0 * kNrBeta));
  uint8_t VAD_low, VAD_high;
  NrVadRange(s, &VAD_low, &VAD_high);
  for (int k = 0; k < 2; k++) {
    NrLoadFrame(s, L, k, buf);
    for (int idx = 0; idx < kNrL; idx++) {          /* Hann window, evaluated like the reference's expression */
      const float w = 0.5 * (float)(1.0 - (cosf(3.1415926535897932384626433832795 * 2.0 * (float)idx / (float)(kNrL - 1))));
      buf[idx * 2] *= w;
    }
    arm_cfft_f32(&arm_cfft_sR_f32_len256, buf, 0, 1);
    for (int i = 0; i < kNrHalf; i++) s->nr_X[i][s->nr_X_pointer] = (buf[i * 2] * buf[i * 2] + buf[i * 2 + 1] * buf[i * 2 + 1]);
    for (int i = VAD_low; i < VAD_high; i++) {
      float sum = 0.0;
      for (int j = 0; j < kNrLFrames; j++) sum = sum + s->nr_X[i][j];
      s->nr_E[i][s->nr_E_pointer] = sum / (float)kNrLFrames;
    }
    for (int i = VAD_low; i < VAD_high; i++) {
      s->nr_M[i] = s->nr_E[i][0];
      for (int j = 1; j < kNrNFrames; j++)
        if (s->nr_E[i][j] < s->nr_M[i]) s->nr_M[i] = s->nr_E[i][j];
    }
    for (int i = VAD_low; i < VAD_high; i++) {
      const float T = s->nr_X[i][s->nr_X_pointer] / s->nr_M[i];
      s->nr_lambda[i] = (T > kNrPsi) ? s->nr_M[i] : s->nr_E[i][s->nr_E_pointer];
    }
    for (int i = VAD_low; i < VAD_high; i++) {      /* NR_use_X == 0 */
      s->nr_G[i] = 1.0 - (s->nr_lambda[i] * NR_KIM_K / s->nr_E[i][s->nr_E_pointer]);
      if (s->nr_G[i] < 0.0) s->nr_G[i] = 0.0;
      s->nr_Gts[i][0] = kNrAlpha * s->nr_Gts[i][1] + (NR_onemalpha) * s->nr_G[i];
      s->nr_Gts[i][1] = s->nr_Gts[i][0];
    }
    for (int i = 1; i < (kNrHalf - 1); i++)
      s->nr_G[i] = kNrBeta * s->nr_Gts[i - 1][0] + NR_onemtwobeta * s->nr_Gts[i][0] + kNrBeta * s->nr_Gts[i + 1][0];
    s->nr_G[0] = (NR_onemtwobeta + kNrBeta) * s->nr_Gts[0][0] + kNrBeta * s->nr_Gts[1][0];
    s->nr_G[kNrHalf - 1] = kNrBeta * s->nr_Gts[kNrHalf - 2][0] + (NR_onemtwobeta + kNrBeta) * s->nr_Gts[kNrHalf - 1][0];
    for (int i = 0; i < kNrHalf; i++) {             /* the "conjugate symmetric" side pairs bin i with bin 255 - i, as written */
      buf[i * 2] = buf[i * 2] * s->nr_G[i];
      buf[i * 2 + 1] = buf[i * 2 + 1] * s->nr_G[i];
      buf[kNrL * 2 - i * 2 - 2] = buf[kNrL * 2 - i * 2 - 2] * s->nr_G[i];
      buf[kNrL * 2 - i * 2 - 1] = buf[kNrL * 2 - i * 2 - 1] * s->nr_G[i];
    }
    if (++s->nr_X_pointer >= kNrLFrames) s->nr_X_pointer = 0;
    if (++s->nr_E_pointer >= kNrNFrames) s->nr_E_pointer = 0;
    arm_cfft_f32(&arm_cfft_sR_f32_len256, buf, 1, 1);
    for (int i = 0; i < kNrHalf; i++) out[i + k * kNrHalf] = buf[i * 2] + s->nr_last_ifft[i];
    for (int i = 0; i < kNrHalf; i++) s->nr_last_ifft[i] = buf[kNrL + i * 2];
  }
  for (int i = 0; i < kNrL; i++) {
    L[i] = out[i];
    R[i] = L[i];
  }
}

/* T41/Noise.cpp:379-655.  Quirks kept: everything from the final weighting to the overlap-add sits inside the
   `first_time == 3` branch (the first 20 frames pass the audio through untouched); the musical-noise smoothing runs once
   per bin of the gain loop; NR_long_tone_gain is never written anywhere in the reference (all zeros in the host
   build), so once the stage has initialised its output is (signed) zero. */
static void SpectralNr(t41o_stream *s, float *L, float *R) {
  float buf[2 * kNrL];
  const float tinc = 0.00533333, tax = 0.0239, tap = 0.05062, psthr = 0.99, pnsaf = 0.01, asnr = 20, psini = 0.5, pspri = 0.5;
  const float ax = expf(-tinc / tax), ap = expf(-tinc / tap);
  const float xih1 = powf(10, (float)asnr / 10.0);
  if (!s->nr_consts_ready) {                 /* function statics: initialised on the first call */
    s->nr_xih1r = 1.0 / (1.0 + xih1) - 1.0;
    s->nr_pfac = (1.0 / pspri - 1.0) * (1.0 + xih1);
    s->nr_consts_ready = 1;
  }
  const float xih1r = s->nr_xih1r, pfac = s->nr_pfac;
  const float snr_prio_min = powf(10, -(float)20 / 20.0);
  const int16_t NR_width = 4;
  const float power_threshold = 0.4;
  float ph1y[kNrHalf];
  float xtr, pre_power, post_power, power_ratio;
  int16_t NN;
  uint8_t VAD_low, VAD_high;
  if (s->nr_first_time_2 == 1) {
    for (int i = 0; i < kNrHalf; i++) {
      s->nr_last_sample[i] = 0.0;
      s->nr_G[i] = 1.0;
      s->nr_Hk_old[i] = 1.0;
      s->nr_Nest[i][0] = 0.0;
      s->nr_Nest[i][1] = 1.0;
      s->nr_pslp[i] = 0.5;
    }
    s->nr_first_time_2 = 2;
  }
  for (int k = 0; k < 2; k++) {
    NrLoadFrame(s, L, k, buf);
    for (int idx = 0; idx < kNrL; idx++) buf[idx * 2] *= t41o_sqrt_hann[idx];
    arm_cfft_f32(&arm_cfft_sR_f32_len256, buf, 0, 1);
    for (int i = 0; i < kNrHalf; i++) s->nr_X[i][0] = (buf[i * 2] * buf[i * 2] + buf[i * 2 + 1] * buf[i * 2 + 1]);
    if (s->nr_first_time_2 == 2) {
      for (int i = 0; i < kNrHalf; i++) {
        s->nr_Nest[i][0] = s->nr_Nest[i][0] + 0.05 * s->nr_X[i][0];
        s->nr_xt[i] = psini * s->nr_Nest[i][0];
      }
      s->nr_init_counter++;
      if (s->nr_init_counter > 19) {
        s->nr_init_counter = 0;
        s->nr_first_time_2 = 3;
      }
    }
    if (s->nr_first_time_2 == 3) {
      for (int i = 0; i < kNrHalf; i++) {
        ph1y[i] = 1.0 / (1.0 + pfac * expf(xih1r * s->nr_X[i][0] / s->nr_xt[i]));
        s->nr_pslp[i] = ap * s->nr_pslp[i] + (1.0 - ap) * ph1y[i];
        if (s->nr_pslp[i] > psthr) ph1y[i] = 1.0 - pnsaf;
        else ph1y[i] = fmin(ph1y[i], 1.0);
        xtr = (1.0 - ph1y[i]) * s->nr_X[i][0] + ph1y[i] * s->nr_xt[i];
        s->nr_xt[i] = ax * s->nr_xt[i] + (1.0 - ax) * xtr;
      }
      for (int i = 0; i < kNrHalf; i++) {
        s->nr_SNR_post[i] = fmax(fmin(s->nr_X[i][0] / s->nr_xt[i], 1000.0), snr_prio_min);
        s->nr_SNR_prio[i] = fmax(kNrAlpha * s->nr_Hk_old[i] + (1.0 - kNrAlpha) * fmax(s->nr_SNR_post[i] - 1.0, 0.0), 0.0);
      }
      NrVadRange(s, &VAD_low, &VAD_high);
      float v;
      for (int i = VAD_low; i < VAD_high; i++) {
        v = s->nr_SNR_prio[i] * s->nr_SNR_post[i] / (1.0 + s->nr_SNR_prio[i]);
        s->nr_G[i] = 1.0 / s->nr_SNR_post[i] * sqrtf((0.7212 * v + v * v));
        s->nr_Hk_old[i] = s->nr_SNR_post[i] * s->nr_G[i] * s->nr_G[i];
        /* musical-noise treatment: inside the gain loop, as written */
        pre_power = 0.0;
        post_power = 0.0;
        for (int m = VAD_low; m < VAD_high; m++) {
          pre_power += s->nr_X[m][0];
          post_power += s->nr_G[m] * s->nr_G[m] * s->nr_X[m][0];
        }
        power_ratio = post_power / pre_power;
        if (power_ratio > power_threshold) {
          power_ratio = 1.0;
          NN = 1;
        } else {
          NN = 1 + 2 * (int)(0.5 + NR_width * (1.0 - power_ratio / power_threshold));
        }
        for (int b = VAD_low + NN / 2; b < VAD_high - NN / 2; b++) {
          s->nr_Nest[b][0] = 0.0;
          for (int m = b - NN / 2; m <= b + NN / 2; m++) s->nr_Nest[b][0] += s->nr_G[m];
          s->nr_Nest[b][0] /= (float)NN;
        }
        for (int b = VAD_low; b < VAD_low + NN / 2; b++) {
          s->nr_Nest[b][0] = 0.0;
          for (int m = b; m < (b + NN); m++) s->nr_Nest[b][0] += s->nr_G[m];
          s->nr_Nest[b][0] /= (float)NN;
        }
        for (int b = VAD_high - NN; b < VAD_high; b++) {
          s->nr_Nest[b][0] = 0.0;
          for (int m = b; m > (b - NN); m--) s->nr_Nest[b][0] += s->nr_G[m];
          s->nr_Nest[b][0] /= (float)NN;
        }
        for (int b = VAD_low + NN / 2; b < VAD_high - NN / 2; b++) s->nr_G[b] = s->nr_Nest[b][0];
      }
      for (int i = 0; i < kNrHalf; i++) {
        buf[i * 2] = buf[i * 2] * s->nr_G[i] * s->nr_long_tone_gain[i];
        buf[i * 2 + 1] = buf[i * 2 + 1] * s->nr_G[i] * s->nr_long_tone_gain[i];
        buf[kNrL * 2 - i * 2 - 2] = buf[kNrL * 2 - i * 2 - 2] * s->nr_G[i] * s->nr_long_tone_gain[i];
        buf[kNrL * 2 - i * 2 - 1] = buf[kNrL * 2 - i * 2 - 1] * s->nr_G[i] * s->nr_long_tone_gain[i];
      }
      arm_cfft_f32(&arm_cfft_sR_f32_len256, buf, 1, 1);
      for (int idx = 0; idx < kNrL; idx++) buf[idx * 2] *= t41o_sqrt_hann[idx];
      for (int i = 0; i < kNrHalf; i++) {
        L[i + k * kNrHalf] = buf[i * 2] + s->nr_last_ifft[i];
        R[i + k * kNrHalf] = L[i + k * kNrHalf];
      }
      for (int i = 0; i < kNrHalf; i++) s->nr_last_ifft[i] = buf[kNrL + i * 2];
    }
  }
}

/* T41/DSP_Fn.cpp:105-362 NoiseBlanker / AltNoiseBlanking: LPC (order 10, Levinson-Durbin on the block's
   autocorrelation), inverse + matched filtering to expose impulses, threshold 2.5 sigma, and 7 samples around each
   impulse replaced by the windowed sum of a forward and a backward linear prediction */
static void NoiseBlank(t41o_stream *s, float *insamp, float *outsamp) {
  enum { Nsam = 256, order = 10, impulse_length = 7, PL = 3, boundary_blank = 14 };
  const float NB_thresh = 2.5;
  int impulse_positions[20];
  int search_pos = 0, impulse_count = 0;
  arm_fir_instance_f32 LPC;
  float lpcs[order + 1], reverse_lpcs[order + 1], firState[Nsam + order], tempsamp[Nsam];
  float sigma2, lpc_power, impulse_threshold;
  float R[11], k, alfa, any[order + 1];
  float Rfw[impulse_length + order], Rbw[impulse_length + order], Wfw[impulse_length], Wbw[impulse_length];
  float sacc;
  memset(R, 0, sizeof(R));
  for (int i = 0; i < impulse_length; i++) {
    Wbw[i] = 1.0 * i / (impulse_length - 1);
    Wfw[impulse_length - i - 1] = Wbw[i];
  }
  for (int i = 0; i < (order + 1); i++) arm_dot_prod_f32(&insamp[0], &insamp[i], Nsam - i, &R[i]);
  R[0] = R[0] * (1.0 + 1.0e-9);
  lpcs[0] = 1;
  for (int i = 1; i < order + 1; i++) lpcs[i] = 0;
  alfa = R[0];
  for (int m = 1; m <= order; m++) {
    sacc = 0.0;
    for (int u = 1; u < m; u++) sacc = sacc + lpcs[u] * R[m - u];
    k = -(R[m] + sacc) / alfa;
    for (int v = 1; v < m; v++) any[v] = lpcs[v] + k * lpcs[m - v];
    for (int w = 1; w < m; w++) lpcs[w] = any[w];
    lpcs[m] = k;
    alfa = alfa * (1 - k * k);
  }
  for (int o = 0; o < order + 1; o++) reverse_lpcs[order - o] = lpcs[o];
  arm_fir_init_f32(&LPC, order + 1, &reverse_lpcs[0], &firState[0], Nsam);
  arm_fir_f32(&LPC, insamp, tempsamp, Nsam);
  arm_fir_init_f32(&LPC, order + 1, &lpcs[0], &firState[0], Nsam);
  arm_fir_f32(&LPC, tempsamp, tempsamp, Nsam);
  arm_var_f32(tempsamp, Nsam, &sigma2);
  arm_power_f32(lpcs, order, &lpc_power);
  impulse_threshold = NB_thresh * sqrtf(sigma2 * lpc_power);
  search_pos = order + PL;
  impulse_count = 0;
  do {
    if ((tempsamp[search_pos] > impulse_threshold) || (tempsamp[search_pos] < (-impulse_threshold))) {
      impulse_positions[impulse_count] = search_pos - order;
      impulse_count++;
      search_pos += PL;
    }
    search_pos++;
  } while (((unsigned int)search_pos < Nsam - (unsigned int)boundary_blank) && ((unsigned int)impulse_count < 20U));
  arm_negate_f32(&lpcs[1], &lpcs[1], order);
  arm_negate_f32(&reverse_lpcs[0], &reverse_lpcs[0], order);
  for (int j = 0; j < impulse_count; j++) {
    for (int q = 0; q < order; q++) {
      if ((impulse_positions[j] - PL - order + q) < 0) Rfw[q] = s->nb_last_frame_end[impulse_positions[j] + q];
      else Rfw[q] = insamp[impulse_positions[j] - PL - order + q];
      Rbw[impulse_length + q] = insamp[impulse_positions[j] + PL + q + 1];
    }
    for (int i = 0; i < impulse_length; i++) {
      arm_dot_prod_f32(&reverse_lpcs[0], &Rfw[i], order, &Rfw[i + order]);
      arm_dot_prod_f32(&lpcs[1], &Rbw[impulse_length - i], order, &Rbw[impulse_length - i - 1]);
    }
    arm_mult_f32(&Wfw[0], &Rfw[order], &Rfw[order], impulse_length);
    arm_mult_f32(&Wbw[0], &Rbw[0], &Rbw[0], impulse_length);
    arm_add_f32(&Rfw[order], &Rbw[0], &insamp[impulse_positions[j] - PL], impulse_length);
  }
  for (int p = 0; p < (order + PL); p++) s->nb_last_frame_end[p] = insamp[Nsam - 1 - order - PL + p];
  for (int q = 0; q < Nsam; q++) outsamp[q] = insamp[q];
}

/* Arduino map() with a float first argument, as the Teensyduino core overloads it (cores/teensy4/wiring.h; the
   core is not under /root/reference and no version is pinned): all arithmetic in the argument's type. */
static inline float MapFloat(float x, int in_min, int in_max, int out_min, int out_max) {
  return (x - (float)in_min) * ((float)out_max - (float)out_min) / ((float)in_max - (float)in_min) + (float)out_min;
}

/* T41/Process.cpp:550-570 (offset 50; LSB reads the mirrored side) and :791-805 (NFM, offset 20): squares of the
   masked spectrum's 1024 floats stored reversed, 3-point average -> 15 log10 -> map() -> int pixel, clamped at 0;
   the largest square feeds the S-meter's running average.  The float -> int conversion of -inf / NaN (empty
   spectrum) is undefined in C++ and INT_MIN or 0 on the targets; both end up 0 after the clamp. */
static void AudioSpectrum(t41o_stream *s, int mode) {
  float sq[2 * kFFT];
  for (int k = 0; k < 2 * kFFT; k++) sq[2 * kFFT - 1 - k] = (s->ifft_buf[k] * s->ifft_buf[k]);
  const int offset = (mode == T41O_DEMOD_NFM) ? 20 : 50;
  for (int k = 0; k < T41O_AUDIO_SPEC_PIXELS; k++) {
    float v;
    if (mode == T41O_DEMOD_USB || mode == T41O_DEMOD_AM || mode == T41O_DEMOD_SAM || mode == T41O_DEMOD_NFM) {
      v = offset + MapFloat(15 * log10f((sq[1021 - k] + sq[1022 - k] + sq[1023 - k]) / 3), 0, 100, 0, 120);
    } else if (mode == T41O_DEMOD_LSB) {
      v = offset + MapFloat(15 * log10f((sq[k] + sq[k + 1] + sq[k + 2]) / 3), 0, 100, 0, 120);
    } else {
      continue;
    }
    s->audio_ypixel[k] = (v >= 0.0f) ? (int)v : 0;       /* false for NaN too */
  }
  float mx = sq[0];
  for (int k = 1; k < 2 * kFFT; k++) if (sq[k] > mx) mx = sq[k];     /* arm_max_f32 */
  s->audio_max_sq_ave = .5 * mx + .5 * s->audio_max_sq_ave;
}

int t41o_process_block(t41o_stream *s, const float *iq, float *audio, int update_display,
                       int16_t *spec_row, uint16_t *wf_row, int8_t *psk_bit, uint8_t *psk_char) {
  if (!s || !iq || !audio) return -1;
  memset(s->sent_spec, 0, sizeof(s->sent_spec));
  memset(s->sent_audio, 0, sizeof(s->sent_audio));
  const int mode = s->prm.mode;
  float *L = s->bufL, *R = s->bufR;

  /* boundary: the caller hands over what arm_q15_to_float produced (Process.cpp:107-108) */
  for (int i = 0; i < kBlock; i++) {
    L[i] = iq[2 * i];
    R[i] = iq[2 * i + 1];
  }

  /* T41/Process.cpp:117-119 */
  float rfGainValue = pow(10, (float)s->prm.rf_gain_all_bands / 20);
  arm_scale_f32(L, rfGainValue, L, kBlock);
  arm_scale_f32(R, rfGainValue, R, kBlock);
  /* T41/Process.cpp:127-128 — one state for I then Q (B6) */
  arm_biquad_cascade_df2T_f32(&s->dc_block, L, L, kBlock);
  arm_biquad_cascade_df2T_f32(&s->dc_block, R, R, kBlock);
  /* T41/Process.cpp:133-134 */
  arm_scale_f32(L, (float)s->rf_gain, L, kBlock);
  arm_scale_f32(R, (float)s->rf_gain, R, kBlock);
  /* T41/Process.cpp:165-174 + Utility.cpp:178-187 (B14) */
  if (mode == T41O_DEMOD_LSB || mode == T41O_DEMOD_AM || mode == T41O_DEMOD_SAM || mode == T41O_DEMOD_USB) {
    arm_scale_f32(L, -s->prm.iq_amp_correction, L, kBlock);
    float factor = s->prm.iq_phase_correction;
    float tmpbuf[kBlock];
    if (factor < 0.0) {
      arm_scale_f32(L, factor, tmpbuf, kBlock);
      arm_add_f32(R, tmpbuf, R, kBlock);
    } else {
      arm_scale_f32(R, factor, tmpbuf, kBlock);
      arm_add_f32(L, tmpbuf, L, kBlock);
    }
  }

  /* T41/Process.cpp:185-187 */
  if (s->prm.spectrum_zoom == 0 && update_display) SpectrumZoom1(s);
  /* T41/Process.cpp:201 */
  ShiftQuarterRate(s);
  /* T41/Process.cpp:212-215 */
  if (s->prm.spectrum_zoom != 0 && update_display) SpectrumZoomN(s);
  /* T41/Process.cpp:236 */
  ShiftNco(s);

  float *io = s->ifft_buf + kFFT;
  if (mode == T41O_DEMOD_NFM) {
    /* T41/Process.cpp:252-276 */
    DesignDecFiltersFor(s, s->prm.nfm_filter_bw);
    arm_fir_decimate_f32(&s->dec1_i, L, L, kBlock);
    arm_fir_decimate_f32(&s->dec1_q, R, R, kBlock);
    arm_fir_decimate_f32(&s->dec2_i, L, L, kBlock / 4);
    arm_fir_decimate_f32(&s->dec2_q, R, R, kBlock / 4);
    for (int i = 0; i < kDec; i++) {
      s->fft_buf[kFFT + i * 2] = L[i];
      s->fft_buf[kFFT + i * 2 + 1] = R[i];
    }
  } else if (mode == T41O_DEMOD_PSK31) {
    /* T41/Process.cpp:376-387 */
    arm_fir_decimate_f32(&s->dec1_i, L, L, kBlock);
    arm_fir_decimate_f32(&s->dec1_q, R, R, kBlock);
    arm_fir_decimate_f32(&s->dec2_i, L, L, kBlock / 4);
    arm_fir_decimate_f32(&s->dec2_q, R, R, kBlock / 4);
  } else {
    /* T41/Process.cpp:470-606 */
    arm_fir_decimate_f32(&s->dec1_i, L, L, kBlock);
    arm_fir_decimate_f32(&s->dec1_q, R, R, kBlock);
    arm_fir_decimate_f32(&s->dec2_i, L, L, kBlock / 4);
    arm_fir_decimate_f32(&s->dec2_q, R, R, kBlock / 4);
    float freqKHzFcut;
    if (mode == T41O_DEMOD_LSB) freqKHzFcut = -(float)s->prm.f_lo_cut * 0.001;
    else freqKHzFcut = (float)s->prm.f_hi_cut * 0.001;
    float volScaleFactor = 7.0874 * pow(freqKHzFcut, -1.232);
    arm_scale_f32(L, volScaleFactor, L, kDec);
    arm_scale_f32(R, volScaleFactor, R, kDec);
    if (s->first_block) {
      for (int i = 0; i < kFFT; i++) s->fft_buf[i] = 0.0;
      s->first_block = 0;
    } else {
      for (int i = 0; i < kDec; i++) {
        s->fft_buf[i * 2] = s->last_sample_l[i];
        s->fft_buf[i * 2 + 1] = s->last_sample_r[i];
      }
    }
    for (int i = 0; i < kDec; i++) {
      s->last_sample_l[i] = L[i];
      s->last_sample_r[i] = R[i];
      s->fft_buf[kFFT + i * 2] = L[i];
      s->fft_buf[kFFT + i * 2 + 1] = R[i];
    }
    arm_cfft_f32(&arm_cfft_sR_f32_len512, s->fft_buf, 0, 1);
    arm_cmplx_mult_cmplx_f32(s->fft_buf, s->mask, s->ifft_buf, kFFT);
    if (update_display) AudioSpectrum(s, mode);          /* T41/Process.cpp:550-570 */
    arm_cfft_f32(&arm_cfft_sR_f32_len512, s->ifft_buf, 1, 1);
    AgcBlock(s);
  }

  /* harness-defined PSK31 tap: one filtered complex sample per
     768 samples at 24 kS/s = the first sample of every third block */
  int8_t bit_out = -1;
  uint8_t char_out = 0;
  if (s->prm.psk31_enable && mode != T41O_DEMOD_NFM && mode != T41O_DEMOD_PSK31) {
    if (s->psk_block_count % 3u == 0u) {
      uint8_t bit = t41o_psk31_dbpsk_bit(&s->psk, io[0], io[1]);
      bit_out = (int8_t)bit;
      char_out = t41o_psk31_varicode_push(&s->psk, bit);
    }
    s->psk_block_count++;
  }
  if (psk_bit) *psk_bit = bit_out;
  if (psk_char) *psk_char = char_out;

  /* T41/Process.cpp:615-761 */
  switch (mode) {
    case T41O_DEMOD_USB:
    case T41O_DEMOD_LSB:
      for (int i = 0; i < kDec; i++) {
        L[i] = io[i * 2];
        R[i] = L[i];
      }
      break;
    case T41O_DEMOD_AM:
      for (int i = 0; i < kDec; i++) {
        float audiotmp = AlphaBetaMagnitude(io[i * 2], io[i * 2 + 1]);
        float w = audiotmp + s->am_wold * 0.99f;
        L[i] = w - s->am_wold;
        s->am_wold = w;
      }
      arm_biquad_cascade_df1_f32(&s->am_lp, L, R, kDec);
      arm_copy_f32(R, L, kDec);
      break;
    case T41O_DEMOD_NFM:
      DemodNfm(s, &s->fft_buf[kFFT], L, kDec);
      for (int i = 1; i < kDec; i++) {     /* limiter skips index 0 (B5) */
        float tmp = L[i];
        tmp = (1 < tmp) ? 1 : tmp;
        tmp = (-1 > tmp) ? -1 : tmp;
        L[i] = tmp;
      }
      break;
    case T41O_DEMOD_SAM:
      DemodSam(s);
      break;
    default:
      break;
  }

  /* T41/Process.cpp:765-816 — NFM audio through the overlap-save filter and the AGC */
  if (mode == T41O_DEMOD_NFM) {
    for (int i = 0; i < kDec; i++) {
      s->fft_buf[i * 2] = s->last_sample_l[i];
      s->fft_buf[i * 2 + 1] = 0;
      s->last_sample_l[i] = L[i];
      s->fft_buf[kFFT + i * 2] = L[i];
      s->fft_buf[kFFT + i * 2 + 1] = 0;
    }
    arm_cfft_f32(&arm_cfft_sR_f32_len512, s->fft_buf, 0, 1);
    arm_cmplx_mult_cmplx_f32(s->fft_buf, s->mask, s->ifft_buf, kFFT);
    if (update_display) AudioSpectrum(s, mode);          /* T41/Process.cpp:791-805 */
    arm_cfft_f32(&arm_cfft_sR_f32_len512, s->ifft_buf, 1, 1);
    AgcBlock(s);
    for (int i = 0; i < kDec; i++) L[i] = io[i * 2];
  }

  /* T41/Process.cpp:818-825: the audio-spectrum bytes for the control app (specData is reused as scratch) */
  if (update_display && s->control_data_flag) {
    for (int i = 0; i < T41O_AUDIO_SPEC_PIXELS; i++) {
      s->spec_data[i] = (uint8_t)(s->audio_ypixel[i] > 255 ? 255 : s->audio_ypixel[i]);
    }
    memcpy(s->sent_audio, s->spec_data, T41O_AUDIO_SPEC_PIXELS);
  }

  /* T41/Process.cpp:827-831 */
  if (s->prm.receive_eq_flag == 1) ReceiveEq(s);

  /* T41/Process.cpp:841-865.  LMS noise reduction (option 3): Xanr leaves its output in float_buffer_R, which
     nothing reads afterwards - what reaches the audio is float_buffer_L x 1.5, and the adaptive filter's state
     (shared with the notch) still advances.  Automatic notch: Xanr's error signal replaces float_buffer_L. */
  if (s->prm.nr_option == 1) {                 /* Process.cpp:845-849 */
    Kim1Nr(s, L, R);
    arm_scale_f32(L, 30, L, kDec);
  } else if (s->prm.nr_option == 2) {          /* Process.cpp:850-852 */
    SpectralNr(s, L, R);
  } else if (s->prm.nr_option == 3) {
    Xanr(s, 0, L, R);
    arm_scale_f32(L, 1.5, L, kDec);
  }
  if (s->prm.anr_notch_on == 1) {
    Xanr(s, 1, L, R);
    arm_copy_f32(R, L, kDec);
  }
  if (s->prm.nb_on != 0) {                     /* Process.cpp:873-876 */
    NoiseBlank(s, L, R);
    arm_copy_f32(R, L, kDec);
  }

  /* T41/Process.cpp:878-914: in the CW receive state (the decoder itself, DoCWReceiveProcessing, is not part of
     this chain) the selected 12-pole low-pass, each filter with its own state */
  if (s->prm.cw_receive == 1 && s->prm.cw_filter_index != 5) {
    arm_biquad_cascade_df2T_f32(&s->cw[s->prm.cw_filter_index], L, R, kDec);
    arm_copy_f32(R, L, kDec);
  }

  /* T41/Process.cpp:917-920 */
  arm_fir_interpolate_f32(&s->int1, L, s->ifft_buf, kDec);
  arm_fir_interpolate_f32(&s->int2, s->ifft_buf, L, 2 * kDec);
  /* T41/Process.cpp:925-931 */
  arm_scale_f32(L, kDF * VolumeGain(s->prm.audio_volume), L, kBlock);
  memcpy(audio, L, kBlock * sizeof(float));
  /* T41/Process.cpp:939 */
  CodecGainStep(s);

  if (update_display) {
    WaterfallRow(s);
    if (spec_row) memcpy(spec_row, s->pixelnew, sizeof(s->pixelnew));
    if (wf_row) memcpy(wf_row, s->waterfall, sizeof(s->waterfall));
  }
  return 0;
}

int t41o_process(t41o_stream *s, const float *iq, float *audio, int n_blocks, int row_every,
                 int16_t *spec_rows, uint16_t *wf_rows, int8_t *psk_bits, uint8_t *psk_chars) {
  int rows = 0;
  for (int b = 0; b < n_blocks; b++) {
    int upd = (row_every > 0) && (b % row_every == 0);
    int rc = t41o_process_block(s, iq + (size_t)b * 2 * kBlock, audio + (size_t)b * kBlock, upd,
                                (upd && spec_rows) ? spec_rows + (size_t)rows * kSpecRes : 0,
                                (upd && wf_rows) ? wf_rows + (size_t)rows * kSpecRes : 0,
                                psk_bits ? psk_bits + b : 0, psk_chars ? psk_chars + b : 0);
    if (rc) return rc;
    if (upd && s->cap_ypixel) {
      for (int k = 0; k < T41O_AUDIO_SPEC_PIXELS; k++) s->cap_ypixel[(size_t)rows * T41O_AUDIO_SPEC_PIXELS + k] = s->audio_ypixel[k];
    }
    if (upd && s->cap_max_ave) s->cap_max_ave[rows] = s->audio_max_sq_ave;
    if (upd && s->cap_frames) memcpy(s->cap_frames + (size_t)rows * T41O_SPEC_FRAME_BYTES, s->sent_spec, T41O_SPEC_FRAME_BYTES);
    if (upd && s->cap_audio_frames) memcpy(s->cap_audio_frames + (size_t)rows * T41O_AUDIO_SPEC_PIXELS, s->sent_audio, T41O_AUDIO_SPEC_PIXELS);
    rows += upd;
  }
  return rows;
}

void t41o_capture_audio_spectrum(t41o_stream *s, int32_t *ypixel_rows, float *max_ave_rows) {
  s->cap_ypixel = ypixel_rows;
  s->cap_max_ave = max_ave_rows;
}

void t41o_capture_control_frames(t41o_stream *s, uint8_t *spec_frame_rows, uint8_t *audio_frame_rows) {
  s->cap_frames = spec_frame_rows;
  s->cap_audio_frames = audio_frame_rows;
  s->control_data_flag = (spec_frame_rows != 0) || (audio_frame_rows != 0);
}

/* T41/Display.cpp:942,958,995-998: smeterPad = map(dbm, -73.0-9*6.0, -73.0, 0, 9*pixels_per_s) with the float overload
   of map(), into an int16_t, then max(0, .) and min(SMETER_BAR_LENGTH = 180, .) */
int32_t t41o_smeter_bar(float dbm) {
  const float pixels_per_s = 12;
  const float v = (dbm - (float)(-73.0 - 9 * 6.0)) * ((float)(9 * pixels_per_s) - (float)0) / ((float)-73.0 - (float)(-73.0 - 9 * 6.0)) + (float)0;
  int16_t pad = (v < -32768.0f || v != v) ? (int16_t)-32768 : (v > 32767.0f ? (int16_t)32767 : (int16_t)v);
  int r = pad;
  r = r > 0 ? r : 0;
  r = r < 180 ? r : 180;
  return r;
}

/* T41/Display.cpp:959-981 (TCVSDR_SMETER): dbm_calibration = 22.0, slope = 10.0, cons = -92 are floats, attenuator = 0
   (Display.cpp:146); the RFgain term's 1.5 is a double literal, so the tail of the sum is FP64 */
float t41o_smeter_dbm(float audio_max_sq_ave, float gain_correction, int32_t rf_gain, int32_t rf_gain_all_bands) {
  const float dbm_calibration = 22.0, slope = 10.0, cons = -92;
  const int attenuator = 0;
  float dbm = dbm_calibration + gain_correction + (float)attenuator + slope * Log10Fast(audio_max_sq_ave) + cons -
              (float)rf_gain * 1.5 - rf_gain_all_bands;
  return dbm;
}

void t41o_calc_fir_coeffs(float *coeffs, int num_coeffs, float fc, float astop, int type, float dfc, float fs) {
  (void)dfc;
  if (type != 0) abort();
  DesignKaiserLowpass(coeffs, num_coeffs, fc, astop, fs);
}

void t41o_calc_cplx_fir_coeffs(float *ci, float *cq, int num_coeffs, float f_lo, float f_hi, float fs) {
  DesignComplexBandpass(ci, cq, num_coeffs, f_lo, f_hi, fs);
}

float t41o_log10f_fast(float x) { return Log10Fast(x); }
float t41o_approx_atan2(float y, float x) { return Atan2Approx(y, x); }
void t41o_cfft512(float *buf, int inverse) { arm_cfft_f32(&arm_cfft_sR_f32_len512, buf, inverse ? 1 : 0, 1); }

/* ---- PSK31 primitives ---- */
void t41o_psk31_reset(t41o_psk31 *st) {
  st->last_phase = 0.0f;
  st->status_shr = 0;
}

/* T41/psk31.cpp:293-310 on a proper (I,Q) pair; psk31.cpp does not include FIR.h so
   PI here is Arduino's double constant and the wrap / threshold compares are FP64 (B22) */
uint8_t t41o_psk31_dbpsk_bit(t41o_psk31 *st, float i, float q) {
  const double PI_D = 3.1415926535897932384626433832795;
  float phase = Atan2Approx(q, i);
  float dphase = phase - st->last_phase;
  while (dphase < -PI_D) dphase += 2 * PI_D;
  while (dphase >= PI_D) dphase -= 2 * PI_D;
  uint8_t bit;
  if ((dphase > (PI_D / 2)) || (dphase < (-PI_D / 2))) bit = 0;
  else bit = 1;
  st->last_phase = phase;
  return bit;
}

/* T41/psk31.cpp:235-264 */
uint8_t t41o_psk31_varicode_push(t41o_psk31 *st, uint8_t symbol) {
  st->status_shr = (st->status_shr << 1) | (uint64_t)(!!symbol);
  if ((st->status_shr & 0xFFFull) == 0) return 0;
  for (int i = 0; i < 128; i++) {
    uint64_t want = ((uint64_t)t41o_varicode[i].code) << 2;
    unsigned nbits = (t41o_varicode[i].bits + 4u) & 63u;
    uint64_t keep = (nbits == 0) ? 0ull : (~0ull >> (64u - nbits));
    if (want == (st->status_shr & keep)) {
      st->status_shr = 0;
      return t41o_varicode[i].ascii;
    }
  }
  return 0;
}

int t41o_psk31_varicode_item(int index, uint64_t *code, uint8_t *ascii) {
  if (index < 0 || index >= 128) return -1;
  if (code) *code = t41o_varicode[index].code;
  if (ascii) *ascii = t41o_varicode[index].ascii;
  return t41o_varicode[index].bits;
}

}  // extern "C"
