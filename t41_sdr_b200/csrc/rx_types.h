/*
 * rx_types.h — plain-data structures shared by the host control path and the CUDA kernel.
 * Everything a virtual receiver carries between blocks lives in StreamState (HBM, one
 * struct per stream; SURVEY.md Appendix A lists the firmware statics each field replaces).
 */
#ifndef T41RX_TYPES_H
#define T41RX_TYPES_H

#include <stdint.h>

namespace t41rx {

constexpr int kBlock = 2048;      /* samples per block at 192 kS/s */
constexpr int kDec1Out = 512;     /* after /4 */
constexpr int kDec = 256;         /* after /8: 24 kS/s */
constexpr int kFft = 512;
constexpr int kDec1Taps = 28;     /* n_dec1_taps, T41_SDR.ino:344 */
constexpr int kDec2Taps = 46;     /* n_dec2_taps, T41_SDR.ino:345 */
constexpr int kInt1Taps = 48;
constexpr int kInt2Taps = 32;
constexpr int kMaskTaps = 257;    /* m_NumTaps, Filter.cpp:18 */
constexpr int kAgcDelay = 97;     /* attack_buffsize, DSP_Fn.cpp:409 (B11) */
constexpr int kAgcRing = 128;     /* power-of-two ring >= kAgcDelay + 1 (firmware: 1921, functionally a 97-sample delay) */
constexpr int kSpecRes = 512;
constexpr int kSpecFrameBytes = 518;    /* specData[], t41Control.cpp:21 */
constexpr int kAudioSpecPixels = 270;   /* AUDIO_SPEC_BOX_W - 2 = 800 - (0 + 1 + 512 + 15) - 2 (Display.h:6,15,20,42,45; Process.cpp:555) */

/* demodulation modes (SDT.h:57-68), same values as T41RX_DEMOD_* */
constexpr int kModeUsb = 0, kModeLsb = 1, kModeAm = 2, kModeNfm = 3, kModePsk31 = 5, kModeSam = 8;

/* AGC constants, DSP_Fn.cpp:408-434 */
struct AgcConsts {
  float max_gain, attack_mult, decay_mult, fast_decay_mult, fast_backmult, onemfast_backmult;
  float out_target, min_volts, slope_constant, inv_max_input, hang_level, hang_backmult;
  float onemhang_backmult, hang_decay_mult, hangtime, fixed_gain;
  float pop_ratio;
  int32_t hang_enable;
  int32_t hang_counter_load;   /* (int)(hangtime * SampleRate / DF), DSP_Fn.cpp:554 */
  int32_t pad_;
};

/* Tables shared by every receiver with the same (mode class, FLoCut, FHiCut, nfmFilterBW). */
struct FilterSet {
  float dec1[kDec1Taps];
  float dec2[kDec2Taps];
  float int1[kInt1Taps];
  float int2[kInt2Taps];
  float pad_[2];
  float mask[2 * kFft];        /* interleaved re,im; natural bin order */
};

/* Per-receiver constants derived from t41rx_params (recomputed on set_params only). */
struct StreamCfg {
  int32_t mode;
  int32_t agc_mode;
  int32_t zoom;                /* spectrumZoom index */
  int32_t filter_id;
  int32_t psk31_enable;
  int32_t mirrored;            /* 1 for USB/LSB/AM/SAM: I *= -amp and phase correction (Process.cpp:165-174) */
  int32_t pixel_add;           /* displayScale[].baseOffset + bands[].pixel_offset */
  int32_t wf_base;             /* spectrumNoiseFloor - currentNF */
  int32_t current_nf;          /* currentNF (FFT.cpp:161: serial frame data = pixelnew + currentNF) */
  int32_t eq_on;               /* receiveEQFlag == ON (Process.cpp:828) */
  int32_t nr_lms;              /* nrOptionSelect == 3 (Process.cpp:852-856) */
  int32_t anr_notch;           /* ANR_notchOn == 1 (Process.cpp:860-865) */
  int32_t cw_filter;           /* CW audio low-pass 0..4 in the chain (T41State == CW_RECEIVE, CWFilterIndex != 5), else -1 */
  int32_t nr_kim;              /* nrOptionSelect == 1: Kim1_NR, then x 30 (Process.cpp:845-849) */
  int32_t nr_spectral;         /* nrOptionSelect == 2: SpectralNoiseReduction (Process.cpp:850-852) */
  int32_t nb_on;               /* NB_on != 0: NoiseBlanker behind the notch (Process.cpp:873-876) */
  int32_t nr_vad_lo, nr_vad_hi;/* VAD_low / VAD_high: the bins the spectral stages work on (Noise.cpp:141-172) */
  int32_t pad_nr_;
  float eq_scale[14];          /* -/+ recEQ_LevelScale[i] = (float)equalizerRec[i] / 100.0, sign as Filter.cpp:136-149 */
  int32_t zoom_samples;        /* min(2048 >> zoom, 512), FFT.cpp:78-81 */
  int32_t nco_epoch;           /* bumped when NCOFreq changes: forces one exact block (amplitude transient) */
  float rf_gain_value;         /* pow(10, rfGainAllBands / 20), Process.cpp:117 */
  float neg_iq_amp;            /* -IQAmpCorrectionFactor */
  float iq_phase;              /* IQPhaseCorrectionFactor */
  float vol_scale;             /* 7.0874 * pow(fcut_kHz, -1.232), Process.cpp:490 */
  float volume;                /* DF * VolumeToAmplification(audioVolume), Process.cpp:929 */
  float db_scale;              /* displayScale[currentScale].dBScale */
  float zoom_mult;             /* FFT.cpp:105-108 */
  float zoom_fir[4];           /* Fir_Zoom_FFT_Decimate_coeffs */
  float am_lp[5];              /* biquad_lowpass1_coeffs */
  float pad_[1];
  AgcConsts agc;
  double osc_cos, osc_sin;     /* OSC_COS / OSC_SIN, Freq_Shift.cpp:123-124 */
  double nco_delta;            /* rotation angle of the (osc_cos, osc_sin) matrix per sample */
  double nco_block_delta;      /* (2048 * nco_delta) mod 2pi */
  double nco_rho;              /* sqrt(osc_cos^2 + osc_sin^2) */
  double nco_r2_fix;           /* fixed point of |V|^2: 1.95 - 1/rho */
  double nco_amp;              /* steady |Osc| = rho * sqrt(nco_r2_fix) */
};

/* Everything a receiver remembers between blocks. */
struct StreamState {
  /* input conditioning: HP_DC_Butter_state2 (Process.cpp:42), shared by I and Q (B6) */
  float dc_d1, dc_d2;
  /* FreqShift2 oscillator (Freq_Shift.cpp:13-14) */
  double osc_q, osc_i;         /* Osc_Vect_Q / Osc_Vect_I while tracked exactly */
  double nco_phase;            /* angle of Osc_Vect while in closed form */
  int32_t nco_closed;          /* 0: (osc_q, osc_i) authoritative; 1: nco_phase authoritative */
  int32_t nco_epoch_seen;
  /* decimators */
  float dec1_hist[2][kDec1Taps - 1];
  float dec2_hist[2][kDec2Taps - 1];
  /* overlap-save */
  float ola_prev[2][kDec];     /* last_sample_buffer_L / _R (T41_SDR.ino:403-404) */
  int32_t first_block;         /* Process.cpp:47 */
  /* AGC (DSP_Fn.cpp:482-492): 97-sample delay line + envelope */
  int32_t agc_pos;             /* ring index of the newest entry */
  float agc_re[kAgcRing], agc_im[kAgcRing], agc_abs[kAgcRing];
  float agc_fast_back, agc_hang_back, agc_ring_max, agc_save_volts, agc_volts;
  int32_t agc_hang_counter, agc_state, agc_decay_type, agc_action;
  /* demodulators */
  float am_wold;               /* Process.cpp:73 */
  float am_lp_state[4];        /* biquad_lowpass1_state */
  float sam_phzerror, sam_fil_out, sam_omega2;   /* Demod.cpp:19-23 */
  float nfm_last_i, nfm_last_q;                  /* Demod.cpp:221-222 */
  /* interpolators */
  float int1_hist[kInt1Taps / 2 - 1];
  float int2_hist[kInt2Taps / 4 - 1];
  /* Codec_gain (Process.cpp:980) */
  uint32_t codec_timer;
  int32_t rf_gain;
  /* PSK31 tap */
  float psk_last_phase;
  uint32_t psk_block_count;
  unsigned long long psk_shr;
  /* display spectrum (FFT.cpp:13-18,72-73) */
  float zoom_iir[2][16];
  float zoom_fir_hist[2][3];
  int32_t zoom_ptr;
  int32_t pad_;
  float zoom_ring[2][kSpecRes];
  float spec_old[kSpecRes];
  /* throughput kernel (rx_fast.cuh): its own forms of the DC-block state and of the oscillator angle, kept
     beside the reference forms above so that a run cut into several calls is bit-identical to one call;
     valid while fast_native != 0 (the bit-exact kernel clears it) */
  int32_t fast_native;
  float fast_dc_w;
  double fast_ph_re, fast_ph_im;
  /* audio-spectrum + S-meter by-product (Process.cpp:32,34): audioMaxSquaredAve and audioYPixel[0..269] */
  float audio_max_sq_ave;
  int32_t pad2_;
  /* receive equaliser: rec_EQ_Band1..14_state (Filter.cpp:43-56), 4 stages x (d1, d2) per band */
  float eq_state[14][8];
  /* variable-leak LMS of the automatic notch / LMS noise reduction (Noise.cpp:33-53): delay line, taps, state */
  float anr_d[512], anr_w[64];
  int32_t anr_in_idx;
  float anr_lidx, anr_ngamma;
  int32_t pad3_;
  /* CW audio low-passes: CW_AudioFilter1..5_state (CWProcessing.cpp:38-42), each filter keeps its own */
  float cw_state[5][12];
  int16_t audio_ypixel[kAudioSpecPixels + 2];
};

/* State of the 256-point spectral noise-reduction stages and the noise blanker (Noise.cpp:17-37,109-110,389-427,
   DSP_Fn.cpp:143): one per receiver, in an array of its own that only exists once a receiver switches one of these
   stages on.  All zero at start, like the reference's statics in the host build (spectral_stage = NR_first_time_2 - 1).
   Kim1_NR and SpectralNoiseReduction share the first group of arrays, as they do in the reference. */
struct NrState {
  float last_sample[128], last_ifft[128];
  float X[128][3], E[128][15], M[128], Nest[128][2], lambda[128], Gts[128][2], G[128];
  float SNR_prio[128], SNR_post[128], Hk_old[128], long_tone_gain[128];
  float pslp[128], xt[128];
  uint32_t x_ptr, e_ptr;
  int32_t spectral_stage, init_counter;
  float nb_last_frame_end[80];
};

}  // namespace t41rx
#endif
