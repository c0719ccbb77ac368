/*
 * oracle/t41_oracle.h — TEST INFRASTRUCTURE ("Tier-B oracle"), not product code.
 *
 * CPU restatement of the T41 receive chain (reference
 * /root/reference/software/T41_SDR/Process.cpp:70-944 `ProcessIQData` and the
 * stage functions it calls), one heap-allocated receiver per handle instead of
 * the firmware's globals.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Parity pin: the reference ships no tests or golden vectors, so
 * this oracle is pinned against the reference's own translation units compiled
 * in place (oracle/Makefile -> oracle/_ref/libt41ref.so, "Tier-A") and against
 * the fixtures under tests/golden/ that Tier-A generated.  CMSIS-DSP itself is
 * not available; see cmsis_port.h for how it is restated.
 */
#ifndef T41_ORACLE_H
#define T41_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* demodulation modes, values of T41/SDT.h:57-68 */
enum {
  T41O_DEMOD_USB = 0,
  T41O_DEMOD_LSB = 1,
  T41O_DEMOD_AM = 2,
  T41O_DEMOD_NFM = 3,
  T41O_DEMOD_PSK31 = 5,
  T41O_DEMOD_SAM = 8
};

/* Same layout as t41rx_params in include/t41rx.h (checked by a test). */
typedef struct t41o_params {
  int32_t mode;                 /* bands[currentBand].mode                        */
  int32_t f_lo_cut;             /* bands[].FLoCut, Hz                             */
  int32_t f_hi_cut;             /* bands[].FHiCut, Hz                             */
  int32_t nco_freq;             /* NCOFreq, Hz                                    */
  int32_t agc_mode;             /* AGCMode 0=off 1=long 2=slow 3=med 4=fast       */
  int32_t agc_thresh;           /* bands[].AGC_thresh (dB), default 20            */
  int32_t audio_volume;         /* audioVolume 0..100, default 30                 */
  int32_t rf_gain_all_bands;    /* rfGainAllBands (dB), default 1                 */
  int32_t rf_gain;              /* bands[].RFgain start value, default 1          */
  int32_t spectrum_zoom;        /* spectrumZoom index 0..4 (x1..x16), default 1   */
  int32_t current_scale;        /* currentScale 0..4, default 1                   */
  int32_t pixel_offset;         /* bands[].pixel_offset, default 20               */
  int32_t current_nf;           /* currentNoiseFloor[band], default 0             */
  int32_t spectrum_noise_floor; /* spectrumNoiseFloor, default 247                */
  int32_t nfm_filter_bw;        /* nfmFilterBW, default 12000                     */
  int32_t psk31_enable;         /* harness-defined DBPSK+varicode tap */
  float iq_amp_correction;      /* IQAmpCorrectionFactor[band], default 1         */
  float iq_phase_correction;    /* IQPhaseCorrectionFactor[band], default 0       */
  int32_t receive_eq_flag;      /* receiveEQFlag, default 0 (OFF)                 */
  int32_t equalizer_rec[14];    /* EEPROMData.equalizerRec[], default 100 each    */
  int32_t nr_option;            /* nrOptionSelect: 0 off, 1 Kim, 2 spectral, 3 LMS */
  int32_t anr_notch_on;         /* ANR_notchOn, default 0                         */
  int32_t cw_receive;           /* T41State == CW_RECEIVE, default 0              */
  int32_t cw_filter_index;      /* CWFilterIndex 0..5, default 5 (off)            */
  int32_t nb_on;                /* NB_on, default 0                               */
} t41o_params;

typedef struct t41o_debug {
  int32_t agc_state;
  int32_t agc_decay_type;
  int32_t agc_hang_counter;
  int32_t agc_action;
  int32_t rf_gain;
  int32_t codec_timer;
  int32_t zoom_sample_ptr;
  int32_t first_block;
  float agc_volts;
  float agc_ring_max;
  float agc_save_volts;
  float agc_fast_backaverage;
  float agc_hang_backaverage;
  float sam_phzerror;
  float sam_omega2;
  float sam_fil_out;
  float dc_state[2];
  float am_wold;
  double osc_vect_q;
  double osc_vect_i;
} t41o_debug;

/* the mode-dependent tables the control path produces (for table parity tests) */
typedef struct t41o_tables {
  float dec1[28];
  float dec2[46];
  float int1[48];
  float int2[32];
  float mask[1024];
  float am_lp[5];
  float zoom_fir[4];
  /* AGC constants: max_gain, attack_mult, decay_mult, fast_decay_mult, fast_backmult,
     onemfast_backmult, out_target, min_volts, slope_constant, inv_max_input,
     hang_level, hang_backmult, onemhang_backmult, hang_decay_mult, hangtime, fixed_gain */
  float agc[16];
  int32_t attack_buffsize;
  int32_t hang_counter_load;
} t41o_tables;

typedef struct t41o_stream t41o_stream;

void t41o_default_params(t41o_params *p);
/* FLoCut/FHiCut presets of T41/Filter.cpp:341-385 (SetupMode) */
void t41o_mode_default_cuts(int32_t mode, int32_t *f_lo_cut, int32_t *f_hi_cut);

t41o_stream *t41o_create(void);
void t41o_destroy(t41o_stream *s);
int t41o_set_params(t41o_stream *s, const t41o_params *p);
void t41o_get_params(const t41o_stream *s, t41o_params *p);
void t41o_get_tables(const t41o_stream *s, t41o_tables *t);
void t41o_get_debug(const t41o_stream *s, t41o_debug *d);

/*
 * One ProcessIQData() call.  iq: 2048 interleaved (I,Q) floats.  audio: 2048
 * floats (float_buffer_L after the volume scale, before arm_float_to_q15).
 * update_display != 0 plays the role of updateDisplayFlag == 1: the block
 * produces a spectrum row (spec_row = pixelnew[512]) and a waterfall colour row
 * (wf_row[512], RGB565; element 511 is never written by the reference and is 0).
 * psk_bit / psk_char (may be NULL): harness-defined PSK31 tap; *psk_bit is -1
 * when this block carries no symbol decision, else 0/1; *psk_char is the decoded
 * ASCII character or 0.
 */
int t41o_process_block(t41o_stream *s, const float *iq, float *audio, int update_display,
                       int16_t *spec_row, uint16_t *wf_row, int8_t *psk_bit, uint8_t *psk_char);

/* n_blocks consecutive calls; iq [n][2048][2], audio [n][2048]; rows written for
 * every block with (row_mask >> (block % 32)) & 1 ... simplified: row_every > 0
 * makes blocks b with b % row_every == 0 row-producing; rows are packed
 * consecutively. Returns number of rows written. */
int t41o_process(t41o_stream *s, const float *iq, float *audio, int n_blocks, int row_every,
                 int16_t *spec_rows, uint16_t *wf_rows, int8_t *psk_bits, uint8_t *psk_chars);

/* Audio-spectrum + S-meter by-product of row-producing blocks (T41/Process.cpp:550-570 and, for NFM, :791-805):
 * audioYPixel[k], k < AUDIO_SPEC_BOX_W - 2 = 270 (T41/Display.h:6,15,20,42,45), and the running average
 * audioMaxSquaredAve (T41/Process.cpp:32,569).  Where the following t41o_process calls on this receiver put
 * them, one row per row-producing block (NULL = stop capturing). */
#define T41O_AUDIO_SPEC_PIXELS 270
void t41o_capture_audio_spectrum(t41o_stream *s, int32_t *ypixel_rows, float *max_ave_rows);
/* What a row-producing block writes to the PC control app's serial port while controlDataFlag is set (capturing
 * sets it; T41/t41Control.cpp:20-48):
 *   spectrum frame, 518 bytes (T41/FFT.cpp:142-194): "FD" + "%03d" of (255 - max) + 512 data bytes + ';'.  Only
 *     ZoomFFTExe (zoom index != 0) sends it; at zoom x1 nothing is sent and the row is all zeros;
 *   audio-spectrum data, 270 bytes (T41/Process.cpp:819-825): min(audioYPixel[i], 255), no header.
 * One row of each per row-producing block (NULL, NULL = stop, flag cleared). */
#define T41O_SPEC_FRAME_BYTES 518
void t41o_capture_control_frames(t41o_stream *s, uint8_t *spec_frame_rows, uint8_t *audio_frame_rows);
/* S-meter reading the display derives from audioMaxSquaredAve (T41/Display.cpp:976-981, TCVSDR_SMETER build,
 * MyConfigurationFile.h:34): dBm.  Display.cpp is not part of the Tier-A build; this is a restatement only. */
float t41o_smeter_dbm(float audio_max_sq_ave, float gain_correction, int32_t rf_gain, int32_t rf_gain_all_bands);
/* T41/Display.cpp:995-998: length of the S-meter bar (restatement, like t41o_smeter_dbm) */
int32_t t41o_smeter_bar(float dbm);

/* ---- control-path functions exposed for unit tests ---- */
void t41o_calc_fir_coeffs(float *coeffs, int num_coeffs, float fc, float astop, int type, float dfc, float fs);
void t41o_calc_cplx_fir_coeffs(float *ci, float *cq, int num_coeffs, float f_lo, float f_hi, float fs);
float t41o_log10f_fast(float x);
float t41o_approx_atan2(float y, float x);
void t41o_cfft512(float *buf, int inverse);

/* ---- PSK31 primitives (T41/psk31.cpp:235-310) ---- */
typedef struct t41o_psk31 {
  float last_phase;
  uint64_t status_shr;
} t41o_psk31;
void t41o_psk31_reset(t41o_psk31 *st);
uint8_t t41o_psk31_dbpsk_bit(t41o_psk31 *st, float i, float q);
uint8_t t41o_psk31_varicode_push(t41o_psk31 *st, uint8_t symbol);
/* varicode table access: returns bitcount, *code receives the code word; -1 past the end */
int t41o_psk31_varicode_item(int index, uint64_t *code, uint8_t *ascii);

#ifdef __cplusplus
}
#endif
#endif
