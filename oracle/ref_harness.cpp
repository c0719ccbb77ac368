/*
 * oracle/ref_harness.cpp — TEST INFRASTRUCTURE (Tier-A cross-check only).
 *
 * Links the reference's own hot-path translation units (compiled in place from
 * /root/reference by oracle/Makefile) into a host library and drives
 * ProcessIQData() through the same C API shape as the Tier-B oracle.  This file
 * supplies what the firmware's sketch file, display code and hardware would have
 * supplied:
 *   - the globals defined in T41_SDR.ino:129-404, gwv.cpp:15-92, Display.cpp:105-135
 *     (same types and start values),
 *   - the start-up sequence of T41_SDR.ino:473-667 (InitializeDataArrays) and
 *     :753-795 (SoftReset -> AGCPrep, NCOFreq = 0),
 *   - inert link stubs for UI / FT8 / noise-reduction / CW entry points that are
 *     default-off on the receive path,
 *   - the caller's side of the hot path: q15 blocks into the record queues
 *     (R -> I, L -> Q, Process.cpp:107-108), updateDisplayFlag, and the waterfall
 *     row arithmetic of Display.cpp:343-358,459-466 (Display.cpp itself is TFT code
 *     and is not compiled).
 * The reference keeps AGC/NFM/AM state in function-local statics, so one loaded
 * copy of this library is exactly one receiver; tests load a private copy per
 * stream.
 */
#include "SDT.h"
#include "Button.h"
#include "ButtonProc.h"
#include "CW_Excite.h"
#include "CWProcessing.h"
#include "Demod.h"
#include "Display.h"
#include "DSP_Fn.h"
#include "EEPROM.h"
#include "Exciter.h"
#include "FFT.h"
#include "Filter.h"
#include "FIR.h"
#include "Freq_Shift.h"
#include "ft8.h"
#include "InfoBox.h"
#include "Menu.h"
#include "MenuProc.h"
#include "Noise.h"
#include "Process.h"
#include "psk31.h"
#include "Tune.h"
#include "t41Control.h"
#include "Utility.h"

#include "t41_oracle.h"

/* ------------------------------------------------------------------ */
/* shim objects                                                        */
/* ------------------------------------------------------------------ */
SerialStub Serial, SerialUSB1, SerialUSB2;
TwoWire Wire, Wire1;
SPIClass SPI;
SDClass SD;
EEPROMClass EEPROM;
RA8875 tft;
volatile uint32_t t41_shim_scratch_reg;
volatile uint32_t TEMPMON_TEMPSENSE0, TEMPMON_TEMPSENSE1, CCM_ANALOG_PLL_AUDIO, CCM_ANALOG_PLL_AUDIO_NUM,
    CCM_ANALOG_PLL_AUDIO_DENOM, CCM_ANALOG_MISC2, CCM_CSCMR1, CCM_CS1CDR, CCM_CS2CDR, CCM_ANALOG_MISC1,
    IOMUXC_GPR_GPR1, HW_OCOTP_ANA1;
extern "C" { uint32_t t41_shim_f_cpu_actual = 600000000u; }

/* ------------------------------------------------------------------ */
/* globals of T41_SDR.ino / gwv.cpp / Display.cpp / others             */
/* ------------------------------------------------------------------ */
const int SampleRate = 192000;                 /* T41_SDR.ino:129 */
long NCOFreq = 0;
long CWFreqShift = 750;
long calFreqShift = 0;
long TxRxFreq = 0;
long centerFreq = 7048000;
int currentFreqA = 7048000;
int currentBand = 2;                           /* any USB band: 20M */
volatile long fineTuneEncoderMove = 0L;
volatile int menuEncoderMove = 0;
uint8_t T41State = 1;
int xrState = 1;
int radioState = 0, lastState = -1;

struct band bands[NUMBER_OF_BANDS] = {         /* T41_SDR.ino:145-168 (ITU region 2 rows) */
    {3700000, 3500000, 4000000, "80M", DEMOD_LSB, -200, -3000, 1, 0, -2.0, 20, 20},
    {7150000, 7000000, 7300000, "40M", DEMOD_LSB, -200, -3000, 1, 0, -2.0, 20, 20},
    {14200000, 14000000, 14350000, "20M", DEMOD_USB, 3000, 200, 1, 0, 2.0, 20, 20},
    {18100000, 18068000, 18168000, "17M", DEMOD_USB, 3000, 200, 1, 0, 2.0, 20, 20},
    {21200000, 21000000, 21450000, "15M", DEMOD_USB, 3000, 200, 1, 0, 5.0, 20, 20},
    {24920000, 24890000, 24990000, "12M", DEMOD_USB, 3000, 200, 1, 0, 6.0, 20, 20},
    {28350000, 28000000, 29700000, "10M", DEMOD_USB, 3000, 200, 1, 0, 8.5, 20, 20}};

uint32_t FFT_length = FFT_LENGTH;
float32_t float_buffer_L_EX[2048];
float32_t float_buffer_R_EX[2048];
float32_t float_buffer_Temp[2048];
byte sharedRAM1[1024 * 8];
byte sharedRAM2[2048 * 13] __attribute__((aligned(4)));

const arm_cfft_instance_f32 *S;
const arm_cfft_instance_f32 *iS;
const arm_cfft_instance_f32 *maskS;
const arm_cfft_instance_f32 *NR_FFT;
const arm_cfft_instance_f32 *NR_iFFT;
const arm_cfft_instance_f32 *spec_FFT;

arm_biquad_casd_df1_inst_f32 biquad_lowpass1;
arm_biquad_casd_df1_inst_f32 IIR_biquad_Zoom_FFT_I;
arm_biquad_casd_df1_inst_f32 IIR_biquad_Zoom_FFT_Q;
arm_fir_decimate_instance_f32 FIR_dec1_I, FIR_dec1_Q, FIR_dec2_I, FIR_dec2_Q;
arm_fir_decimate_instance_f32 Fir_Zoom_FFT_Decimate_I, Fir_Zoom_FFT_Decimate_Q;
arm_fir_interpolate_instance_f32 FIR_int1_I, FIR_int1_Q, FIR_int2_I, FIR_int2_Q;
arm_lms_norm_instance_f32 LMS_Norm_instance;
arm_lms_instance_f32 LMS_instance;
arm_fir_instance_f32 FIR_Hilbert_L, FIR_Hilbert_R;
/* CWProcessing.cpp:38-49 (that translation unit is the CW decoder and is not built): states and instances of the CW
   audio low-passes ProcessIQData() applies in the CW receive state; coefficients from the reference's FIR.cpp */
extern float32_t CW_AudioFilterCoeffs1[30], CW_AudioFilterCoeffs2[30], CW_AudioFilterCoeffs3[30], CW_AudioFilterCoeffs4[30],
    CW_AudioFilterCoeffs5[30];
float32_t CW_AudioFilter1_state[12], CW_AudioFilter2_state[12], CW_AudioFilter3_state[12], CW_AudioFilter4_state[12],
    CW_AudioFilter5_state[12];
arm_biquad_cascade_df2T_instance_f32 S1_CW_AudioFilter1 = {6, CW_AudioFilter1_state, CW_AudioFilterCoeffs1};
arm_biquad_cascade_df2T_instance_f32 S1_CW_AudioFilter2 = {6, CW_AudioFilter2_state, CW_AudioFilterCoeffs2};
arm_biquad_cascade_df2T_instance_f32 S1_CW_AudioFilter3 = {6, CW_AudioFilter3_state, CW_AudioFilterCoeffs3};
arm_biquad_cascade_df2T_instance_f32 S1_CW_AudioFilter4 = {6, CW_AudioFilter4_state, CW_AudioFilterCoeffs4};
arm_biquad_cascade_df2T_instance_f32 S1_CW_AudioFilter5 = {6, CW_AudioFilter5_state, CW_AudioFilterCoeffs5};

/* T41_SDR.ino:333-345 — same expressions, evaluated once here */
const float32_t DF1 = 4.0;
const float32_t DF2 = 2.0;
const float32_t DF = DF1 * DF2;
const float32_t n_att = 90.0;
static const float32_t n_desired_BW = 9.0;
static const float32_t n_samplerate = 176.0;
static const float32_t n_fpass1 = n_desired_BW / n_samplerate;
static const float32_t n_fpass2 = n_desired_BW / (n_samplerate / DF1);
static const float32_t n_fstop1 = ((n_samplerate / DF1) - n_desired_BW) / n_samplerate;
static const float32_t n_fstop2 = ((n_samplerate / (DF1 * DF2)) - n_desired_BW) / (n_samplerate / DF1);
const uint16_t n_dec1_taps = (1 + (uint16_t)(n_att / (22.0 * (n_fstop1 - n_fpass1))));
const uint16_t n_dec2_taps = (1 + (uint16_t)(n_att / (22.0 * (n_fstop2 - n_fpass2))));
enum { kDec1Taps = 28, kDec2Taps = 46 }; /* values of the two expressions above; checked in t41ref_init */

const uint32_t N_B = FFT_LENGTH / 2 / BUFFER_SIZE * 8;
uint32_t N_BLOCKS = N_B;
float32_t bin_BW = 1.0 / (8.0f * FFT_LENGTH) * 192000;
static float32_t biquad_lowpass1_state[4];
float32_t biquad_lowpass1_coeffs[5] = {0, 0, 0, 0, 0};
float32_t float_buffer_L[BUFFER_SIZE * N_B];
float32_t float_buffer_R[BUFFER_SIZE * N_B];
float32_t iFFT_buffer[FFT_LENGTH * 2 + 1];
static float32_t IIR_biquad_Zoom_FFT_I_state[16];
static float32_t IIR_biquad_Zoom_FFT_Q_state[16];
float temp;

static float32_t FIR_dec1_I_state[kDec1Taps + 2048 - 1], FIR_dec1_Q_state[kDec1Taps + 2048 - 1];
static float32_t FIR_dec2_I_state[kDec2Taps + 512 - 1], FIR_dec2_Q_state[kDec2Taps + 512 - 1];
static float32_t FIR_int1_I_state[24 + 256 - 1], FIR_int1_Q_state[24 + 256 - 1];
static float32_t FIR_int2_I_state[8 + 512 - 1], FIR_int2_Q_state[8 + 512 - 1];
static float32_t Fir_Zoom_FFT_Decimate_I_state[4 + 2048 - 1], Fir_Zoom_FFT_Decimate_Q_state[4 + 2048 - 1];
float32_t FIR_dec1_coeffs[kDec1Taps];
float32_t FIR_dec2_coeffs[kDec2Taps];
float32_t last_sample_buffer_L[BUFFER_SIZE * 2];
float32_t last_sample_buffer_R[BUFFER_SIZE * 2];

AudioRecordQueue Q_in_L, Q_in_R, Q_in_L_Ex, Q_in_R_Ex;
AudioPlayQueue Q_out_L, Q_out_R, Q_out_L_Ex, Q_out_R_Ex;
elapsedMicros usec = 0;

/* gwv.cpp:15-92 */
int AGCMode = 1;
int audioVolume = 30;
int rfGainAllBands = 1;
int spectrumNoiseFloor = 247;
int xmtMode = SSB_MODE;
int nrOptionSelect = 0;
float NR_PSI = 0.0, NR_alpha = 0.95, NR_beta = 0.85;   /* gwv.cpp:61-63 (Noise.cpp's spectral NR, not driven here) */
int currentScale = 1;
long spectrumZoom = 1;
int CWFilterIndex = 5;
float omegaN = 200.0;
float pll_fmax = +4000.0;
float IQAmpCorrectionFactor[NUMBER_OF_BANDS] = {1, 1, 1, 1, 1, 1, 1};
float IQPhaseCorrectionFactor[NUMBER_OF_BANDS] = {0, 0, 0, 0, 0, 0, 0};
int currentNoiseFloor[NUMBER_OF_BANDS] = {0, 0, 0, 0, 0, 0, 0};
int equalizerXmt[14];
config_t EEPROMData;

/* Display.cpp:105-135,223 */
int16_t pixelCurrent[SPECTRUM_RES];
int16_t pixelnew[SPECTRUM_RES];
int16_t pixelold[SPECTRUM_RES];
int updateDisplayFlag = 1;
int wfRows = 0;
int displayScreen = 0;
int currentNF = 0;
dispSc displayScale[] = {{"20 dB/", 10.0, 2, 24, 1.00},
                         {"10 dB/", 20.0, 4, 10, 0.50},
                         {"5 dB/", 40.0, 8, 58, 0.25},
                         {"2 dB/", 100.0, 20, 120, 0.10},
                         {"1 dB/", 200.0, 40, 200, 0.05}};

/* assorted flags owned by UI / FT8 / control files that are not compiled */
bool buttonInterruptsEnabled = false;
int calibrateFlag = -1;                        /* MenuProc.cpp:29 */
bool controlDataFlag = false;
int currentDataMode = 0;
uint8_t keyPressedOn = 0;
bool nfmBWFilterActive = false;                /* ButtonProc.cpp:29 */
int receiveEQFlag = 0;
uint8_t specData[518];
int DSP_Flag = 0, FT_8_counter = 0, ft8State = 0, ft8_decode_flag = 0, ft8_flag = 0, num_decoded_msg = 0;
bool syncFlag = false;
static q15_t ft8_dsp_storage[4096];
q15_t *ft8_dsp_buffer = ft8_dsp_storage;

/* inert link stubs (all default-off on the receive path) */
int ReadSelectedPushButton() { return -1; }
void DoCWReceiveProcessing() {}
void ShowFrequency() {}
void ShowOperatingStats() {}
void ShowSpectrumFreqValues() {}
void DrawBandwidthBar() {}
void ShowBandwidthBarValues() {}
void MyDrawFloat(float, int, int, int, char *) {}
void UpdateInfoBoxItem(uint8_t) {}
void CalibrateOptions() {}
void SetFreq() {}
void process_FT8_FFT() {}
int ft8_decode(void) { return 0; }
void DisplayMessages() {}
void update_synchronization() {}
void auto_sync_FT8() {}
/* the control serial port: what a block writes to it is captured per block (see t41ref_capture_control_frames) */
static uint8_t g_sent_spec[T41O_SPEC_FRAME_BYTES], g_sent_audio[T41O_AUDIO_SPEC_PIXELS];
void T41ControlSendData(uint8_t *data, int len) {
  if (len == T41O_SPEC_FRAME_BYTES) memcpy(g_sent_spec, data, len);               /* FFT.cpp:193 */
  else if (len == T41O_AUDIO_SPEC_PIXELS) memcpy(g_sent_audio, data, len);        /* Process.cpp:824 */
}

/* AGC variables of DSP_Fn.cpp that have external linkage */
extern uint8_t agc_action;
extern uint8_t NB_on;                                   /* Process.cpp:39 */
extern int attack_buffsize;
extern int hang_counter;
extern int out_index;
extern uint32_t in_index;
extern float32_t attack_mult, decay_mult, fast_backmult, fast_decay_mult, fixed_gain, hang_backmult,
    hang_decay_mult, hang_level, hangtime, inv_max_input, max_gain, min_volts, onemfast_backmult,
    onemhang_backmult, slope_constant, out_target;
/* Demod.cpp / Freq_Shift.cpp / Process.cpp / FFT.cpp */
extern float32_t phzerror, fil_out, omega2;
extern double Osc_Vect_Q, Osc_Vect_I;
extern float32_t HP_DC_Butter_state2[2];
extern int8_t first_block;
extern int zoom_sample_ptr;
extern float32_t coefficient_set[];
/* audio-spectrum by-product of row-producing blocks, Process.cpp:32,34,550-570,791-805 */
extern int audioYPixel[];
extern float32_t audioMaxSquaredAve;

static t41o_params g_prm;
static int g_last_set_rf_gain;
static int g_inited = 0;
static unsigned g_psk_block_count = 0;
static uint16_t g_waterfall[SPECTRUM_RES];
static int32_t *g_cap_ypixel = 0;   /* [rows][T41O_AUDIO_SPEC_PIXELS], set by t41ref_capture_audio_spectrum */
static float *g_cap_max_ave = 0;    /* [rows] */
static uint8_t *g_cap_frames = 0;   /* [rows][518] */
static uint8_t *g_cap_audio_frames = 0;   /* [rows][270] */

/* RGB565 ramp of Display.cpp:148-161 comes from the generated data header */
#include "t41_tables_data.h"

extern "C" {

/* mirror of InitializeDataArrays() T41_SDR.ino:473-667 + SoftReset() :753-795 */
int t41ref_init(void) {
  if (g_inited) return 0;
  if (n_dec1_taps != kDec1Taps || n_dec2_taps != kDec2Taps) return -2;
  /* start values of the globals above, in boundary form */
  memset(&g_prm, 0, sizeof(g_prm));
  g_prm.mode = bands[currentBand].mode;
  g_prm.f_lo_cut = bands[currentBand].FLoCut;
  g_prm.f_hi_cut = bands[currentBand].FHiCut;
  g_prm.nco_freq = 0;
  g_prm.agc_mode = AGCMode;
  g_prm.agc_thresh = bands[currentBand].AGC_thresh;
  g_prm.audio_volume = audioVolume;
  g_prm.rf_gain_all_bands = rfGainAllBands;
  g_prm.rf_gain = bands[currentBand].RFgain;
  g_prm.spectrum_zoom = (int32_t)spectrumZoom;
  g_prm.current_scale = currentScale;
  g_prm.pixel_offset = bands[currentBand].pixel_offset;
  g_prm.current_nf = currentNoiseFloor[currentBand];
  g_prm.spectrum_noise_floor = spectrumNoiseFloor;
  g_prm.nfm_filter_bw = nfmFilterBW;
  g_prm.psk31_enable = 0;
  g_prm.iq_amp_correction = IQAmpCorrectionFactor[currentBand];
  g_prm.iq_phase_correction = IQPhaseCorrectionFactor[currentBand];
  g_prm.receive_eq_flag = 0;
  g_prm.nr_option = 0;
  g_prm.anr_notch_on = 0;
  g_prm.cw_receive = 0;
  g_prm.cw_filter_index = 5;
  g_prm.nb_on = 0;
  for (int i = 0; i < 14; i++) {
    EEPROMData.equalizerRec[i] = 100;       /* EEPROM.cpp:59,698 */
    g_prm.equalizer_rec[i] = 100;
  }
  g_last_set_rf_gain = g_prm.rf_gain;

  CalcCplxFIRCoeffs(FIR_Coef_I, FIR_Coef_Q, m_NumTaps, (float32_t)bands[currentBand].FLoCut,
                    (float32_t)bands[currentBand].FHiCut, (float)SampleRate / DF);
  S = &arm_cfft_sR_f32_len512;
  iS = &arm_cfft_sR_f32_len512;
  maskS = &arm_cfft_sR_f32_len512;
  spec_FFT = &arm_cfft_sR_f32_len512;
  NR_FFT = &arm_cfft_sR_f32_len256;
  NR_iFFT = &arm_cfft_sR_f32_len256;
  InitFilterMask();

  biquad_lowpass1.numStages = 1;
  biquad_lowpass1.pCoeffs = biquad_lowpass1_coeffs;
  for (unsigned i = 0; i < 4; i++) biquad_lowpass1_state[i] = 0.0;
  biquad_lowpass1.pState = biquad_lowpass1_state;
  int LP_F_help = bands[currentBand].FHiCut;
  if (LP_F_help < -bands[currentBand].FLoCut) LP_F_help = -bands[currentBand].FLoCut;
  SetIIRCoeffs((float32_t)LP_F_help, 1.3, (float32_t)SampleRate / DF, 0);
  for (int i = 0; i < 5; i++) biquad_lowpass1_coeffs[i] = coefficient_set[i];

  CalcFIRCoeffs(FIR_dec1_coeffs, n_dec1_taps, (float32_t)(n_desired_BW * 1000.0), n_att, 0, 0.0, (float32_t)SampleRate);
  if (arm_fir_decimate_init_f32(&FIR_dec1_I, n_dec1_taps, (uint32_t)DF1, FIR_dec1_coeffs, FIR_dec1_I_state, BUFFER_SIZE * N_BLOCKS)) return -3;
  if (arm_fir_decimate_init_f32(&FIR_dec1_Q, n_dec1_taps, (uint32_t)DF1, FIR_dec1_coeffs, FIR_dec1_Q_state, BUFFER_SIZE * N_BLOCKS)) return -3;
  CalcFIRCoeffs(FIR_dec2_coeffs, n_dec2_taps, (float32_t)(n_desired_BW * 1000.0), n_att, 0, 0.0, (float32_t)(SampleRate / DF1));
  if (arm_fir_decimate_init_f32(&FIR_dec2_I, n_dec2_taps, (uint32_t)DF2, FIR_dec2_coeffs, FIR_dec2_I_state, BUFFER_SIZE * N_BLOCKS / (uint32_t)DF1)) return -3;
  if (arm_fir_decimate_init_f32(&FIR_dec2_Q, n_dec2_taps, (uint32_t)DF2, FIR_dec2_coeffs, FIR_dec2_Q_state, BUFFER_SIZE * N_BLOCKS / (uint32_t)DF1)) return -3;
  CalcFIRCoeffs(FIR_int1_coeffs, 48, (float32_t)(n_desired_BW * 1000.0), n_att, 0, 0.0, SampleRate / 4.0);
  if (arm_fir_interpolate_init_f32(&FIR_int1_I, (uint8_t)DF2, 48, FIR_int1_coeffs, FIR_int1_I_state, BUFFER_SIZE * N_BLOCKS / (uint32_t)DF)) return -3;
  if (arm_fir_interpolate_init_f32(&FIR_int1_Q, (uint8_t)DF2, 48, FIR_int1_coeffs, FIR_int1_Q_state, BUFFER_SIZE * N_BLOCKS / (uint32_t)DF)) return -3;
  CalcFIRCoeffs(FIR_int2_coeffs, 32, (float32_t)(n_desired_BW * 1000.0), n_att, 0, 0.0, (float32_t)SampleRate);
  if (arm_fir_interpolate_init_f32(&FIR_int2_I, (uint8_t)DF1, 32, FIR_int2_coeffs, FIR_int2_I_state, BUFFER_SIZE * N_BLOCKS / (uint32_t)DF1)) return -3;
  if (arm_fir_interpolate_init_f32(&FIR_int2_Q, (uint8_t)DF1, 32, FIR_int2_coeffs, FIR_int2_Q_state, BUFFER_SIZE * N_BLOCKS / (uint32_t)DF1)) return -3;
  SetDecIntFilters();

  float32_t Fstop_Zoom = 0.5 * (float32_t)SampleRate / (1 << spectrumZoom);
  CalcFIRCoeffs(Fir_Zoom_FFT_Decimate_coeffs, 4, Fstop_Zoom, 60, 0, 0.0, (float32_t)SampleRate);
  if (arm_fir_decimate_init_f32(&Fir_Zoom_FFT_Decimate_I, 4, 128, Fir_Zoom_FFT_Decimate_coeffs, Fir_Zoom_FFT_Decimate_I_state, BUFFER_SIZE * N_BLOCKS)) return -3;
  if (arm_fir_decimate_init_f32(&Fir_Zoom_FFT_Decimate_Q, 4, 128, Fir_Zoom_FFT_Decimate_coeffs, Fir_Zoom_FFT_Decimate_Q_state, BUFFER_SIZE * N_BLOCKS)) return -3;
  IIR_biquad_Zoom_FFT_I.numStages = 4;
  IIR_biquad_Zoom_FFT_Q.numStages = 4;
  for (unsigned i = 0; i < 16; i++) {
    IIR_biquad_Zoom_FFT_I_state[i] = 0.0;
    IIR_biquad_Zoom_FFT_Q_state[i] = 0.0;
  }
  IIR_biquad_Zoom_FFT_I.pState = IIR_biquad_Zoom_FFT_I_state;
  IIR_biquad_Zoom_FFT_Q.pState = IIR_biquad_Zoom_FFT_Q_state;
  IIR_biquad_Zoom_FFT_I.pCoeffs = mag_coeffs[spectrumZoom];
  IIR_biquad_Zoom_FFT_Q.pCoeffs = mag_coeffs[spectrumZoom];
  ZoomFFTPrep();

  AGCPrep();      /* SoftReset, T41_SDR.ino:791 */
  NCOFreq = 0;    /* T41_SDR.ino:793 */
  CalcFilters();  /* the receive-state entry in loop() runs SetupMode()/CalcFilters() before the first block */
  g_inited = 1;
  return 0;
}

/* what ChangeDemodMode/SetupMode (ButtonProc.cpp:237, Filter.cpp:341), SetBWFilters
   (Encoders.cpp:50), AGCOptions (MenuProc.cpp:275) and SetZoom (Display.cpp:1402) do
   to the hot path's inputs */
int t41ref_set_params(const t41o_params *p) {
  if (!g_inited && t41ref_init()) return -2;
  const t41o_params old = g_prm;
  g_prm = *p;
  bands[currentBand].mode = p->mode;
  bands[currentBand].FLoCut = p->f_lo_cut;
  bands[currentBand].FHiCut = p->f_hi_cut;
  bands[currentBand].AGC_thresh = p->agc_thresh;
  bands[currentBand].pixel_offset = (int16_t)p->pixel_offset;
  if (p->rf_gain != g_last_set_rf_gain) {
    g_last_set_rf_gain = p->rf_gain;
    bands[currentBand].RFgain = p->rf_gain;
  }
  NCOFreq = p->nco_freq;
  audioVolume = p->audio_volume;
  rfGainAllBands = p->rf_gain_all_bands;
  currentScale = p->current_scale;
  currentNoiseFloor[currentBand] = p->current_nf;
  currentNF = p->current_nf;
  spectrumNoiseFloor = p->spectrum_noise_floor;
  nfmFilterBW = p->nfm_filter_bw;
  IQAmpCorrectionFactor[currentBand] = p->iq_amp_correction;
  IQPhaseCorrectionFactor[currentBand] = p->iq_phase_correction;
  receiveEQFlag = p->receive_eq_flag;
  if (p->nr_option < 0 || p->nr_option > 3) return -1;
  nrOptionSelect = p->nr_option;
  ANR_notchOn = (uint8_t)p->anr_notch_on;
  NB_on = (uint8_t)(p->nb_on != 0);                    /* Process.cpp:39 */
  if (p->cw_filter_index < 0 || p->cw_filter_index > 5) return -1;
  T41State = p->cw_receive == 1 ? CW_RECEIVE : 1;      /* 1: the state the harness otherwise sits in */
  CWFilterIndex = p->cw_filter_index;
  for (int i = 0; i < 14; i++) EEPROMData.equalizerRec[i] = p->equalizer_rec[i];
  if (p->mode != old.mode || p->f_lo_cut != old.f_lo_cut || p->f_hi_cut != old.f_hi_cut) CalcFilters();
  if (p->agc_mode != old.agc_mode || p->agc_thresh != old.agc_thresh) {
    AGCMode = p->agc_mode;
    AGCLoadValues();
  }
  if (p->spectrum_zoom != old.spectrum_zoom) {
    spectrumZoom = p->spectrum_zoom;
    ZoomFFTPrep();
  }
  return 0;
}

int t41ref_process_block(const float *iq, float *audio, int update_display, int16_t *spec_row,
                         uint16_t *wf_row, int8_t *psk_bit, uint8_t *psk_char) {
  if (!g_inited && t41ref_init()) return -2;
  /* the chain starts from q15 samples: the float input must sit on the q15 grid */
  int16_t qi[2048], qq[2048];
  for (int i = 0; i < 2048; i++) {
    float a = iq[2 * i] * 32768.0f, b = iq[2 * i + 1] * 32768.0f;
    if (a != (float)(int)a || b != (float)(int)b || a < -32768.0f || a > 32767.0f || b < -32768.0f || b > 32767.0f) return -4;
    qi[i] = (int16_t)a;
    qq[i] = (int16_t)b;
  }
  Q_in_L.clear();
  Q_in_R.clear();
  for (unsigned blk = 0; blk < 16; blk++) {
    Q_in_R.push(qi + 128 * blk); /* right channel carries I (B1) */
    Q_in_L.push(qq + 128 * blk);
  }
  static const int16_t silence[128] = {0};
  Q_in_R.push(silence);          /* ProcessIQData wants MORE than N_BLOCKS queued (Process.cpp:93) */
  Q_in_L.push(silence);
  updateDisplayFlag = update_display ? 1 : 0;
  memset(g_sent_spec, 0, sizeof(g_sent_spec));
  memset(g_sent_audio, 0, sizeof(g_sent_audio));
  ProcessIQData();
  Q_in_L.clear();
  Q_in_R.clear();
  memcpy(audio, float_buffer_L, 2048 * sizeof(float));

  int8_t bit_out = -1;
  uint8_t char_out = 0;
  if (g_prm.psk31_enable && g_prm.mode != DEMOD_NFM && g_prm.mode != DEMOD_PSK31) {
    if (g_psk_block_count % 3u == 0u) {
      float pair[2] = {iFFT_buffer[FFT_LENGTH], iFFT_buffer[FFT_LENGTH + 1]};
      uint8_t bit = 0;
      dbpsk_decoder_c_u8(pair, &bit, 1);       /* psk31.cpp:293-310 on one proper (I,Q) pair */
      bit_out = (int8_t)bit;
      char_out = (uint8_t)psk31_varicode_decoder_push(bit);  /* psk31.cpp:235-264 */
    }
    g_psk_block_count++;
  }
  if (psk_bit) *psk_bit = bit_out;
  if (psk_char) *psk_char = char_out;

  if (update_display) {
    /* Display.cpp:343-358,459-466 for x1 = 0..510 */
    for (int x1 = 0; x1 < SPECTRUM_RES - 1; x1++) {
      int y_new_plot = spectrumNoiseFloor - pixelnew[x1] - currentNF;
      if (y_new_plot > SPECTRUM_BOTTOM) y_new_plot = SPECTRUM_BOTTOM;
      if (y_new_plot < SPECTRUM_TOP_Y) y_new_plot = SPECTRUM_TOP_Y;
      int test1 = -y_new_plot + 230;
      if (test1 < 0) test1 = 0;
      if (test1 > 116) test1 = 116;
      g_waterfall[x1] = t41o_gradient[test1];
    }
    if (spec_row) memcpy(spec_row, pixelnew, sizeof(pixelnew));
    if (wf_row) memcpy(wf_row, g_waterfall, sizeof(g_waterfall));
  }
  return 0;
}

int t41ref_process(const float *iq, float *audio, int n_blocks, int row_every, int16_t *spec_rows,
                   uint16_t *wf_rows, int8_t *psk_bits, uint8_t *psk_chars) {
  int rows = 0;
  for (int b = 0; b < n_blocks; b++) {
    int upd = (row_every > 0) && (b % row_every == 0);
    int rc = t41ref_process_block(iq + (size_t)b * 4096, audio + (size_t)b * 2048, upd,
                                  (upd && spec_rows) ? spec_rows + (size_t)rows * 512 : 0,
                                  (upd && wf_rows) ? wf_rows + (size_t)rows * 512 : 0,
                                  psk_bits ? psk_bits + b : 0, psk_chars ? psk_chars + b : 0);
    if (rc) return rc;
    if (upd && g_cap_ypixel) {
      for (int k = 0; k < T41O_AUDIO_SPEC_PIXELS; k++) g_cap_ypixel[(size_t)rows * T41O_AUDIO_SPEC_PIXELS + k] = audioYPixel[k];
    }
    if (upd && g_cap_max_ave) g_cap_max_ave[rows] = audioMaxSquaredAve;
    if (upd && g_cap_frames) memcpy(g_cap_frames + (size_t)rows * T41O_SPEC_FRAME_BYTES, g_sent_spec, T41O_SPEC_FRAME_BYTES);
    if (upd && g_cap_audio_frames) memcpy(g_cap_audio_frames + (size_t)rows * T41O_AUDIO_SPEC_PIXELS, g_sent_audio, T41O_AUDIO_SPEC_PIXELS);
    rows += upd;
  }
  return rows;
}

/* where the following t41ref_process calls put audioYPixel[0..269] and audioMaxSquaredAve after every
   row-producing block (NULL = stop capturing) */
void t41ref_capture_audio_spectrum(int32_t *ypixel_rows, float *max_ave_rows) {
  g_cap_ypixel = ypixel_rows;
  g_cap_max_ave = max_ave_rows;
}

/* capture what every row-producing block sends to the control serial port (all zeros when it sends nothing):
   the 518-byte spectrum frame and the 270 audio-spectrum bytes; sets controlDataFlag like the control app does */
void t41ref_capture_control_frames(uint8_t *spec_frame_rows, uint8_t *audio_frame_rows) {
  g_cap_frames = spec_frame_rows;
  g_cap_audio_frames = audio_frame_rows;
  controlDataFlag = (spec_frame_rows != 0) || (audio_frame_rows != 0);
}

/* the reference's WAV reader on a host file (SD.open is backed by stdio in the shim) */
int t41ref_load_wav(const char *path, uint32_t num_samples) { return load_wav(path, num_samples); }
int t41ref_read_wave(float *buf, int size_buf) { return readWave(buf, size_buf) ? 1 : 0; }

void t41ref_get_params(t41o_params *p) {
  if (!g_inited) t41ref_init();
  *p = g_prm;
}

void t41ref_get_tables(t41o_tables *t) {
  memset(t, 0, sizeof(*t));
  memcpy(t->dec1, FIR_dec1_coeffs, sizeof(t->dec1));
  memcpy(t->dec2, FIR_dec2_coeffs, sizeof(t->dec2));
  memcpy(t->int1, FIR_int1_coeffs, sizeof(t->int1));
  memcpy(t->int2, FIR_int2_coeffs, sizeof(t->int2));
  memcpy(t->mask, FIR_filter_mask, sizeof(t->mask));
  memcpy(t->am_lp, biquad_lowpass1_coeffs, sizeof(t->am_lp));
  memcpy(t->zoom_fir, Fir_Zoom_FFT_Decimate_coeffs, sizeof(t->zoom_fir));
  const float v[16] = {max_gain, attack_mult, decay_mult, fast_decay_mult, fast_backmult, onemfast_backmult,
                       out_target, min_volts, slope_constant, inv_max_input, hang_level, hang_backmult,
                       onemhang_backmult, hang_decay_mult, hangtime, fixed_gain};
  memcpy(t->agc, v, sizeof(v));
  t->attack_buffsize = attack_buffsize;
  t->hang_counter_load = (int)(hangtime * SampleRate / DF);
}

/* only what has external linkage in the reference; the AGC's function statics
   (state, volts, ring_max, ...) are reported as -1 / NaN */
void t41ref_get_debug(t41o_debug *d) {
  memset(d, 0, sizeof(*d));
  d->agc_state = -1;
  d->agc_decay_type = -1;
  d->agc_hang_counter = hang_counter;
  d->agc_action = agc_action;
  d->rf_gain = bands[currentBand].RFgain;
  d->codec_timer = -1;
  d->zoom_sample_ptr = zoom_sample_ptr;
  d->first_block = first_block;
  d->agc_volts = NAN;
  d->agc_ring_max = NAN;
  d->agc_save_volts = NAN;
  d->agc_fast_backaverage = NAN;
  d->agc_hang_backaverage = NAN;
  d->sam_phzerror = phzerror;
  d->sam_omega2 = omega2;
  d->sam_fil_out = fil_out;
  d->dc_state[0] = HP_DC_Butter_state2[0];
  d->dc_state[1] = HP_DC_Butter_state2[1];
  d->am_wold = NAN;
  d->osc_vect_q = Osc_Vect_Q;
  d->osc_vect_i = Osc_Vect_I;
}

/* control-path and helper functions of the reference, for unit-level cross-checks */
void t41ref_calc_fir_coeffs(float *coeffs, int num_coeffs, float fc, float astop, int type, float dfc, float fs) {
  CalcFIRCoeffs(coeffs, num_coeffs, fc, astop, type, dfc, fs);
}
void t41ref_calc_cplx_fir_coeffs(float *ci, float *cq, int num_coeffs, float f_lo, float f_hi, float fs) {
  CalcCplxFIRCoeffs(ci, cq, num_coeffs, f_lo, f_hi, fs);
}
float t41ref_log10f_fast(float x) { return log10f_fast(x); }
float t41ref_approx_atan2(float y, float x) { return ApproxAtan2(y, x); }
float t41ref_alpha_beta_mag(float i, float q) { return AlphaBetaMag(i, q); }

}  // extern "C"
