/*
 * oracle/cmsis_port.h — TEST INFRASTRUCTURE, not product code.
 *
 * Portable C restatement of the ARM CMSIS-DSP primitives that the T41 receive
 * chain calls (call sites: SURVEY.md §2.3; reference software/T41_SDR/Process.cpp,
 * FFT.cpp, Filter.cpp, Demod.cpp).  CMSIS-DSP is a third-party dependency of the
 * reference that is NOT vendored under /root/reference and has no pinned version
 * there (Teensyduino ships a prebuilt libarm_cortexM7lfsp_math of the CMSIS 4.5 /
 * DSP 1.4.x era).  The functions below restate the published CMSIS-DSP algorithms
 * (names, argument order, state layouts and operation order) in our own code.
 *
 * Rounding model (the oracle's definition, replicated op-for-op by the CUDA path):
 *   - every arithmetic operation is a separately rounded IEEE-754 binary32
 *     operation (build with -ffp-contract=off), EXCEPT
 *   - multiply-accumulate loops of the FIR family (arm_fir_f32,
 *     arm_fir_decimate_f32, arm_fir_interpolate_f32, arm_dot_prod_f32,
 *     arm_power_f32) accumulate with one fused multiply-add per tap in tap order,
 *     i.e. acc = fmaf(x, c, acc).  CMSIS writes these as `sum0 += x * c;`, which
 *     GCC for the Cortex-M7 (FPv5, -ffp-contract=fast default) emits as VFMA.F32.
 *
 * The same header doubles as the `arm_math.h` seen by the reference translation
 * units when they are compiled in place for the Tier-A cross-check
 * (oracle/ref_shim/arm_math.h includes it).
 */
#ifndef T41_ORACLE_CMSIS_PORT_H
#define T41_ORACLE_CMSIS_PORT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float float32_t;
typedef double float64_t;
typedef int16_t q15_t;
typedef int32_t q31_t;

typedef enum {
  ARM_MATH_SUCCESS = 0,
  ARM_MATH_ARGUMENT_ERROR = -1,
  ARM_MATH_LENGTH_ERROR = -2
} arm_status;

/* ---- instance structures (field order as in CMSIS-DSP arm_math.h) ---- */
typedef struct {
  uint16_t fftLen;
  const float32_t *pTwiddle;
  const uint16_t *pBitRevTable;
  uint16_t bitRevLength;
} arm_cfft_instance_f32;

typedef struct {
  uint32_t numStages;
  float32_t *pState;  /* 4 per stage: x[n-1], x[n-2], y[n-1], y[n-2] */
  const float32_t *pCoeffs; /* 5 per stage: b0 b1 b2 a1 a2 (a's pre-negated) */
} arm_biquad_casd_df1_inst_f32;

typedef struct {
  uint8_t numStages;
  float32_t *pState;  /* 2 per stage: d1, d2 */
  const float32_t *pCoeffs;
} arm_biquad_cascade_df2T_instance_f32;

typedef struct {
  uint8_t M;
  uint16_t numTaps;
  const float32_t *pCoeffs;
  float32_t *pState; /* numTaps + blockSize - 1 */
} arm_fir_decimate_instance_f32;

typedef struct {
  uint8_t L;
  uint16_t phaseLength;
  const float32_t *pCoeffs;
  float32_t *pState; /* phaseLength + blockSize - 1 */
} arm_fir_interpolate_instance_f32;

typedef struct {
  uint16_t numTaps;
  float32_t *pState; /* numTaps + blockSize - 1 */
  const float32_t *pCoeffs;
} arm_fir_instance_f32;

/* only declared so that the reference's extern declarations parse */
typedef struct {
  uint16_t numTaps;
  float32_t *pState;
  float32_t *pCoeffs;
  float32_t mu;
} arm_lms_instance_f32;

typedef struct {
  uint16_t numTaps;
  float32_t *pState;
  float32_t *pCoeffs;
  float32_t mu;
  float32_t energy;
  float32_t x0;
} arm_lms_norm_instance_f32;

/* only so that the reference's Noise.cpp links (its LMS-norm set-up is not on the receive path) */
void arm_fill_f32(float32_t value, float32_t *pDst, uint32_t blockSize);
void arm_lms_norm_init_f32(arm_lms_norm_instance_f32 *S, uint16_t numTaps, float32_t *pCoeffs, float32_t *pState,
                           float32_t mu, uint32_t blockSize);

extern const arm_cfft_instance_f32 arm_cfft_sR_f32_len256;
extern const arm_cfft_instance_f32 arm_cfft_sR_f32_len512;
extern const arm_cfft_instance_f32 arm_cfft_sR_f32_len1024;
extern const arm_cfft_instance_f32 arm_cfft_sR_f32_len2048;

/* ---- elementwise / reductions ---- */
void arm_q15_to_float(const q15_t *pSrc, float32_t *pDst, uint32_t n);
void arm_float_to_q15(const float32_t *pSrc, q15_t *pDst, uint32_t n);
void arm_scale_f32(const float32_t *pSrc, float32_t scale, float32_t *pDst, uint32_t n);
void arm_add_f32(const float32_t *a, const float32_t *b, float32_t *pDst, uint32_t n);
void arm_mult_f32(const float32_t *a, const float32_t *b, float32_t *pDst, uint32_t n);
void arm_negate_f32(const float32_t *pSrc, float32_t *pDst, uint32_t n);
void arm_copy_f32(const float32_t *pSrc, float32_t *pDst, uint32_t n);
void arm_cmplx_mult_cmplx_f32(const float32_t *a, const float32_t *b, float32_t *pDst, uint32_t numSamples);
void arm_dot_prod_f32(const float32_t *a, const float32_t *b, uint32_t n, float32_t *result);
void arm_power_f32(const float32_t *pSrc, uint32_t n, float32_t *result);
void arm_var_f32(const float32_t *pSrc, uint32_t n, float32_t *result);
void arm_max_f32(const float32_t *pSrc, uint32_t n, float32_t *pResult, uint32_t *pIndex);

/* ---- table trig ---- */
float32_t arm_sin_f32(float32_t x);
float32_t arm_cos_f32(float32_t x);
/* the 513-entry table the two functions above interpolate in */
const float32_t *t41_cmsis_sin_table(void);

/* ---- filters ---- */
arm_status arm_fir_decimate_init_f32(arm_fir_decimate_instance_f32 *S, uint16_t numTaps, uint8_t M,
                                     const float32_t *pCoeffs, float32_t *pState, uint32_t blockSize);
void arm_fir_decimate_f32(const arm_fir_decimate_instance_f32 *S, const float32_t *pSrc,
                          float32_t *pDst, uint32_t blockSize);
arm_status arm_fir_interpolate_init_f32(arm_fir_interpolate_instance_f32 *S, uint8_t L, uint16_t numTaps,
                                        const float32_t *pCoeffs, float32_t *pState, uint32_t blockSize);
void arm_fir_interpolate_f32(const arm_fir_interpolate_instance_f32 *S, const float32_t *pSrc,
                             float32_t *pDst, uint32_t blockSize);
void arm_fir_init_f32(arm_fir_instance_f32 *S, uint16_t numTaps, const float32_t *pCoeffs,
                      float32_t *pState, uint32_t blockSize);
void arm_fir_f32(const arm_fir_instance_f32 *S, const float32_t *pSrc, float32_t *pDst, uint32_t blockSize);
void arm_biquad_cascade_df1_init_f32(arm_biquad_casd_df1_inst_f32 *S, uint8_t numStages,
                                     const float32_t *pCoeffs, float32_t *pState);
void arm_biquad_cascade_df1_f32(const arm_biquad_casd_df1_inst_f32 *S, const float32_t *pSrc,
                                float32_t *pDst, uint32_t blockSize);
void arm_biquad_cascade_df2T_init_f32(arm_biquad_cascade_df2T_instance_f32 *S, uint8_t numStages,
                                      const float32_t *pCoeffs, float32_t *pState);
void arm_biquad_cascade_df2T_f32(const arm_biquad_cascade_df2T_instance_f32 *S, const float32_t *pSrc,
                                 float32_t *pDst, uint32_t blockSize);

/* ---- complex FFT (in place, interleaved re/im) ---- */
void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag);

#ifdef __cplusplus
}
#endif
#endif
