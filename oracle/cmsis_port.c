/*
 * oracle/cmsis_port.c — TEST INFRASTRUCTURE, not product code.
 * See cmsis_port.h for the rounding model.  Build with -ffp-contract=off.
 *
 * Each function restates the CMSIS-DSP algorithm of the same name as used by the
 * reference receive chain (call sites cited per function; T41/ =
 * /root/reference/software/T41_SDR/).
 */
#include "cmsis_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* elementwise                                                        */
/* ------------------------------------------------------------------ */

/* T41/Process.cpp:107-108 — q15 sample / 32768 */
void arm_q15_to_float(const q15_t *pSrc, float32_t *pDst, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i) pDst[i] = (float32_t)pSrc[i] / 32768.0f;
}

/* T41/Process.cpp:936 — saturate((q31)(x * 32768)), truncation toward zero */
void arm_float_to_q15(const float32_t *pSrc, q15_t *pDst, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i) {
    float32_t v = pSrc[i] * 32768.0f;
    int32_t q;
    if (v != v) q = 0;                             /* NaN: Cortex-M7 VCVT.S32.F32 converts NaN to 0 (not the x86 INT32_MIN) */
    else if (!(v > -2147483648.0f)) q = INT32_MIN;
    else if (v >= 2147483648.0f) q = INT32_MAX;
    else q = (int32_t)v;
    if (q > 32767) q = 32767;
    if (q < -32768) q = -32768;
    pDst[i] = (q15_t)q;
  }
}

/* T41/Process.cpp:118-119,133-134,166,491-492,929 */
void arm_scale_f32(const float32_t *pSrc, float32_t scale, float32_t *pDst, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i) pDst[i] = pSrc[i] * scale;
}

/* T41/Utility.cpp:182,185 */
void arm_add_f32(const float32_t *a, const float32_t *b, float32_t *pDst, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i) pDst[i] = a[i] + b[i];
}

void arm_mult_f32(const float32_t *a, const float32_t *b, float32_t *pDst, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i) pDst[i] = a[i] * b[i];
}

void arm_negate_f32(const float32_t *pSrc, float32_t *pDst, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i) pDst[i] = -pSrc[i];
}

/* T41/Process.cpp:706 */
void arm_copy_f32(const float32_t *pSrc, float32_t *pDst, uint32_t n) {
  if (pSrc != pDst) memmove(pDst, pSrc, (size_t)n * sizeof(float32_t));
}

/* T41/Process.cpp:547,788 — (a+jb)(c+jd) = (ac - bd) + j(ad + bc) */
void arm_cmplx_mult_cmplx_f32(const float32_t *a, const float32_t *b, float32_t *pDst, uint32_t numSamples) {
  for (uint32_t i = 0; i < numSamples; ++i) {
    float32_t ar = a[2 * i], ai = a[2 * i + 1];
    float32_t br = b[2 * i], bi = b[2 * i + 1];
    float32_t rr = ar * br, ii = ai * bi, ri = ar * bi, ir = ai * br;
    pDst[2 * i] = rr - ii;
    pDst[2 * i + 1] = ri + ir;
  }
}

void arm_dot_prod_f32(const float32_t *a, const float32_t *b, uint32_t n, float32_t *result) {
  float32_t acc = 0.0f;
  for (uint32_t i = 0; i < n; ++i) acc = fmaf(a[i], b[i], acc);
  *result = acc;
}

void arm_power_f32(const float32_t *pSrc, uint32_t n, float32_t *result) {
  float32_t acc = 0.0f;
  for (uint32_t i = 0; i < n; ++i) acc = fmaf(pSrc[i], pSrc[i], acc);
  *result = acc;
}

/* unbiased variance: (sum(x^2) - sum(x)^2 / n) / (n - 1) */
void arm_var_f32(const float32_t *pSrc, uint32_t n, float32_t *result) {
  if (n <= 1u) { *result = 0.0f; return; }
  float32_t sum = 0.0f, sumsq = 0.0f;
  for (uint32_t i = 0; i < n; ++i) { sum += pSrc[i]; sumsq = fmaf(pSrc[i], pSrc[i], sumsq); }
  float32_t msq = sumsq / (float32_t)(n - 1u);
  float32_t sqm = (sum * sum) / ((float32_t)n * (float32_t)(n - 1u));
  *result = msq - sqm;
}

/* T41/Process.cpp:568,803 — maximum and the index of its first occurrence */
void arm_fill_f32(float32_t value, float32_t *pDst, uint32_t blockSize) {
  for (uint32_t i = 0; i < blockSize; ++i) pDst[i] = value;
}

void arm_lms_norm_init_f32(arm_lms_norm_instance_f32 *S, uint16_t numTaps, float32_t *pCoeffs, float32_t *pState,
                           float32_t mu, uint32_t blockSize) {
  S->numTaps = numTaps;
  S->pCoeffs = pCoeffs;
  memset(pState, 0, ((size_t)numTaps + (blockSize - 1u)) * sizeof(float32_t));
  S->pState = pState;
  S->mu = mu;
  S->energy = 0.0f;
  S->x0 = 0.0f;
}

void arm_max_f32(const float32_t *pSrc, uint32_t n, float32_t *pResult, uint32_t *pIndex) {
  float32_t best = pSrc[0];
  uint32_t where = 0;
  for (uint32_t i = 1; i < n; ++i) {
    if (best < pSrc[i]) { best = pSrc[i]; where = i; }
  }
  *pResult = best;
  *pIndex = where;
}

/* ------------------------------------------------------------------ */
/* table sine / cosine (T41/Demod.cpp:75-76)                           */
/* ------------------------------------------------------------------ */
#define T41_SIN_TABLE_SIZE 512

static float32_t g_sin_table[T41_SIN_TABLE_SIZE + 1];
static int g_sin_table_ready = 0;

const float32_t *t41_cmsis_sin_table(void) {
  if (!g_sin_table_ready) {
    for (int k = 0; k <= T41_SIN_TABLE_SIZE; ++k) {
      g_sin_table[k] = (float32_t)sin(2.0 * 3.14159265358979323846 * (double)k / (double)T41_SIN_TABLE_SIZE);
    }
    g_sin_table[0] = 0.0f;
    g_sin_table[T41_SIN_TABLE_SIZE / 2] = 0.0f;
    g_sin_table[T41_SIN_TABLE_SIZE] = 0.0f;
    g_sin_table_ready = 1;
  }
  return g_sin_table;
}

/* shared tail: `in` is the angle in turns (x / 2pi [+ 0.25 for cosine]) */
static float32_t table_lookup_turns(float32_t in) {
  const float32_t *tab = t41_cmsis_sin_table();
  int32_t n = (int32_t)in;
  if (in < 0.0f) n--;
  in = in - (float32_t)n;                     /* fractional turn, [0,1] */
  float32_t findex = (float32_t)T41_SIN_TABLE_SIZE * in;
  uint16_t index = (uint16_t)findex;
  if (index >= T41_SIN_TABLE_SIZE) {          /* in rounded up to 1.0 */
    index = 0;
    findex -= (float32_t)T41_SIN_TABLE_SIZE;
  }
  float32_t fract = findex - (float32_t)index;
  float32_t a = tab[index];
  float32_t b = tab[index + 1];
  float32_t wa = (1.0f - fract) * a;
  float32_t wb = fract * b;
  return wa + wb;
}

float32_t arm_sin_f32(float32_t x) {
  float32_t in = x * 0.159154943092f;
  return table_lookup_turns(in);
}

float32_t arm_cos_f32(float32_t x) {
  float32_t in = x * 0.159154943092f + 0.25f;
  return table_lookup_turns(in);
}

/* ------------------------------------------------------------------ */
/* FIR family                                                          */
/* ------------------------------------------------------------------ */

/* T41/T41_SDR.ino:574-590 */
arm_status arm_fir_decimate_init_f32(arm_fir_decimate_instance_f32 *S, uint16_t numTaps, uint8_t M,
                                     const float32_t *pCoeffs, float32_t *pState, uint32_t blockSize) {
  if ((blockSize % M) != 0u) return ARM_MATH_LENGTH_ERROR;
  S->numTaps = numTaps;
  S->M = M;
  S->pCoeffs = pCoeffs;
  S->pState = pState;
  memset(pState, 0, ((size_t)numTaps + blockSize - 1u) * sizeof(float32_t));
  return ARM_MATH_SUCCESS;
}

/*
 * T41/Process.cpp:262-267,378-386,474-479; T41/FFT.cpp:87-88.
 * State = [numTaps-1 oldest-first history | incoming].  Each output appends M new
 * inputs and correlates the window that starts numTaps-1+M(m) .. with the
 * coefficients in stored order: y[m] = sum_i c[i] * w[m*M + i].
 * In-place (pSrc == pDst) is supported, as the reference uses it.
 */
void arm_fir_decimate_f32(const arm_fir_decimate_instance_f32 *S, const float32_t *pSrc,
                          float32_t *pDst, uint32_t blockSize) {
  const uint32_t numTaps = S->numTaps;
  const uint32_t M = S->M;
  float32_t *state = S->pState;
  float32_t *fill = state + (numTaps - 1u);
  const uint32_t outCount = blockSize / M;
  uint32_t base = 0;
  for (uint32_t m = 0; m < outCount; ++m) {
    for (uint32_t k = 0; k < M; ++k) *fill++ = *pSrc++;
    const float32_t *w = state + base;
    float32_t acc = 0.0f;
    for (uint32_t i = 0; i < numTaps; ++i) acc = fmaf(w[i], S->pCoeffs[i], acc);
    *pDst++ = acc;
    base += M;
  }
  /* keep the last numTaps-1 samples as next call's history */
  memmove(state, state + base, (size_t)(numTaps - 1u) * sizeof(float32_t));
}

/* T41/T41_SDR.ino:595-616 */
arm_status arm_fir_interpolate_init_f32(arm_fir_interpolate_instance_f32 *S, uint8_t L, uint16_t numTaps,
                                        const float32_t *pCoeffs, float32_t *pState, uint32_t blockSize) {
  if ((numTaps % L) != 0u) return ARM_MATH_LENGTH_ERROR;
  S->L = L;
  S->phaseLength = (uint16_t)(numTaps / L);
  S->pCoeffs = pCoeffs;
  S->pState = pState;
  memset(pState, 0, ((size_t)S->phaseLength + blockSize - 1u) * sizeof(float32_t));
  return ARM_MATH_SUCCESS;
}

/*
 * T41/Process.cpp:917,920.  Polyphase: for every input, L outputs; output phase p
 * uses coefficients c[(L-1-p) + k*L], k = 0..phaseLength-1, against the
 * oldest-first window of phaseLength samples ending at the new input.  No xL gain.
 */
void arm_fir_interpolate_f32(const arm_fir_interpolate_instance_f32 *S, const float32_t *pSrc,
                             float32_t *pDst, uint32_t blockSize) {
  const uint32_t L = S->L;
  const uint32_t P = S->phaseLength;
  float32_t *state = S->pState;
  float32_t *fill = state + (P - 1u);
  for (uint32_t n = 0; n < blockSize; ++n) {
    *fill++ = pSrc[n];
    const float32_t *w = state + n;
    for (uint32_t p = 0; p < L; ++p) {
      const float32_t *c = S->pCoeffs + (L - 1u - p);
      float32_t acc = 0.0f;
      for (uint32_t k = 0; k < P; ++k) acc = fmaf(w[k], c[k * L], acc);
      *pDst++ = acc;
    }
  }
  memmove(state, state + blockSize, (size_t)(P - 1u) * sizeof(float32_t));
}

void arm_fir_init_f32(arm_fir_instance_f32 *S, uint16_t numTaps, const float32_t *pCoeffs,
                      float32_t *pState, uint32_t blockSize) {
  S->numTaps = numTaps;
  S->pCoeffs = pCoeffs;
  S->pState = pState;
  memset(pState, 0, ((size_t)numTaps + blockSize - 1u) * sizeof(float32_t));
}

/* y[n] = sum_i c[i] * w[n + i], window oldest-first (coefficients stored time-reversed) */
void arm_fir_f32(const arm_fir_instance_f32 *S, const float32_t *pSrc, float32_t *pDst, uint32_t blockSize) {
  const uint32_t numTaps = S->numTaps;
  float32_t *state = S->pState;
  float32_t *fill = state + (numTaps - 1u);
  for (uint32_t n = 0; n < blockSize; ++n) {
    *fill++ = pSrc[n];
    const float32_t *w = state + n;
    float32_t acc = 0.0f;
    for (uint32_t i = 0; i < numTaps; ++i) acc = fmaf(w[i], S->pCoeffs[i], acc);
    pDst[n] = acc;
  }
  memmove(state, state + blockSize, (size_t)(numTaps - 1u) * sizeof(float32_t));
}

/* ------------------------------------------------------------------ */
/* biquads                                                             */
/* ------------------------------------------------------------------ */
void arm_biquad_cascade_df1_init_f32(arm_biquad_casd_df1_inst_f32 *S, uint8_t numStages,
                                     const float32_t *pCoeffs, float32_t *pState) {
  S->numStages = numStages;
  S->pCoeffs = pCoeffs;
  S->pState = pState;
  memset(pState, 0, 4u * (size_t)numStages * sizeof(float32_t));
}

/*
 * T41/Process.cpp:705 (AM low-pass, 1 stage), T41/FFT.cpp:83-84 (zoom IIR, 4 stages).
 * Direct form I per stage, evaluated left to right:
 *   y = ((((b0*x) + (b1*x1)) + (b2*x2)) + (a1*y1)) + (a2*y2)
 * Stage s filters the whole block before stage s+1 starts (stage 0 reads pSrc,
 * later stages run in place on pDst).
 */
void arm_biquad_cascade_df1_f32(const arm_biquad_casd_df1_inst_f32 *S, const float32_t *pSrc,
                                float32_t *pDst, uint32_t blockSize) {
  const float32_t *in = pSrc;
  for (uint32_t s = 0; s < S->numStages; ++s) {
    const float32_t *c = S->pCoeffs + 5u * s;
    float32_t *st = S->pState + 4u * s;
    const float32_t b0 = c[0], b1 = c[1], b2 = c[2], a1 = c[3], a2 = c[4];
    float32_t x1 = st[0], x2 = st[1], y1 = st[2], y2 = st[3];
    for (uint32_t n = 0; n < blockSize; ++n) {
      float32_t x = in[n];
      float32_t acc = b0 * x;
      acc = acc + b1 * x1;
      acc = acc + b2 * x2;
      acc = acc + a1 * y1;
      acc = acc + a2 * y2;
      x2 = x1; x1 = x;
      y2 = y1; y1 = acc;
      pDst[n] = acc;
    }
    st[0] = x1; st[1] = x2; st[2] = y1; st[3] = y2;
    in = pDst;
  }
}

void arm_biquad_cascade_df2T_init_f32(arm_biquad_cascade_df2T_instance_f32 *S, uint8_t numStages,
                                      const float32_t *pCoeffs, float32_t *pState) {
  S->numStages = numStages;
  S->pCoeffs = pCoeffs;
  S->pState = pState;
  memset(pState, 0, 2u * (size_t)numStages * sizeof(float32_t));
}

/*
 * T41/Process.cpp:127-128 (DC-block high-pass, 1 stage; ONE instance/state used
 * for the I block and then the Q block).  Transposed direct form II:
 *   y  = (b0*x) + d1
 *   d1 = ((b1*x) + (a1*y)) + d2
 *   d2 = (b2*x) + (a2*y)
 */
void arm_biquad_cascade_df2T_f32(const arm_biquad_cascade_df2T_instance_f32 *S, const float32_t *pSrc,
                                 float32_t *pDst, uint32_t blockSize) {
  const float32_t *in = pSrc;
  for (uint32_t s = 0; s < S->numStages; ++s) {
    const float32_t *c = S->pCoeffs + 5u * s;
    float32_t *st = S->pState + 2u * s;
    const float32_t b0 = c[0], b1 = c[1], b2 = c[2], a1 = c[3], a2 = c[4];
    float32_t d1 = st[0], d2 = st[1];
    for (uint32_t n = 0; n < blockSize; ++n) {
      float32_t x = in[n];
      float32_t y = b0 * x + d1;
      float32_t t = b1 * x + a1 * y;
      d1 = t + d2;
      d2 = b2 * x + a2 * y;
      pDst[n] = y;
    }
    st[0] = d1; st[1] = d2;
    in = pDst;
  }
}

/* ------------------------------------------------------------------ */
/* complex FFT, N = 512 = 8^3: three radix-8 decimation-in-frequency   */
/* stages in place, then base-8 digit reversal                         */
/* (T41/Process.cpp:535,595,787,808; Filter.cpp:282; FFT.cpp:134,226)  */
/* ------------------------------------------------------------------ */
#define T41_FFT_N 512

static float32_t g_twiddle512[2 * T41_FFT_N]; /* cos(2*pi*i/512), sin(2*pi*i/512) */
static int g_twiddle_ready = 0;

static const float32_t *twiddle512(void) {
  if (!g_twiddle_ready) {
    for (int i = 0; i < T41_FFT_N; ++i) {
      double a = 2.0 * 3.14159265358979323846 * (double)i / (double)T41_FFT_N;
      g_twiddle512[2 * i] = (float32_t)cos(a);
      g_twiddle512[2 * i + 1] = (float32_t)sin(a);
    }
    /* exact quadrant values */
    g_twiddle512[2 * 0] = 1.0f;   g_twiddle512[2 * 0 + 1] = 0.0f;
    g_twiddle512[2 * 128] = 0.0f; g_twiddle512[2 * 128 + 1] = 1.0f;
    g_twiddle512[2 * 256] = -1.0f; g_twiddle512[2 * 256 + 1] = 0.0f;
    g_twiddle512[2 * 384] = 0.0f; g_twiddle512[2 * 384 + 1] = -1.0f;
    g_twiddle_ready = 1;
  }
  return g_twiddle512;
}

const arm_cfft_instance_f32 arm_cfft_sR_f32_len256 = {256, 0, 0, 0};
const arm_cfft_instance_f32 arm_cfft_sR_f32_len512 = {512, 0, 0, 0};
const arm_cfft_instance_f32 arm_cfft_sR_f32_len1024 = {1024, 0, 0, 0};
const arm_cfft_instance_f32 arm_cfft_sR_f32_len2048 = {2048, 0, 0, 0};

#define T41_C81 0.70710678118f

/*
 * 8-point DFT of (xr[m] + j xi[m]), m = 0..7, outputs in natural order.
 * Split as 4 radix-2 butterflies (distance 4), a 4-point DFT of the sums (even
 * bins) and a 4-point DFT of the rotated differences (odd bins); the +-45 degree
 * rotations of differences 1 and 3 are folded into four scaled sums.
 * The operation order below IS the definition the CUDA kernel replicates.
 */
static void dft8(const float32_t *xr, const float32_t *xi, float32_t *yr, float32_t *yi) {
  /* distance-4 butterflies */
  float32_t ar0 = xr[0] + xr[4], ai0 = xi[0] + xi[4];
  float32_t br0 = xr[0] - xr[4], bi0 = xi[0] - xi[4];
  float32_t ar1 = xr[1] + xr[5], ai1 = xi[1] + xi[5];
  float32_t br1 = xr[1] - xr[5], bi1 = xi[1] - xi[5];
  float32_t ar2 = xr[2] + xr[6], ai2 = xi[2] + xi[6];
  float32_t br2 = xr[2] - xr[6], bi2 = xi[2] - xi[6];
  float32_t ar3 = xr[3] + xr[7], ai3 = xi[3] + xi[7];
  float32_t br3 = xr[3] - xr[7], bi3 = xi[3] - xi[7];

  /* even bins: 4-point DFT of a */
  float32_t cr0 = ar0 + ar2, ci0 = ai0 + ai2;
  float32_t dr0 = ar0 - ar2, di0 = ai0 - ai2;
  float32_t cr1 = ar1 + ar3, ci1 = ai1 + ai3;
  float32_t dr1 = ar1 - ar3, di1 = ai1 - ai3;
  yr[0] = cr0 + cr1; yi[0] = ci0 + ci1;
  yr[4] = cr0 - cr1; yi[4] = ci0 - ci1;
  yr[2] = dr0 + di1; yi[2] = di0 - dr1;   /* d0 + (-j) d1 */
  yr[6] = dr0 - di1; yi[6] = di0 + dr1;

  /* odd bins */
  float32_t p = (br1 - br3) * T41_C81;
  float32_t q = (br1 + br3) * T41_C81;
  float32_t u = (bi1 - bi3) * T41_C81;
  float32_t v = (bi1 + bi3) * T41_C81;
  float32_t er0 = br0 + bi2, ei0 = bi0 - br2;   /* b0 + (-j) b2 */
  float32_t fr0 = br0 - bi2, fi0 = bi0 + br2;   /* b0 - (-j) b2 */
  float32_t er1 = p + v, ei1 = u - q;           /* W8 b1 + W8^3 b3 */
  float32_t fr1 = v - p, fi1 = q + u;           /* (-j)(W8 b1 - W8^3 b3) = (fr1, -fi1) */
  yr[1] = er0 + er1; yi[1] = ei0 + ei1;
  yr[5] = er0 - er1; yi[5] = ei0 - ei1;
  yr[3] = fr0 + fr1; yi[3] = fi0 - fi1;
  yr[7] = fr0 - fr1; yi[7] = fi0 + fi1;
}

static void radix8_dif_512(float32_t *buf) {
  const float32_t *tw = twiddle512();
  uint32_t n2 = T41_FFT_N;
  uint32_t stride = 1;            /* twiddle index step = N / n1 */
  while (n2 > 1u) {
    const uint32_t n1 = n2;
    n2 >>= 3;
    for (uint32_t j = 0; j < n2; ++j) {
      for (uint32_t i0 = j; i0 < T41_FFT_N; i0 += n1) {
        float32_t xr[8], xi[8], yr[8], yi[8];
        for (uint32_t m = 0; m < 8u; ++m) {
          xr[m] = buf[2u * (i0 + m * n2)];
          xi[m] = buf[2u * (i0 + m * n2) + 1u];
        }
        dft8(xr, xi, yr, yi);
        buf[2u * i0] = yr[0];
        buf[2u * i0 + 1u] = yi[0];
        for (uint32_t k = 1; k < 8u; ++k) {
          float32_t re = yr[k], im = yi[k];
          if (j != 0u) {
            /* multiply by exp(-j*2*pi*(j*k*stride)/N) */
            const uint32_t t = j * k * stride;
            const float32_t co = tw[2u * t], si = tw[2u * t + 1u];
            float32_t rc = yr[k] * co, is = yi[k] * si;
            float32_t ic = yi[k] * co, rs = yr[k] * si;
            re = rc + is;
            im = ic - rs;
          }
          buf[2u * (i0 + k * n2)] = re;
          buf[2u * (i0 + k * n2) + 1u] = im;
        }
      }
    }
    stride <<= 3;
  }
}

static uint32_t octal_reverse3(uint32_t p) {
  return ((p & 7u) << 6) | (p & 0x38u) | ((p >> 6) & 7u);
}

/* ------------------------------------------------------------------ */
/* complex FFT, N = 256 = 4 * 8 * 8 (the noise-reduction stages,        */
/* T41/Noise.cpp:206,279,454,636): one radix-4 decimation-in-frequency  */
/* stage over the four quarters, then two radix-8 stages inside each    */
/* 64-point quarter (CMSIS: arm_cfft_radix8by4_f32), then the mixed     */
/* digit reversal.  The operation order below IS the definition the     */
/* CUDA side replicates.                                                */
/* ------------------------------------------------------------------ */
#define T41_FFT_N256 256
static float32_t g_twiddle256[2 * T41_FFT_N256]; /* cos(2*pi*i/256), sin(2*pi*i/256) */
static int g_twiddle256_ready = 0;

static const float32_t *twiddle256(void) {
  if (!g_twiddle256_ready) {
    for (int i = 0; i < T41_FFT_N256; ++i) {
      double a = 2.0 * 3.14159265358979323846 * (double)i / (double)T41_FFT_N256;
      g_twiddle256[2 * i] = (float32_t)cos(a);
      g_twiddle256[2 * i + 1] = (float32_t)sin(a);
    }
    g_twiddle256[2 * 0] = 1.0f;    g_twiddle256[2 * 0 + 1] = 0.0f;
    g_twiddle256[2 * 64] = 0.0f;   g_twiddle256[2 * 64 + 1] = 1.0f;
    g_twiddle256[2 * 128] = -1.0f; g_twiddle256[2 * 128 + 1] = 0.0f;
    g_twiddle256[2 * 192] = 0.0f;  g_twiddle256[2 * 192 + 1] = -1.0f;
    g_twiddle256_ready = 1;
  }
  return g_twiddle256;
}

/* (re, im) * exp(-j*2*pi*t/256): the same four products and two sums as the 512-point stages */
static void twiddle_mul256(const float32_t *tw, uint32_t t, float32_t *re, float32_t *im) {
  const float32_t co = tw[2u * t], si = tw[2u * t + 1u];
  const float32_t rc = *re * co, is = *im * si;
  const float32_t ic = *im * co, rs = *re * si;
  *re = rc + is;
  *im = ic - rs;
}

static void dif_256(float32_t *buf) {
  const float32_t *tw = twiddle256();
  /* radix-4 stage: quarter q of the result holds the inputs of the 64-point transform of bins 4 k + q */
  for (uint32_t j = 0; j < 64u; ++j) {
    const float32_t ar = buf[2u * j], ai = buf[2u * j + 1u];
    const float32_t br = buf[2u * (j + 64u)], bi = buf[2u * (j + 64u) + 1u];
    const float32_t cr = buf[2u * (j + 128u)], ci = buf[2u * (j + 128u) + 1u];
    const float32_t dr = buf[2u * (j + 192u)], di = buf[2u * (j + 192u) + 1u];
    const float32_t s0r = ar + cr, s0i = ai + ci;      /* a + c */
    const float32_t d0r = ar - cr, d0i = ai - ci;      /* a - c */
    const float32_t s1r = br + dr, s1i = bi + di;      /* b + d */
    const float32_t d1r = br - dr, d1i = bi - di;      /* b - d */
    float32_t y0r = s0r + s1r, y0i = s0i + s1i;
    float32_t y2r = s0r - s1r, y2i = s0i - s1i;
    float32_t y1r = d0r + d1i, y1i = d0i - d1r;        /* (a - c) - j (b - d) */
    float32_t y3r = d0r - d1i, y3i = d0i + d1r;        /* (a - c) + j (b - d) */
    if (j != 0u) {
      twiddle_mul256(tw, j, &y1r, &y1i);
      twiddle_mul256(tw, 2u * j, &y2r, &y2i);
      twiddle_mul256(tw, 3u * j, &y3r, &y3i);
    }
    buf[2u * j] = y0r;            buf[2u * j + 1u] = y0i;
    buf[2u * (j + 64u)] = y1r;    buf[2u * (j + 64u) + 1u] = y1i;
    buf[2u * (j + 128u)] = y2r;   buf[2u * (j + 128u) + 1u] = y2i;
    buf[2u * (j + 192u)] = y3r;   buf[2u * (j + 192u) + 1u] = y3i;
  }
  /* two radix-8 stages inside each quarter: n1 = 64 (twiddle step 4 of the 256-table), then n1 = 8 (no twiddles) */
  for (uint32_t q = 0; q < 4u; ++q) {
    float32_t *b = buf + 2u * 64u * q;
    uint32_t n2 = 64u, stride = 4u;
    while (n2 > 1u) {
      const uint32_t n1 = n2;
      n2 >>= 3;
      for (uint32_t j = 0; j < n2; ++j) {
        for (uint32_t i0 = j; i0 < 64u; i0 += n1) {
          float32_t xr[8], xi[8], yr[8], yi[8];
          for (uint32_t m = 0; m < 8u; ++m) {
            xr[m] = b[2u * (i0 + m * n2)];
            xi[m] = b[2u * (i0 + m * n2) + 1u];
          }
          dft8(xr, xi, yr, yi);
          b[2u * i0] = yr[0];
          b[2u * i0 + 1u] = yi[0];
          for (uint32_t k = 1; k < 8u; ++k) {
            float32_t re = yr[k], im = yi[k];
            if (j != 0u) twiddle_mul256(tw, j * k * stride, &re, &im);
            b[2u * (i0 + k * n2)] = re;
            b[2u * (i0 + k * n2) + 1u] = im;
          }
        }
      }
      stride <<= 3;
    }
  }
}

/* position p = 64 q + 8 d1 + d0 holds bin 4 (8 d0 + d1) + q */
static uint32_t digit_reverse_256(uint32_t p) {
  return 4u * (8u * (p & 7u) + ((p >> 3) & 7u)) + (p >> 6);
}

static void cfft256(float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag) {
  if (ifftFlag) {
    for (uint32_t i = 0; i < T41_FFT_N256; ++i) p1[2u * i + 1u] = -p1[2u * i + 1u];
  }
  dif_256(p1);
  if (bitReverseFlag) {
    float32_t tmp[2 * T41_FFT_N256];
    memcpy(tmp, p1, sizeof(tmp));
    for (uint32_t p = 0; p < T41_FFT_N256; ++p) {
      const uint32_t k = digit_reverse_256(p);
      p1[2u * k] = tmp[2u * p];
      p1[2u * k + 1u] = tmp[2u * p + 1u];
    }
  }
  if (ifftFlag) {
    const float32_t invL = 1.0f / (float32_t)T41_FFT_N256;
    for (uint32_t i = 0; i < T41_FFT_N256; ++i) {
      p1[2u * i] = p1[2u * i] * invL;
      p1[2u * i + 1u] = -p1[2u * i + 1u] * invL;
    }
  }
}

void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag) {
  if (S->fftLen == T41_FFT_N256) {
    cfft256(p1, ifftFlag, bitReverseFlag);
    return;
  }
  if (S->fftLen != T41_FFT_N) abort(); /* only the 512- and 256-point transforms are on the RX path */
  if (ifftFlag) {
    for (uint32_t i = 0; i < T41_FFT_N; ++i) p1[2u * i + 1u] = -p1[2u * i + 1u];
  }
  radix8_dif_512(p1);
  if (bitReverseFlag) {
    for (uint32_t p = 0; p < T41_FFT_N; ++p) {
      uint32_t r = octal_reverse3(p);
      if (r > p) {
        float32_t tr = p1[2u * p], ti = p1[2u * p + 1u];
        p1[2u * p] = p1[2u * r]; p1[2u * p + 1u] = p1[2u * r + 1u];
        p1[2u * r] = tr; p1[2u * r + 1u] = ti;
      }
    }
  }
  if (ifftFlag) {
    const float32_t invL = 1.0f / (float32_t)T41_FFT_N;
    for (uint32_t i = 0; i < T41_FFT_N; ++i) {
      p1[2u * i] = p1[2u * i] * invL;
      p1[2u * i + 1u] = -p1[2u * i + 1u] * invL;
    }
  }
}
