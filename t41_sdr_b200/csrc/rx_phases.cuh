/*
 * rx_phases.cuh — the fused T41 receive chain as a sequence of barrier-separated phases.
 *
 * One CTA owns G consecutive receivers ("slots") for all n_blocks blocks of a launch and
 * keeps every intermediate of the chain in shared memory: per block a receiver reads
 * 16 KiB of I/Q from HBM and writes 8 KiB of audio (plus 2 KiB of spectrum / waterfall row
 * on row-producing blocks) and nothing else.  Per-receiver state (StreamState) is loaded
 * from HBM once per launch and written back once.
 *
 * The chain is ProcessIQData() (reference Process.cpp:70-944).  Stage arithmetic follows
 * the CPU oracle operation by operation (separately rounded float ops, fmaf() only in FIR
 * tap accumulation), so everything except the default closed-form NCO is bit-identical to
 * the oracle; see DESIGN.md "Numerics".  Compile with --fmad=false.
 *
 * Each phase is a function of (cta, tid) that needs no intra-phase synchronisation, which
 * is what lets tests/devtools compile this very file for the host and step the phases with
 * a plain loop over tid (a development aid, never part of libt41rx.so).
 */
#ifndef T41RX_PHASES_CUH
#define T41RX_PHASES_CUH

#include <math.h>
#include <stdint.h>

#include "rx_fft.h"
#include "rx_types.h"

#ifdef T41RX_HOST_EMUL
#define T41RX_DEV inline
template <class T> inline T LdgRO(const T *p) { return *p; }
#else
#define T41RX_DEV __device__ __forceinline__
template <class T> __device__ __forceinline__ T LdgRO(const T *p) { return __ldg(p); }
#endif

namespace t41rx {

constexpr int kG = 4;                 /* receivers per CTA */
constexpr int kNT = 64 * kG;          /* threads per CTA: 64 per receiver in the FFT phases */
constexpr int kDcChunks = 8;          /* time-parallel chunks of the 4096-step DC-block chain */
constexpr int kDcChunkLen = 513;      /* chunk c starts at c*513 (odd: bank-conflict-free lanes) */
constexpr int kDcWarm = 256;          /* speculative warm-up length (a1 = 0.854: 0.854^256 ~ 3e-18) */

/* ---- shared-memory slot layout, in floats ---- */
constexpr int kRawLen = 2076;                     /* 27 history + 2048 new (+1 pad) */
constexpr int oRawI = 0;
constexpr int oRawQ = oRawI + kRawLen;            /* 2076 */
constexpr int oD1I = oRawQ + kRawLen;             /* 4152 : 45 history + 512 new */
constexpr int kD1Len = 560;
constexpr int oD1Q = oD1I + kD1Len;
constexpr int oOla = oD1Q + kD1Len;               /* 5272 : prev I[256], prev Q[256] */
constexpr int oTaps = oOla + 512;                 /* 5784 : dec1 28 | dec2 46 | int1 48 | int2 32 */
constexpr int kTapDec1 = 0, kTapDec2 = 28, kTapInt1 = 74, kTapInt2 = 122;
constexpr int oAgc = oTaps + 156;                 /* 5940 : re[128] im[128] abs[128] */
constexpr int oNco = oAgc + 384;                  /* 6324 : W[64] C[32] CB[32] as double2 */
constexpr int oD1H = oNco + 512;                  /* 6836 : dec1 history 2 x 27 (-> 56) */
constexpr int oD2H = oD1H + 56;                   /* 6892 : dec2 history 2 x 45 (-> 92) */
constexpr int oIntH = oD2H + 92;                  /* 6984 : int1 23 (-> 24) | int2 7 (-> 8) */
constexpr int oMisc = oIntH + 32;                 /* 7016 : scalars */
constexpr int kSlotRaw = oMisc + 64;              /* 7080 */
constexpr int kSlot = ((kSlotRaw - 8 + 31) / 32) * 32 + 8;   /* == 8 (mod 32): distinct banks per slot */
static_assert(kSlot % 32 == 8 && kSlot >= kSlotRaw, "slot stride");
static_assert((oNco % 4) == 0, "double2 alignment");
constexpr int kSmemFloats = kG * kSlot;

/* overlays on the raw region once dec1 has consumed it */
constexpr int vFftA = 0;                          /* 512 complex */
constexpr int vFftB = 1024;                       /* 512 complex */
constexpr int vZext = 2048;                       /* 353 complex: 97 delayed + 256 new AGC samples */
constexpr int vAbs = vZext + 708;                 /* 2756 : 353 */
constexpr int vLvlA = vAbs + 356;                 /* 3112 : 353 */
constexpr int vLvlB = vLvlA + 356;                /* 3468 : 353 */
constexpr int vRm = vLvlB + 356;                  /* 3824 : 256 -> 4080 */
static_assert(vRm + 256 <= 2 * kRawLen, "overlay fits");
/* after the AGC the FFT buffers are dead */
constexpr int vVolt = vFftA;                      /* 256 volts samples (written by the serial AGC pass) */
constexpr int vDem = 256;                         /* 256 complex AGC output -> 768 */
constexpr int vAud = 768;                         /* int1 state: 23 history + 256 -> 1047 */
constexpr int vAmTmp = 1056;                      /* 256 */
constexpr int vInt2 = 1312;                       /* int2 state: 7 history + 512 -> 1831 */
/* spectrum scratch (row blocks, before dec1): the D1 region */
constexpr int vSpecFft = oD1I;                    /* 512 complex = 1024 floats <= 1120 */

/* misc scalar indices */
enum { mDcD1 = 0, mDcD2 = 1, mDcSpec = 2 /* 8 x 2 */, mDcEnd = 18 /* 8 x 2 */, mNcoMode = 34, mRowFlag = 35 };

struct LaunchArgs {
  const float *iq;
  float *audio;
  int16_t *spec_rows;
  uint16_t *wf_rows;
  int8_t *psk_bits;
  uint8_t *psk_chars;
  const StreamCfg *cfg;
  StreamState *st;
  const FilterSet *fsets;
  const double *nco_tab;      /* per stream: W[64] then C[32], (cos, sin) pairs */
  const float2 *twiddle;      /* 512 */
  const double *hann;         /* 512: 0.5 - 0.5*cos(6.28*i/512) */
  const float *sin_table;     /* 513: sin(2*pi*k/512), arm_sin_f32's table */
  const float *zoom_iir;      /* 4 x 20: zoom x2..x16 biquad coefficients (FIR.cpp:582-885) */
  const float *sam_consts;    /* omega_min, omega_max, g1, g2 (Demod.cpp:13-18) */
  const uint16_t *gradient;   /* 117 */
  const uint32_t *varicode;   /* 128: code | bits << 16 | ascii << 24 */
  int n_streams, n_blocks, row_every, n_rows;
  uint32_t flags;
};

struct Cta {
  LaunchArgs a;
  float *smem;
  int s0;      /* first receiver of this CTA */
  int ng;      /* receivers handled (<= kG) */
  int t;       /* block index within the launch */
  int row;     /* this block produces a spectrum row */
  int row_idx;
};

T41RX_DEV float *Slot(const Cta &c, int g) { return c.smem + g * kSlot; }

/* ------------------------------------------------------------------ */
/* scalar helpers (Utility.cpp / Demod.cpp restated for the device)    */
/* ------------------------------------------------------------------ */
T41RX_DEV float Log10Fast(float x) {              /* Utility.cpp:245-258 */
  int e;
  const float f = frexpf(fabsf(x), &e);
  float y = 1.23149591368684f;
  y *= f;
  y += -4.11852516267426f;
  y *= f;
  y += 6.02197014179219f;
  y *= f;
  y += -3.13396450166353f;
  y += (float)e;
  return y * 0.3010299956639812f;
}

T41RX_DEV float AlphaBetaMag(float i, float q) {  /* Utility.cpp:269-285 */
  const float alpha = 0.960433870103f;
  const float beta = 0.397824734759f;
  const float ai = fabsf(i), aq = fabsf(q);
  if (ai > aq) return alpha * ai + beta * aq;
  return alpha * aq + beta * ai;
}

T41RX_DEV float AtanPoly(float z) {               /* Utility.cpp:298-302 */
  const float n1 = 0.97239411f;
  const float n2 = -0.19194795f;
  return (n1 + n2 * z * z) * z;
}

T41RX_DEV float Atan2Approx(float y, float x) {   /* Demod.cpp:148-197 (TPI quirk kept) */
  const float pi = 3.1415926535897932384626433832795f;
  const float tpi = 6.283185307179586476925286766559f;
  if (x != 0.0f) {
    if (fabsf(x) > fabsf(y)) {
      const float z = y / x;
      if (x > 0.0f) return AtanPoly(z);
      if (y >= 0.0f) return AtanPoly(z) + pi;
      return AtanPoly(z) - pi;
    }
    const float z = x / y;
    if (y > 0.0f) return -AtanPoly(z) + tpi;
    return -AtanPoly(z) - tpi;
  }
  if (y > 0.0f) return tpi;
  if (y < 0.0f) return -tpi;
  return 0.0f;
}

/* arm_sin_f32 / arm_cos_f32 tail: linear interpolation in the 513-entry table */
T41RX_DEV float TableTurns(const float *tab, float in) {
  int32_t n = (int32_t)in;
  if (in < 0.0f) n--;
  in = in - (float)n;
  float findex = 512.0f * in;
  uint32_t index = (uint32_t)findex & 0xFFFFu;
  if (index >= 512u) {
    index = 0;
    findex -= 512.0f;
  }
  const float fract = findex - (float)index;
  const float a = LdgRO(tab + index);
  const float b = LdgRO(tab + index + 1);
  const float wa = (1.0f - fract) * a;
  const float wb = fract * b;
  return wa + wb;
}

/* Process.cpp:165-174 + Utility.cpp:178-187 on one sample of the conditioned buffers */
T41RX_DEV void IqCorr(const StreamCfg &cf, float &i, float &q) {
  if (cf.mirrored) {
    if (cf.iq_phase < 0.0f) q = q + i * cf.iq_phase;
    else i = i + q * cf.iq_phase;
  }
}

/* FreqShift1 (Freq_Shift.cpp:42-65): multiply sample n by exp(+j*pi*n/2) */
T41RX_DEV void QuarterShift(int n, float &i, float &q) {
  const float a = i, b = q;
  switch (n & 3) {
    case 1: i = -b; q = a; break;
    case 2: i = -a; q = -b; break;
    case 3: i = b; q = -a; break;
    default: break;
  }
}

/* one step of the DC-block biquad (arm_biquad_cascade_df2T_f32, 1 stage; FIR.cpp:87-89) */
struct DcCoef { float b0, b1, b2, a1, a2; };
T41RX_DEV float DcStep(const DcCoef &k, float x, float &d1, float &d2) {
  const float y = k.b0 * x + d1;
  const float t = k.b1 * x + k.a1 * y;
  d1 = t + d2;
  d2 = k.b2 * x + k.a2 * y;
  return y;
}
T41RX_DEV DcCoef DcCoefs() {
  return DcCoef{(float)0.927176191943378969, (float)-0.927176191943378969, (float)0.0,
                (float)0.854352383886757938, (float)0.0};
}

/* raw sequence index (I block then Q block, B6) -> float offset inside the slot */
T41RX_DEV int SeqOff(int i) { return i < kBlock ? (oRawI + 27 + i) : (oRawQ + 27 + (i - kBlock)); }

/* ------------------------------------------------------------------ */
/* launch prologue / epilogue: state <-> shared memory                 */
/* ------------------------------------------------------------------ */
T41RX_DEV void PhStateIn(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamState &st = c.a.st[c.s0 + g];
    const StreamCfg &cf = c.a.cfg[c.s0 + g];
    const FilterSet &fs = c.a.fsets[cf.filter_id];
    for (int i = tid; i < 512; i += kNT) s[oOla + i] = st.ola_prev[i >> 8][i & 255];
    for (int i = tid; i < 154; i += kNT) {
      float v;
      if (i < kTapDec2) v = fs.dec1[i];
      else if (i < kTapInt1) v = fs.dec2[i - kTapDec2];
      else if (i < kTapInt2) v = fs.int1[i - kTapInt1];
      else v = fs.int2[i - kTapInt2];
      s[oTaps + i] = v;
    }
    for (int i = tid; i < 128; i += kNT) {
      s[oAgc + i] = st.agc_re[i];
      s[oAgc + 128 + i] = st.agc_im[i];
      s[oAgc + 256 + i] = st.agc_abs[i];
    }
    double *nco = reinterpret_cast<double *>(s + oNco);
    const double *tab = c.a.nco_tab + (size_t)(c.s0 + g) * 192;
    for (int i = tid; i < 192; i += kNT) nco[i] = tab[i];
    for (int i = tid; i < 54; i += kNT) s[oD1H + i] = st.dec1_hist[i / 27][i % 27];
    for (int i = tid; i < 90; i += kNT) s[oD2H + i] = st.dec2_hist[i / 45][i % 45];
    for (int i = tid; i < 23; i += kNT) s[oIntH + i] = st.int1_hist[i];
    for (int i = tid; i < 7; i += kNT) s[oIntH + 24 + i] = st.int2_hist[i];
    if (tid == 0) {
      s[oMisc + mDcD1] = st.dc_d1;
      s[oMisc + mDcD2] = st.dc_d2;
    }
  }
}

T41RX_DEV void PhStateOut(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    StreamState &st = c.a.st[c.s0 + g];
    for (int i = tid; i < 512; i += kNT) st.ola_prev[i >> 8][i & 255] = s[oOla + i];
    for (int i = tid; i < 128; i += kNT) {
      st.agc_re[i] = s[oAgc + i];
      st.agc_im[i] = s[oAgc + 128 + i];
      st.agc_abs[i] = s[oAgc + 256 + i];
    }
    for (int i = tid; i < 54; i += kNT) st.dec1_hist[i / 27][i % 27] = s[oD1H + i];
    for (int i = tid; i < 90; i += kNT) st.dec2_hist[i / 45][i % 45] = s[oD2H + i];
    for (int i = tid; i < 23; i += kNT) st.int1_hist[i] = s[oIntH + i];
    for (int i = tid; i < 7; i += kNT) st.int2_hist[i] = s[oIntH + 24 + i];
    if (tid == 0) {
      st.dc_d1 = s[oMisc + mDcD1];
      st.dc_d2 = s[oMisc + mDcD2];
    }
  }
}

/* ------------------------------------------------------------------ */
/* P0: HBM -> shared, de-interleave; restore dec1 history              */
/* ------------------------------------------------------------------ */
T41RX_DEV void PhLoad(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const float4 *src = reinterpret_cast<const float4 *>(
        c.a.iq + ((size_t)(c.s0 + g) * c.a.n_blocks + c.t) * (2 * kBlock));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = tid + kNT * k;            /* float4 index: samples 2j, 2j+1 */
      const float4 v = LdgRO(src + j);
      s[oRawI + 27 + 2 * j] = v.x;
      s[oRawQ + 27 + 2 * j] = v.y;
      s[oRawI + 27 + 2 * j + 1] = v.z;
      s[oRawQ + 27 + 2 * j + 1] = v.w;
    }
    if (tid < 54) {
      const int ch = tid / 27, i = tid % 27;
      s[(ch ? oRawQ : oRawI) + i] = s[oD1H + tid];
    }
  }
}

/* ------------------------------------------------------------------ */
/* P1: input conditioning (Process.cpp:117-134,165-166)                */
/*   x * rfGainValue -> DC-block biquad over I[0..2047] then Q[0..2047] with ONE state (B6)
 *   -> * RFgain -> (I only, mirrored modes) * -IQAmp.
 * The 4096-step recurrence is split into kDcChunks chunks that run on separate lanes.
 * A chunk other than the first starts from a state obtained by filtering the kDcWarm
 * samples before it from zero: the filter's pole (0.854) makes that state converge to the
 * true one to the last bit.  P1c checks every chunk's start state against the previous
 * chunk's end state bit for bit and recomputes serially from the first mismatch, so the
 * result is always exactly the serial recurrence's.                                       */
/* ------------------------------------------------------------------ */
T41RX_DEV void PhDcWarm(Cta &c, int tid) {
  if (tid >= c.ng * kDcChunks) return;
  const int g = tid / kDcChunks, ch = tid % kDcChunks;
  if (ch == 0) return;
  float *s = Slot(c, g);
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  const DcCoef k = DcCoefs();
  const float rfg = cf.rf_gain_value;
  float d1 = 0.0f, d2 = 0.0f;
  const int start = ch * kDcChunkLen - kDcWarm;
  for (int i = start; i < start + kDcWarm; ++i) {
    const float x = s[SeqOff(i)] * rfg;
    (void)DcStep(k, x, d1, d2);
  }
  s[oMisc + mDcSpec + 2 * ch] = d1;
  s[oMisc + mDcSpec + 2 * ch + 1] = d2;
}

T41RX_DEV void DcRunChunk(float *s, const StreamCfg &cf, float rfgain_f, int begin, int end, float &d1, float &d2) {
  const DcCoef k = DcCoefs();
  const float rfg = cf.rf_gain_value;
  for (int i = begin; i < end; ++i) {
    const int off = SeqOff(i);
    const float x = s[off] * rfg;
    float y = DcStep(k, x, d1, d2) * rfgain_f;
    if (cf.mirrored && i < kBlock) y = y * cf.neg_iq_amp;
    s[off] = y;
  }
}

T41RX_DEV void PhDcMain(Cta &c, int tid) {
  if (tid >= c.ng * kDcChunks) return;
  const int g = tid / kDcChunks, ch = tid % kDcChunks;
  float *s = Slot(c, g);
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  const float rfgain_f = (float)c.a.st[c.s0 + g].rf_gain;
  float d1, d2;
  if (ch == 0) {
    d1 = s[oMisc + mDcD1];
    d2 = s[oMisc + mDcD2];
  } else {
    d1 = s[oMisc + mDcSpec + 2 * ch];
    d2 = s[oMisc + mDcSpec + 2 * ch + 1];
  }
  const int begin = ch * kDcChunkLen;
  const int end = (ch == kDcChunks - 1) ? 2 * kBlock : begin + kDcChunkLen;
  DcRunChunk(s, cf, rfgain_f, begin, end, d1, d2);
  s[oMisc + mDcEnd + 2 * ch] = d1;
  s[oMisc + mDcEnd + 2 * ch + 1] = d2;
}

T41RX_DEV bool SameBits(float a, float b) {
  union { float f; uint32_t u; } x, y;
  x.f = a;
  y.f = b;
  return x.u == y.u;
}

T41RX_DEV void PhDcVerify(Cta &c, int tid) {
  if (tid >= c.ng) return;
  const int g = tid;
  float *s = Slot(c, g);
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  const float rfgain_f = (float)c.a.st[c.s0 + g].rf_gain;
  int bad = -1;
  for (int ch = 1; ch < kDcChunks; ++ch) {
    if (!SameBits(s[oMisc + mDcSpec + 2 * ch], s[oMisc + mDcEnd + 2 * (ch - 1)]) ||
        !SameBits(s[oMisc + mDcSpec + 2 * ch + 1], s[oMisc + mDcEnd + 2 * (ch - 1) + 1])) {
      bad = ch;
      break;
    }
  }
  if (bad >= 0) {
    /* speculation missed (vanishingly rare): redo serially from the first bad chunk,
       re-reading the raw samples from HBM/L2 because the chunk was filtered in place */
    const float *src = c.a.iq + ((size_t)(c.s0 + g) * c.a.n_blocks + c.t) * (2 * kBlock);
    float d1 = s[oMisc + mDcEnd + 2 * (bad - 1)];
    float d2 = s[oMisc + mDcEnd + 2 * (bad - 1) + 1];
    for (int i = bad * kDcChunkLen; i < 2 * kBlock; ++i)
      s[SeqOff(i)] = LdgRO(src + (i < kBlock ? 2 * i : 2 * (i - kBlock) + 1));
    DcRunChunk(s, cf, rfgain_f, bad * kDcChunkLen, 2 * kBlock, d1, d2);
    s[oMisc + mDcEnd + 2 * (kDcChunks - 1)] = d1;
    s[oMisc + mDcEnd + 2 * (kDcChunks - 1) + 1] = d2;
  }
  s[oMisc + mDcD1] = s[oMisc + mDcEnd + 2 * (kDcChunks - 1)];
  s[oMisc + mDcD2] = s[oMisc + mDcEnd + 2 * (kDcChunks - 1) + 1];
}

/* ------------------------------------------------------------------ */
/* P2: display spectrum on row-producing blocks (FFT.cpp:67-251)        */
/* ------------------------------------------------------------------ */
/* zoom index >= 1: 4-stage elliptic DF1 biquad cascade on the Fs/4-shifted I and Q, 4-tap
 * FIR decimation by 2^zoom, first zoom_samples outputs into the 512-deep ring.  One lane
 * per (receiver, channel); the cascade is evaluated sample by sample through all four
 * stages, which yields the same values as the reference's stage-by-stage order.           */
T41RX_DEV void PhZoomIir(Cta &c, int tid) {
  if (!c.row || tid >= c.ng * 2) return;
  const int g = tid >> 1, chn = tid & 1;
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  if (cf.zoom == 0) return;
  StreamState &st = c.a.st[c.s0 + g];
  const float *s = Slot(c, g);
  float k[20];
  for (int i = 0; i < 20; ++i) k[i] = LdgRO(c.a.zoom_iir + (cf.zoom - 1) * 20 + i);
  float zs[16];
  for (int i = 0; i < 16; ++i) zs[i] = st.zoom_iir[chn][i];
  float h0 = st.zoom_fir_hist[chn][0], h1 = st.zoom_fir_hist[chn][1], h2 = st.zoom_fir_hist[chn][2];
  const int M = 1 << cf.zoom;
  int ptr = st.zoom_ptr;
  int produced = 0;
  for (int n = 0; n < kBlock; ++n) {
    float vi = s[oRawI + 27 + n], vq = s[oRawQ + 27 + n];
    IqCorr(cf, vi, vq);
    QuarterShift(n, vi, vq);
    float x = chn ? vq : vi;
#pragma unroll
    for (int sg = 0; sg < 4; ++sg) {
      float *z = zs + 4 * sg;
      const float *kk = k + 5 * sg;
      float acc = kk[0] * x;
      acc = acc + kk[1] * z[0];
      acc = acc + kk[2] * z[1];
      acc = acc + kk[3] * z[2];
      acc = acc + kk[4] * z[3];
      z[1] = z[0]; z[0] = x;
      z[3] = z[2]; z[2] = acc;
      x = acc;
    }
    if ((n & (M - 1)) == 0) {
      /* arm_fir_decimate_f32, 4 taps: output n/M correlates the 3 older IIR outputs and this one */
      float acc = 0.0f;
      acc = fmaf(h0, cf.zoom_fir[0], acc);
      acc = fmaf(h1, cf.zoom_fir[1], acc);
      acc = fmaf(h2, cf.zoom_fir[2], acc);
      acc = fmaf(x, cf.zoom_fir[3], acc);
      if (produced < cf.zoom_samples) {
        st.zoom_ring[chn][ptr] = acc;
        if (++ptr >= kSpecRes) ptr = 0;
      }
      ++produced;
    }
    h0 = h1; h1 = h2; h2 = x;
  }
  for (int i = 0; i < 16; ++i) st.zoom_iir[chn][i] = zs[i];
  st.zoom_fir_hist[chn][0] = h0;
  st.zoom_fir_hist[chn][1] = h1;
  st.zoom_fir_hist[chn][2] = h2;
  if (chn == 1) st.zoom_ptr = ptr;   /* both channels advance identically; written after the ring writes */
}

/* window the 512 samples into the spectrum FFT buffer */
T41RX_DEV void PhSpecWindow(Cta &c, int tid) {
  if (!c.row) return;
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  float *s = Slot(c, g);
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  const StreamState &st = c.a.st[c.s0 + g];
  float2 *buf = reinterpret_cast<float2 *>(s + vSpecFft);
  for (int j = 0; j < 8; ++j) {
    const int i = u + 64 * j;
    const double w = LdgRO(c.a.hann + i);
    float re, im;
    if (cf.zoom == 0) {          /* CalcZoom1Magn, FFT.cpp:220-223: raw (pre-shift) samples */
      float vi = s[oRawI + 27 + i], vq = s[oRawQ + 27 + i];
      IqCorr(cf, vi, vq);
      re = (float)((double)vi * w);
      im = (float)((double)vq * w);
    } else {                     /* ZoomFFTExe, FFT.cpp:109-116: ring, oldest first */
      const int p = (st.zoom_ptr + i) & (kSpecRes - 1);
      const float a = cf.zoom_mult * st.zoom_ring[0][p];
      const float b = cf.zoom_mult * st.zoom_ring[1][p];
      re = (float)((double)a * w);
      im = (float)((double)b * w);
    }
    buf[i] = float2{re, im};
  }
}

T41RX_DEV void PhSpecFftPass(Cta &c, int tid, int pass) {
  if (!c.row) return;
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  Radix8Butterfly(reinterpret_cast<float2 *>(Slot(c, g) + vSpecFft), c.a.twiddle, pass, u);
}

/* |X|^2 with half swap -> smoothing -> log -> pixel -> spectrum and waterfall rows */
T41RX_DEV void PhSpecRow(Cta &c, int tid) {
  if (!c.row) return;
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  float *s = Slot(c, g);
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  StreamState &st = c.a.st[c.s0 + g];
  const float2 *buf = reinterpret_cast<const float2 *>(s + vSpecFft);
  const size_t row_base = ((size_t)(c.s0 + g) * c.a.n_rows + c.row_idx) * kSpecRes;
  const float lpf = 0.7f;
  for (int j = 0; j < 8; ++j) {
    const int x = u + 64 * j;
    const int bin = (x + 256) & 511;
    const float2 v = buf[OctRev3((unsigned)bin)];
    const float pw = v.x * v.x + v.y * v.y;
    const float old = st.spec_old[x];
    float shown;
    if (cf.zoom == 0) {
      /* spec_help = LPFcoeff * new + (1.0 - LPFcoeff) * old: the second product and the sum
         are double (FFT.cpp:241); the pixel uses the UNSMOOTHED value (B8) */
      const float a = lpf * pw;
      const float smooth = (float)((double)a + (1.0 - (double)lpf) * (double)old);
      st.spec_old[x] = smooth;
      shown = pw;
    } else {
      const float onem = (float)(1.0 - (double)lpf);
      const float smooth = lpf * pw + onem * old;
      st.spec_old[x] = smooth;
      shown = smooth;
    }
    const int16_t dbpix = (int16_t)(cf.db_scale * Log10Fast(shown));   /* truncation toward zero (B10) */
    const int16_t pix = (int16_t)(cf.pixel_add + (int)dbpix);
    if (c.a.spec_rows) c.a.spec_rows[row_base + x] = pix;
    if (c.a.wf_rows) {
      uint16_t colour = 0;
      if (x < kSpecRes - 1) {    /* Display.cpp:259: x1 = 0..510 (B18) */
        int y = cf.wf_base - (int)pix;
        if (y > 249) y = 249;
        if (y < 100) y = 100;
        int idx = 230 - y;
        if (idx < 0) idx = 0;
        if (idx > 116) idx = 116;
        colour = LdgRO(c.a.gradient + idx);
      }
      c.a.wf_rows[row_base + x] = colour;
    }
  }
}

/* ------------------------------------------------------------------ */
/* P3: FreqShift1 + FreqShift2 (Freq_Shift.cpp:42-141)                  */
/* ------------------------------------------------------------------ */
struct D2 { double x, y; };

/* per block: decide exact vs closed form, and build CB[m] = amp * exp(j(phi + delta + 64 m delta)) */
T41RX_DEV void PhNcoPrep(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  float *s = Slot(c, g);
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  const StreamState &st = c.a.st[c.s0 + g];
  const bool exact = (c.a.flags & 1u) || !st.nco_closed || (st.nco_epoch_seen != cf.nco_epoch);
  if (u == 0) s[oMisc + mNcoMode] = exact ? 1.0f : 0.0f;
  if (exact || u >= 32) return;
  D2 *tab = reinterpret_cast<D2 *>(s + oNco);
  double sn, cs;
  sincos(st.nco_phase + cf.nco_delta, &sn, &cs);
  const double bx = cf.nco_amp * cs, by = cf.nco_amp * sn;
  const D2 cm = tab[64 + u];
  tab[96 + u] = D2{bx * cm.x - by * cm.y, bx * cm.y + by * cm.x};
}

T41RX_DEV void MixStore(float *s, int n, float vi, float vq, double oq, double oi) {
  const float f = 1.1f;                       /* freqAdjFactor */
  const float a = vi * f, b = vq * f;
  s[oRawI + 27 + n] = (float)(((double)a * oq) + ((double)b * oi));
  s[oRawQ + 27 + n] = (float)(((double)b * oq) - ((double)a * oi));
}

T41RX_DEV void PhMix(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamCfg &cf = c.a.cfg[c.s0 + g];
    if (s[oMisc + mNcoMode] != 0.0f) {
      /* exact path: the FP64 oscillator recurrence, one lane per receiver */
      if (tid != g) continue;
      StreamState &st = c.a.st[c.s0 + g];
      double vq, vi;
      if (st.nco_closed) {
        /* leaving closed form (retune or forced): rebuild the vector at the settled radius */
        double sn, cs;
        sincos(st.nco_phase, &sn, &cs);
        const double r = sqrt(st.osc_q * st.osc_q + st.osc_i * st.osc_i);
        vq = r * cs;
        vi = r * sn;
      } else {
        vq = st.osc_q;
        vi = st.osc_i;
      }
      const double oc = cf.osc_cos, os = cf.osc_sin;
      for (int n = 0; n < kBlock; ++n) {
        const double oq = (vq * oc) - (vi * os);
        const double oi = (vi * oc) + (vq * os);
        const double gain = 1.95 - ((vq * vq) + (vi * vi));
        vq = gain * oq;
        vi = gain * oi;
        float xi = s[oRawI + 27 + n], xq = s[oRawQ + 27 + n];
        IqCorr(cf, xi, xq);
        QuarterShift(n, xi, xq);
        MixStore(s, n, xi, xq, oq, oi);
      }
      st.osc_q = vq;
      st.osc_i = vi;
      st.nco_epoch_seen = cf.nco_epoch;
      const double r2 = vq * vq + vi * vi;
      const bool settled = fabs(r2 - cf.nco_r2_fix) < 4.0e-15;
      if (settled && !(c.a.flags & 1u)) {
        st.nco_closed = 1;
        double ph = atan2(vi, vq);
        if (ph < 0) ph += 6.283185307179586476925286766559;
        st.nco_phase = ph;
      } else {
        st.nco_closed = 0;
      }
    } else {
      const D2 *tab = reinterpret_cast<const D2 *>(s + oNco);
      const D2 w = tab[tid & 63];
#pragma unroll
      for (int k = 0; k < kBlock / kNT; ++k) {
        const int n = tid + kNT * k;
        const D2 cb = tab[96 + (n >> 6)];
        const double oq = cb.x * w.x - cb.y * w.y;
        const double oi = cb.x * w.y + cb.y * w.x;
        float xi = s[oRawI + 27 + n], xq = s[oRawQ + 27 + n];
        IqCorr(cf, xi, xq);
        QuarterShift(n, xi, xq);
        MixStore(s, n, xi, xq, oq, oi);
      }
    }
  }
}

/* closed form: advance the phase by one block */
T41RX_DEV void PhNcoAdvance(Cta &c, int tid) {
  if (tid >= c.ng) return;
  float *s = Slot(c, tid);
  if (s[oMisc + mNcoMode] != 0.0f) return;
  const StreamCfg &cf = c.a.cfg[c.s0 + tid];
  StreamState &st = c.a.st[c.s0 + tid];
  double ph = st.nco_phase + cf.nco_block_delta;
  const double two_pi = 6.283185307179586476925286766559;
  if (ph >= two_pi) ph -= two_pi;
  st.nco_phase = ph;
}

/* ------------------------------------------------------------------ */
/* P4/P5: arm_fir_decimate_f32 x4 (28 taps) then x2 (46 taps)           */
/* (Process.cpp:262-267,378-386,474-479)                                */
/* ------------------------------------------------------------------ */
T41RX_DEV void PhDec1(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    float taps[kDec1Taps];
#pragma unroll
    for (int i = 0; i < kDec1Taps; ++i) taps[i] = s[oTaps + kTapDec1 + i];
    if (tid < 90) {               /* dec2 history back in front of the dec1 output */
      const int ch = tid / 45, i = tid % 45;
      s[(ch ? oD1Q : oD1I) + i] = s[oD2H + tid];
    }
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float *x = s + (ch ? oRawQ : oRawI);
      float *y = s + (ch ? oD1Q : oD1I) + 45;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int m = tid + kNT * r;
        const float4 *w4 = reinterpret_cast<const float4 *>(x + 4 * m);
        float acc = 0.0f;
#pragma unroll
        for (int q = 0; q < kDec1Taps / 4; ++q) {
          const float4 v = w4[q];
          acc = fmaf(v.x, taps[4 * q + 0], acc);
          acc = fmaf(v.y, taps[4 * q + 1], acc);
          acc = fmaf(v.z, taps[4 * q + 2], acc);
          acc = fmaf(v.w, taps[4 * q + 3], acc);
        }
        y[m] = acc;
      }
    }
  }
}

/* dec2, level adjust (Process.cpp:482-492), overlap-save assembly (Process.cpp:498-522) */
T41RX_DEV void PhDec2(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamCfg &cf = c.a.cfg[c.s0 + g];
    StreamState &st = c.a.st[c.s0 + g];
    /* dec1's input region is about to be overlaid: keep its last 27 samples */
    if (tid < 54) {
      const int ch = tid / 27, i = tid % 27;
      s[oD1H + tid] = s[(ch ? oRawQ : oRawI) + kBlock + i];
    }
    float acc[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float2 *w2 = reinterpret_cast<const float2 *>(s + (ch ? oD1Q : oD1I) + 2 * tid);
      float a = 0.0f;
#pragma unroll
      for (int q = 0; q < kDec2Taps / 2; ++q) {
        const float2 v = w2[q];
        a = fmaf(v.x, s[oTaps + kTapDec2 + 2 * q], a);
        a = fmaf(v.y, s[oTaps + kTapDec2 + 2 * q + 1], a);
      }
      acc[ch] = a;
    }
    float2 *fa = reinterpret_cast<float2 *>(s + vFftA);
    if (cf.mode == kModePsk31) {
      s[vAud + 23 + tid] = acc[0];                 /* Process.cpp:376-387,745: raw decimated I */
    } else if (cf.mode == kModeNfm) {
      fa[kDec + tid] = float2{acc[0], acc[1]};     /* Process.cpp:272-275 */
    } else {
      const float li = acc[0] * cf.vol_scale, lq = acc[1] * cf.vol_scale;
      float2 prev = float2{s[oOla + tid], s[oOla + 256 + tid]};
      if (st.first_block) prev = float2{0.0f, 0.0f};
      fa[tid] = prev;
      fa[kDec + tid] = float2{li, lq};
      s[oOla + tid] = li;
      s[oOla + 256 + tid] = lq;
    }
  }
}

/* dec2 history, NFM discriminator (Demod.cpp:220-235, Process.cpp:716-727,768-779) */
T41RX_DEV void PhPostDec2(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamCfg &cf = c.a.cfg[c.s0 + g];
    StreamState &st = c.a.st[c.s0 + g];
    if (tid < 90) {
      const int ch = tid / 45, i = tid % 45;
      s[oD2H + tid] = s[(ch ? oD1Q : oD1I) + kDec1Out + i];
    }
    if (tid == 0 && cf.mode != kModePsk31 && cf.mode != kModeNfm) st.first_block = 0;
    if (cf.mode == kModeNfm) {
      const float2 *fa = reinterpret_cast<const float2 *>(s + vFftA);
      /* fmdemod_quadri_K is a double literal (Demod.h:7): K * num / den runs in double */
      const double kq = 0.340447550238101026565118445432744920253753662109375;
      const float2 now = fa[kDec + tid];
      const float den = now.x * now.x + now.y * now.y;
      float out;
      if (tid == 0) {
        const float li = st.nfm_last_i, lq = st.nfm_last_q;
        const float num = now.x * (now.y - lq) - now.y * (now.x - li);
        out = (float)(kq * (double)num / (double)den);
      } else {
        const float2 last = fa[kDec + tid - 1];
        const float num = now.y * last.x - now.x * last.y;
        out = (float)(kq * (double)num / (double)den);
        out = (1.0f < out) ? 1.0f : out;           /* limiter skips index 0 (B5) */
        out = (-1.0f > out) ? -1.0f : out;
      }
      s[vAmTmp + tid] = out;
    }
  }
}

/* NFM: build the real-input overlap-save buffer from the demodulated audio */
T41RX_DEV void PhNfmAssemble(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamCfg &cf = c.a.cfg[c.s0 + g];
    if (cf.mode != kModeNfm) continue;
    StreamState &st = c.a.st[c.s0 + g];
    float2 *fa = reinterpret_cast<float2 *>(s + vFftA);
    if (tid == 0) {                                /* "last sample" = complex sample 127 (B4) */
      st.nfm_last_i = fa[kDec + 127].x;
      st.nfm_last_q = fa[kDec + 127].y;
    }
  }
}
T41RX_DEV void PhNfmAssemble2(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamCfg &cf = c.a.cfg[c.s0 + g];
    if (cf.mode != kModeNfm) continue;
    float2 *fa = reinterpret_cast<float2 *>(s + vFftA);
    const float a = s[vAmTmp + tid];
    fa[tid] = float2{s[oOla + tid], 0.0f};
    fa[kDec + tid] = float2{a, 0.0f};
    s[oOla + tid] = a;
  }
}

/* ------------------------------------------------------------------ */
/* P6-P8: fast convolution (Process.cpp:535-595,787-808)                */
/* ------------------------------------------------------------------ */
T41RX_DEV bool UsesFilter(int mode) { return mode != kModePsk31; }

T41RX_DEV void PhFftPass(Cta &c, int tid, int which, int pass) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  if (!UsesFilter(c.a.cfg[c.s0 + g].mode)) return;
  Radix8Butterfly(reinterpret_cast<float2 *>(Slot(c, g) + (which ? vFftB : vFftA)), c.a.twiddle, pass, u);
}

/* digit-reverse the forward result, multiply by the mask, conjugate for the inverse */
T41RX_DEV void PhMask(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  if (!UsesFilter(cf.mode)) return;
  float *s = Slot(c, g);
  const float2 *fa = reinterpret_cast<const float2 *>(s + vFftA);
  float2 *fb = reinterpret_cast<float2 *>(s + vFftB);
  const float2 *mask = reinterpret_cast<const float2 *>(c.a.fsets[cf.filter_id].mask);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = u + 64 * j;
    const float2 x = fa[OctRev3((unsigned)k)];
    const float2 h = LdgRO(mask + k);
    const float rr = x.x * h.x, ii = x.y * h.y, ri = x.x * h.y, ir = x.y * h.x;
    fb[k] = float2{rr - ii, -(ri + ir)};
  }
}

/* ------------------------------------------------------------------ */
/* P9-P12: AGC (DSP_Fn.cpp:479-632)                                     */
/* ------------------------------------------------------------------ */
/* undo the inverse transform's conjugate/scale for the 256 valid outputs; AGC off: x20;
 * AGC on: build the 97-sample-delayed views and |z|                                       */
T41RX_DEV void PhAgcPre(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  if (!UsesFilter(cf.mode)) return;
  float *s = Slot(c, g);
  const float2 *fb = reinterpret_cast<const float2 *>(s + vFftB);
  float2 *zext = reinterpret_cast<float2 *>(s + vZext);
  const float inv = 1.0f / 512.0f;
  float2 z[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = u + 64 * j;
    const float2 v = fb[OctRev3((unsigned)(kDec + i))];
    z[j] = float2{v.x * inv, -v.y * inv};
  }
  if (cf.agc_mode == 0) {
    float2 *dem = reinterpret_cast<float2 *>(s + vDem);   /* overlays FFT_A only: safe while FFT_B is read */
#pragma unroll
    for (int j = 0; j < 4; ++j) dem[u + 64 * j] = float2{cf.agc.fixed_gain * z[j].x, cf.agc.fixed_gain * z[j].y};
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = u + 64 * j;
    zext[kAgcDelay + i] = z[j];
    s[vAbs + kAgcDelay + i] = sqrtf(z[j].x * z[j].x + z[j].y * z[j].y);
  }
  /* history: sample -k (k = 1..97) sits at ring index 128 - k */
  for (int e = u; e < kAgcDelay; e += 64) {
    const int r = kAgcRing - (kAgcDelay - e);
    zext[e] = float2{s[oAgc + r], s[oAgc + 128 + r]};
    s[vAbs + e] = s[oAgc + 256 + r];
  }
}

/* sliding maximum over 97 entries by doubling: level L holds max over 2^L trailing entries */
T41RX_DEV void PhAgcMaxLevel(Cta &c, int tid, int level) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  if (!UsesFilter(cf.mode) || cf.agc_mode == 0) return;
  float *s = Slot(c, g);
  const int n = kAgcDelay + kDec;   /* 353 */
  if (level <= 6) {
    const float *src = s + (level == 1 ? vAbs : ((level & 1) ? vLvlB : vLvlA));
    float *dst = s + ((level & 1) ? vLvlA : vLvlB);
    const int d = 1 << (level - 1);
    if (level == 1) {
      /* the reference's `if (abs > ring_max)` never lets a NaN magnitude (NFM discriminator on
         exact silence: 0/0) become the maximum: NaNs count as 0 here */
      for (int e = u; e < n; e += 64) {
        const float a = src[e], b = (e >= 1) ? src[e - 1] : 0.0f;
        dst[e] = fmaxf((a == a) ? a : 0.0f, (b == b) ? b : 0.0f);
      }
    } else {
      for (int e = u; e < n; e += 64) dst[e] = (e >= d) ? fmaxf(src[e], src[e - d]) : src[e];
    }
  } else {
    /* level 6 result lives in vLvlB; window [i+1, i+97] = [e-96, e] with e = i + 97 */
    const float *m6 = s + vLvlB;
    for (int i = u; i < kDec; i += 64) {
      const int e = i + kAgcDelay;
      s[vRm + i] = fmaxf(m6[e], m6[e - 33]);
    }
  }
}

/* the serial envelope state machine: one lane per receiver */
T41RX_DEV void PhAgcSerial(Cta &c, int tid) {
  if (tid >= c.ng) return;
  const int g = tid;
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  if (!UsesFilter(cf.mode) || cf.agc_mode == 0) return;
  float *s = Slot(c, g);
  StreamState &st = c.a.st[c.s0 + g];
  const AgcConsts &a = cf.agc;
  float fast = st.agc_fast_back, hang = st.agc_hang_back, v = st.agc_volts, save = st.agc_save_volts;
  int hc = st.agc_hang_counter, state = st.agc_state, dtype = st.agc_decay_type, action = st.agc_action;
  float rm = st.agc_ring_max;
  for (int i = 0; i < kDec; ++i) {
    const float abs_out = s[vAbs + i];
    fast = a.fast_backmult * abs_out + a.onemfast_backmult * fast;
    hang = a.hang_backmult * abs_out + a.onemhang_backmult * hang;
    rm = s[vRm + i];
    if (hc > 0) --hc;
    if (rm >= v) {
      if (state >= 2) save = v;
      state = 0;
      v += (rm - v) * a.attack_mult;
    } else {
      switch (state) {
        case 0:
          if (v > a.pop_ratio * fast) {
            state = 1;
            v += (rm - v) * a.fast_decay_mult;
          } else if (a.hang_enable && (hang > a.hang_level)) {
            state = 2;
            hc = a.hang_counter_load;
            dtype = 1;
          } else {
            state = 3;
            v += (rm - v) * a.decay_mult;
            dtype = 0;
          }
          break;
        case 1:
          if (v > save) {
            v += (rm - v) * a.fast_decay_mult;
          } else if (hc > 0) {
            state = 2;
          } else if (dtype == 0) {
            state = 3;
            v += (rm - v) * a.decay_mult;
          } else {
            state = 4;
            v += (rm - v) * a.hang_decay_mult;
          }
          break;
        case 2:
          if (hc == 0) {
            state = 4;
            v += (rm - v) * a.hang_decay_mult;
          }
          break;
        case 3: {
          const float step = (rm - v) * a.decay_mult;
          v = (float)((double)v + (double)step * .05);   /* double product and sum (DSP_Fn.cpp:607) */
          break;
        }
        default:
          v += (rm - v) * a.hang_decay_mult;
          break;
      }
    }
    if (v < a.min_volts) {
      v = a.min_volts;
      action = 0;
    } else {
      action = 1;
    }
    s[vVolt + i] = v;
  }
  st.agc_fast_back = fast;
  st.agc_hang_back = hang;
  st.agc_volts = v;
  st.agc_save_volts = save;
  st.agc_ring_max = rm;
  st.agc_hang_counter = hc;
  st.agc_state = state;
  st.agc_decay_type = dtype;
  st.agc_action = action;
}

/* gain from volts (DSP_Fn.cpp:628) applied to the delayed samples; refresh the delay line */
T41RX_DEV void PhAgcPost(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  if (!UsesFilter(cf.mode) || cf.agc_mode == 0) return;
  float *s = Slot(c, g);
  const AgcConsts &a = cf.agc;
  const float2 *zext = reinterpret_cast<const float2 *>(s + vZext);
  float2 *dem = reinterpret_cast<float2 *>(s + vDem);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = u + 64 * j;
    const float v = s[vVolt + i];
    const double lg = (double)Log10Fast(a.inv_max_input * v);
    const double clipped = (0.0 < lg) ? 0.0 : lg;
    const float mult = (float)(((double)a.out_target - (double)a.slope_constant * clipped) / (double)v);
    const float2 o = zext[i];
    dem[i] = float2{o.x * mult, o.y * mult};
  }
  /* ring index r holds new sample 128 + r of this block = zext[97 + 128 + r] */
  for (int r = u; r < kAgcRing; r += 64) {
    const float2 z = zext[kAgcDelay + 128 + r];
    s[oAgc + r] = z.x;
    s[oAgc + 128 + r] = z.y;
    s[oAgc + 256 + r] = s[vAbs + kAgcDelay + 128 + r];
  }
}

/* ------------------------------------------------------------------ */
/* P13: demodulators (Process.cpp:615-761, Demod.cpp:40-139)            */
/* ------------------------------------------------------------------ */
T41RX_DEV void PhDemodParallel(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  if (!UsesFilter(cf.mode)) return;
  float *s = Slot(c, g);
  const float2 *dem = reinterpret_cast<const float2 *>(s + vDem);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = u + 64 * j;
    const float2 z = dem[i];
    if (cf.mode == kModeAm) s[vAmTmp + i] = AlphaBetaMag(z.x, z.y);
    else if (cf.mode != kModeSam) s[vAud + 23 + i] = z.x;   /* USB / LSB / NFM: real part */
  }
}

T41RX_DEV void PhDemodSerial(Cta &c, int tid) {
  if (tid >= c.ng) return;
  const int g = tid;
  const StreamCfg &cf = c.a.cfg[c.s0 + g];
  float *s = Slot(c, g);
  StreamState &st = c.a.st[c.s0 + g];
  const float2 *dem = reinterpret_cast<const float2 *>(s + vDem);
  if (cf.mode == kModeAm) {
    /* DC removal (Process.cpp:698-704) then 1-stage DF1 low-pass (Process.cpp:705) */
    float wold = st.am_wold;
    float x1 = st.am_lp_state[0], x2 = st.am_lp_state[1], y1 = st.am_lp_state[2], y2 = st.am_lp_state[3];
    const float b0 = cf.am_lp[0], b1 = cf.am_lp[1], b2 = cf.am_lp[2], a1 = cf.am_lp[3], a2 = cf.am_lp[4];
    for (int i = 0; i < kDec; ++i) {
      const float w = s[vAmTmp + i] + wold * 0.99f;
      const float x = w - wold;
      wold = w;
      float acc = b0 * x;
      acc = acc + b1 * x1;
      acc = acc + b2 * x2;
      acc = acc + a1 * y1;
      acc = acc + a2 * y2;
      x2 = x1; x1 = x;
      y2 = y1; y1 = acc;
      s[vAud + 23 + i] = acc;
    }
    st.am_wold = wold;
    st.am_lp_state[0] = x1; st.am_lp_state[1] = x2; st.am_lp_state[2] = y1; st.am_lp_state[3] = y2;
  } else if (cf.mode == kModeSam) {
    /* PLL constants of Demod.cpp:13-18 (omegaN = 200, zeta = 0.65, pll_fmax = 4000), host-computed */
    const float tpi = 6.283185307179586476925286766559f;
    const float omega_min = LdgRO(c.a.sam_consts + 0);
    const float omega_max = LdgRO(c.a.sam_consts + 1);
    const float g1 = LdgRO(c.a.sam_consts + 2);
    const float g2 = LdgRO(c.a.sam_consts + 3);
    float phz = st.sam_phzerror, fil = st.sam_fil_out, om2 = st.sam_omega2;
    for (int i = 0; i < kDec; ++i) {
      const float2 z = dem[i];
      const float sn = TableTurns(c.a.sin_table, phz * 0.159154943092f);
      const float cs = TableTurns(c.a.sin_table, phz * 0.159154943092f + 0.25f);
      const float ai = cs * z.x, bi = sn * z.x, aq = cs * z.y, bq = sn * z.y;
      const float corr0 = +ai + bq;
      const float corr1 = -bi + aq;
      float audio = (ai - bi) + (aq + bq);
      /* fade leveller with mtau = exp(0) = 1, onem = 0, dc = dc_insert = 0 at block start
         (locals, B3): dc = 1*dc + 0*audio; dc_insert = 1*dc_insert + 0*corr0; audio + dc_insert - dc.
         dc and dc_insert stay +0 for finite input, so audio + 0 - 0 leaves audio's bits except
         for -0 -> +0; keep that. */
      audio = (audio + 0.0f) - 0.0f;
      s[vAud + 23 + i] = audio;
      const float det = Atan2Approx(corr1, corr0);
      const float del_out = fil;
      om2 = om2 + g2 * det;
      if (om2 < omega_min) om2 = omega_min;
      else if (om2 > omega_max) om2 = omega_max;
      fil = g1 * det + om2;
      phz = phz + del_out;
      while (phz >= tpi) phz -= tpi;
      while (phz < 0.0f) phz += tpi;
    }
    st.sam_phzerror = phz;
    st.sam_fil_out = fil;
    st.sam_omega2 = om2;
  }
  /* PSK31 tap (psk31.cpp:235-310): first filtered sample of every third block */
  if (cf.psk31_enable && cf.mode != kModeNfm && cf.mode != kModePsk31) {
    int8_t bit_out = -1;
    uint8_t char_out = 0;
    if (st.psk_block_count % 3u == 0u) {
      const double pi_d = 3.1415926535897932384626433832795;
      const float2 z = dem[0];
      const float phase = Atan2Approx(z.y, z.x);
      float dphase = phase - st.psk_last_phase;
      while ((double)dphase < -pi_d) dphase = (float)((double)dphase + 2 * pi_d);
      while ((double)dphase >= pi_d) dphase = (float)((double)dphase - 2 * pi_d);
      const uint8_t bit = (((double)dphase > (pi_d / 2)) || ((double)dphase < (-pi_d / 2))) ? 0 : 1;
      st.psk_last_phase = phase;
      bit_out = (int8_t)bit;
      unsigned long long shr = (st.psk_shr << 1) | (unsigned long long)bit;
      if ((shr & 0xFFFull) != 0) {
        for (int i = 0; i < 128; ++i) {
          const uint32_t e = LdgRO(c.a.varicode + i);
          const unsigned long long want = ((unsigned long long)(e & 0xFFFFu)) << 2;
          const unsigned nbits = (((e >> 16) & 0xFFu) + 4u) & 63u;
          const unsigned long long keep = (nbits == 0) ? 0ull : (~0ull >> (64u - nbits));
          if (want == (shr & keep)) {
            shr = 0;
            char_out = (uint8_t)(e >> 24);
            break;
          }
        }
      }
      st.psk_shr = shr;
    }
    st.psk_block_count++;
    const size_t o = (size_t)(c.s0 + g) * c.a.n_blocks + c.t;
    if (c.a.psk_bits) c.a.psk_bits[o] = bit_out;
    if (c.a.psk_chars) c.a.psk_chars[o] = char_out;
  } else {
    const size_t o = (size_t)(c.s0 + g) * c.a.n_blocks + c.t;
    if (c.a.psk_bits) c.a.psk_bits[o] = -1;
    if (c.a.psk_chars) c.a.psk_chars[o] = 0;
  }
}

/* ------------------------------------------------------------------ */
/* P14/P15: arm_fir_interpolate_f32 x2 (48 taps) and x4 (32 taps), volume */
/* (Process.cpp:917-931)                                                 */
/* ------------------------------------------------------------------ */
T41RX_DEV void PhInterp1(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    if (tid < 23) s[vAud + tid] = s[oIntH + tid];
  }
}
T41RX_DEV void PhInterp1b(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const float *w = s + vAud + tid;          /* oldest-first window of 24 ending at input tid */
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
    for (int k = 0; k < 24; ++k) {
      const float x = w[k];
      a0 = fmaf(x, s[oTaps + kTapInt1 + 2 * k + 1], a0);   /* phase 0: c[(L-1) + kL] */
      a1 = fmaf(x, s[oTaps + kTapInt1 + 2 * k], a1);       /* phase 1: c[0 + kL]     */
    }
    s[vInt2 + 7 + 2 * tid] = a0;
    s[vInt2 + 7 + 2 * tid + 1] = a1;
    if (tid < 7) s[vInt2 + tid] = s[oIntH + 24 + tid];
  }
}

T41RX_DEV void PhInterp2(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamCfg &cf = c.a.cfg[c.s0 + g];
    float4 *dst = reinterpret_cast<float4 *>(c.a.audio + ((size_t)(c.s0 + g) * c.a.n_blocks + c.t) * kBlock);
    if (tid < 23) s[oIntH + tid] = s[vAud + kDec + tid];          /* int1 history for the next block */
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int n = tid + kNT * r;
      const float *w = s + vInt2 + n;
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float x = w[k];
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[p] = fmaf(x, s[oTaps + kTapInt2 + 4 * k + (3 - p)], acc[p]);
      }
      dst[n] = float4{acc[0] * cf.volume, acc[1] * cf.volume, acc[2] * cf.volume, acc[3] * cf.volume};
    }
  }
}

/* int2 history + Codec_gain (Process.cpp:979-1016 with the clip flags never set) */
T41RX_DEV void PhBlockEnd(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    if (tid < 7) s[oIntH + 24 + tid] = s[vInt2 + 2 * kDec + tid];
    if (tid == 32) {
      StreamState &st = c.a.st[c.s0 + g];
      uint32_t timer = st.codec_timer + 1;
      if (timer > 10000) timer = 10000;
      if (timer >= 50) {
        int rg = st.rf_gain + 1;
        if (rg > 15) rg = 15;
        st.rf_gain = rg;
        timer = 0;
      }
      st.codec_timer = timer;
    }
  }
}

/* ------------------------------------------------------------------ */
/* the block schedule; RX_PHASE(stmt) runs stmt for every tid then syncs */
/* ------------------------------------------------------------------ */
#define T41RX_BLOCK_SCHEDULE(RX_PHASE)                                   \
  RX_PHASE(PhLoad(c, tid));                                              \
  RX_PHASE(PhDcWarm(c, tid));                                            \
  RX_PHASE(PhDcMain(c, tid));                                            \
  RX_PHASE(PhDcVerify(c, tid));                                          \
  if (c.row) {                                                           \
    RX_PHASE(PhZoomIir(c, tid));                                         \
    RX_PHASE(PhSpecWindow(c, tid));                                      \
    RX_PHASE(PhSpecFftPass(c, tid, 0));                                  \
    RX_PHASE(PhSpecFftPass(c, tid, 1));                                  \
    RX_PHASE(PhSpecFftPass(c, tid, 2));                                  \
    RX_PHASE(PhSpecRow(c, tid));                                         \
  }                                                                      \
  RX_PHASE(PhNcoPrep(c, tid));                                           \
  RX_PHASE(PhMix(c, tid));                                               \
  RX_PHASE(PhNcoAdvance(c, tid); PhDec1(c, tid));                        \
  RX_PHASE(PhDec2(c, tid));                                              \
  RX_PHASE(PhPostDec2(c, tid));                                          \
  RX_PHASE(PhNfmAssemble(c, tid));                                       \
  RX_PHASE(PhNfmAssemble2(c, tid));                                      \
  RX_PHASE(PhFftPass(c, tid, 0, 0));                                     \
  RX_PHASE(PhFftPass(c, tid, 0, 1));                                     \
  RX_PHASE(PhFftPass(c, tid, 0, 2));                                     \
  RX_PHASE(PhMask(c, tid));                                              \
  RX_PHASE(PhFftPass(c, tid, 1, 0));                                     \
  RX_PHASE(PhFftPass(c, tid, 1, 1));                                     \
  RX_PHASE(PhFftPass(c, tid, 1, 2));                                     \
  RX_PHASE(PhAgcPre(c, tid));                                            \
  RX_PHASE(PhAgcMaxLevel(c, tid, 1));                                    \
  RX_PHASE(PhAgcMaxLevel(c, tid, 2));                                    \
  RX_PHASE(PhAgcMaxLevel(c, tid, 3));                                    \
  RX_PHASE(PhAgcMaxLevel(c, tid, 4));                                    \
  RX_PHASE(PhAgcMaxLevel(c, tid, 5));                                    \
  RX_PHASE(PhAgcMaxLevel(c, tid, 6));                                    \
  RX_PHASE(PhAgcMaxLevel(c, tid, 7));                                    \
  RX_PHASE(PhAgcSerial(c, tid));                                         \
  RX_PHASE(PhAgcPost(c, tid));                                           \
  RX_PHASE(PhDemodParallel(c, tid); PhInterp1(c, tid));                  \
  RX_PHASE(PhDemodSerial(c, tid));                                       \
  RX_PHASE(PhInterp1b(c, tid));                                          \
  RX_PHASE(PhInterp2(c, tid));                                           \
  RX_PHASE(PhBlockEnd(c, tid));

}  // namespace t41rx
#endif
