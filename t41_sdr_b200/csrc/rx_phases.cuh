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
#define T41RX_DEV_NOINLINE inline
template <class T> inline T LdgRO(const T *p) { return *p; }
#else
#define T41RX_DEV __device__ __forceinline__
#define T41RX_DEV_NOINLINE static __device__ __noinline__
template <class T> __device__ __forceinline__ T LdgRO(const T *p) { return __ldg(p); }
#endif

/* hint: pull one 128-byte line towards L2 (next block's I/Q while this block is computed) */
T41RX_DEV void PrefetchL2(const void *p) {
#ifndef T41RX_HOST_EMUL
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

namespace t41rx {

#ifndef T41RX_G
#define T41RX_G 4
#endif
constexpr int kG = T41RX_G;           /* receivers per CTA */
constexpr int kNT = 64 * kG;          /* threads per CTA: 64 per receiver in the FFT phases */
constexpr int kDcChunks = 64;         /* time-parallel chunks of the 4096-step DC-block chain: both warps of a receiver */
constexpr int kDcHalf = kDcChunks / 2;   /* chunks 0..31 cut the I block, 32..63 the Q block: no chunk reaches across sample 2048 */
constexpr int kDcChunkLen = 65;       /* a chunk starts 65 after its neighbour (odd stride: the lanes of a warp hit 32 different
                                         banks); the last chunk of a channel is the short one */
#ifndef T41RX_DC_SPEC_WARM
#define T41RX_DC_SPEC_WARM 192
#endif
constexpr int kDcSpecWarm = T41RX_DC_SPEC_WARM;      /* warm-up of a chunk's speculative start state (a1 = 0.854: 0.854^192 ~ 7e-14, a
                                         millionth of a float's last place; every start state is VERIFIED bit for bit
                                         against the previous chunk's end state and redone serially on a mismatch) */
constexpr int kDcWarm = 256;          /* warm-up where nothing can verify it (the rows-only kernel's block-start state):
                                         0.854^256 ~ 3e-18 */
static_assert(kDcHalf * kDcChunkLen >= kBlock && (kDcHalf - 1) * kDcChunkLen < kBlock, "chunking");
static_assert(kDcSpecWarm <= kDcWarm && kDcSpecWarm > kDcChunkLen, "bridge size");
/* the warm-ups that reach across the seam of the sequence (sample 2048: the I block ends, the Q block begins) are those of
   the Q chunks 1 .. (W - 1) / 65; together they read the samples [kDcBridgeLo, kDcBridgeHi) of the sequence */
constexpr int kDcBridgeLo = kBlock - kDcSpecWarm + kDcChunkLen;
constexpr int kDcBridgeHi = kBlock + ((kDcSpecWarm - 1) / kDcChunkLen) * kDcChunkLen;
constexpr int kDcBridgeLen = kDcBridgeHi - kDcBridgeLo;                 /* 257 at W = 192 */

/* ---- shared-memory slot layout, in floats ---- */
constexpr int kRawLen = 2076;                     /* 27 history + 2048 new (+1 pad) */
constexpr int oRawI = 0;
constexpr int oRawQ = oRawI + kRawLen;            /* 2076 */
#ifndef T41RX_ROWS_LAYOUT
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
constexpr int kMiscLen = 160;
constexpr int kSlotRaw = oMisc + kMiscLen;
#else
/* The rows-only kernel's slot (rx_rows.cu is compiled with T41RX_ROWS_LAYOUT): the raw tile, the DC-block scratch, the
   scalars: 19.1 KB instead of 28.8, so that THREE of its CTAs share an SM (its cascade phase is one warp's 2093-step
   chain while the CTA's other warps wait: rows in flight per SM are what its throughput is made of).  The regions of
   the audio chain are aliased onto the scratch: no phase of this kernel touches them (PhLoad's copy of the dec1 history
   reads 54 floats of the scalar area and puts them in front of the tile, where nothing looks). */
constexpr int oD1I = oRawQ + kRawLen;             /* 4152 : chunk states (256) + bridge (257) */
constexpr int kD1Len = (4 * kDcChunks + kDcBridgeLen + 2) / 2;
constexpr int oD1Q = oD1I, oOla = oD1I, oTaps = oD1I, oAgc = oD1I, oNco = oD1I, oD2H = oD1I, oIntH = oD1I;
constexpr int kTapDec1 = 0, kTapDec2 = 28, kTapInt1 = 74, kTapInt2 = 122;
constexpr int oMisc = oD1I + 2 * kD1Len;
constexpr int oD1H = oMisc;
constexpr int kMiscLen = 96;
constexpr int kSlotRaw = oMisc + kMiscLen;
#endif
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
constexpr int vSamSin = 1056;                     /* SAM: arm_sin_f32's 513-entry table, staged for the serial PLL (the
                                                     kernel has next to no L1: a table read from L2 costs ~300 clk, and
                                                     the PLL makes 4 dependent ones per sample) -> 1569 */
/* receive equaliser, between the demodulator and the interpolators: 14 band outputs of 256 samples; bands 0..11
   behind the audio buffer (over vAmTmp / vInt2, dead at that point), bands 12 and 13 in the D1 region */
constexpr int vEqBand = 1056;
static_assert(vEqBand + 12 * kDec <= 2 * kRawLen, "equaliser bands fit the raw region");
/* spectrum scratch (row blocks, before dec1): the D1 region */
#ifndef T41RX_ROWS_LAYOUT
constexpr int vSpecFft = oD1I;                    /* 512 complex = 1024 floats <= 1120 */
#else
/* rows-only layout: the top of the Q tile (a zoomed receiver's tile is dead behind the decimator; one at zoom x1 reads
   samples 0..511 of both tiles for this buffer: PhSpecWindow) */
constexpr int vSpecFft = 2 * kRawLen - 2 * kSpecRes;
static_assert(vSpecFft % 4 == 0 && vSpecFft >= oRawQ + 27 + kSpecRes, "spectrum scratch inside the Q tile, clear of its first 512 samples");
#endif

/* misc scalar indices */
enum { mDcD1 = 0, mDcD2 = 1, mNcoMode = 2, mDcBad = 3, mSid = 4 /* int: the receiver this slot serves */ };
constexpr int mCfg = 8;               /* from here: a copy of the receiver's StreamCfg (PhCtaInit) */
static_assert(sizeof(StreamCfg) % 8 == 0 && mCfg * 4 + sizeof(StreamCfg) <= kMiscLen * 4 && (oMisc + mCfg) % 2 == 0 && kSlot % 2 == 0,
              "the configuration record fits the scalar area, 8-byte aligned");
/* the chunks' speculative start states and end states, 64 x 2 each, during the DC phases only: the dec1 output region
   (its history is restored by PhDec1, the spectrum scratch of a row block is written after PhDcFix) */
constexpr int vDcSpec = oD1I, vDcEnd = oD1I + 2 * kDcChunks;
/* behind them, until PhDcFix: the samples either side of the seam of the sequence, I[kDcBridgeLo ..] then
   Q[.. kDcBridgeHi - 2049], as one run (PhLoad writes it beside the raw tile): the warm-up of the first Q chunks reads its
   I part and its Q part from here in one piece */
constexpr int vDcBridge = oD1I + 4 * kDcChunks;
static_assert(vDcBridge + kDcBridgeLen <= oD1I + 2 * kD1Len, "DC chunk states and the bridge fit the dec1 region");

struct LaunchArgs {
  const float *iq;
  float *audio;
  /* the firmware's q15 block format instead (both set, iq / audio NULL): int16 [receiver][block][2048][2] in,
     int16 [receiver][block][2048] out; every kernel converts at its own load / store (arm_q15_to_float = x / 32768,
     arm_float_to_q15 = saturate(trunc(x * 32768)); Process.cpp:107-108,936) */
  const int16_t *iq16;
  int16_t *audio16;
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
  const float *eq_coeffs;     /* 14 x 20: receive-equaliser band-pass biquads (FIR.cpp:279-371) */
  const float *cw_coeffs;     /* 5 x 30: CW audio low-passes (FIR.cpp:15-65) */
  const float *sam_consts;    /* omega_min, omega_max, g1, g2 (Demod.cpp:13-18) */
  const uint16_t *gradient;   /* 117 */
  const uint32_t *varicode;   /* 128: code | bits << 16 | ascii << 24 */
  int n_streams, n_blocks, row_every, n_rows;   /* n_streams: receivers this launch works on */
  /* a call may be cut into launches over block ranges: this launch covers blocks t0 .. t0 + n_blocks - 1 of a call
     whose buffers hold t_stride blocks per receiver (iq / audio / psk pointers are already offset by t0; row_every,
     n_rows and the row outputs refer to the whole call) */
  int t0, t_stride;
  uint32_t flags;
  const int32_t *stream_ids;  /* receiver index of each of them, or NULL for stream_base .. stream_base + n_streams-1 */
  int stream_base;
  /* audio-spectrum + S-meter by-product of the row-producing blocks (all NULL when not bound) */
  float2 *aspec;              /* scratch [receiver][n_rows][512]: the masked spectrum (iFFT_buffer before the inverse FFT) */
  int32_t *audio_ypixel;      /* [receiver][n_rows][270] */
  float *audio_max_ave;       /* [receiver][n_rows] */
  /* what the row-producing blocks write to the control app's serial port (NULL when not bound) */
  uint8_t *spec_frames;       /* [receiver][n_rows][518] */
  uint8_t *audio_frames;      /* [receiver][n_rows][270] */
  /* split form of the bit-exact chain (front kernel | serial kernel, thread = receiver | back kernel): hand-over
     through HBM, receiver-minor so that the serial kernel's lanes read and write consecutive words:
       ser_in  [n_blocks][256][n_streams]  (delayed z.re, delayed z.im, delayed |z|, window maximum)
       ser_out [n_blocks][256][n_streams]  demodulated audio sample */
  float4 *ser_in;
  float *ser_out;
  /* spectral noise reduction / noise blanker (rx_nr.cuh): per-receiver state (NULL until a receiver uses a stage) and
     the host-computed tables */
  NrState *nr;
  const float *nr_tab;
};

struct Cta {
  LaunchArgs a;
  float *smem;
  int s0;      /* first receiver of this CTA */
  int ng;      /* receivers handled (<= kG) */
  int t;       /* block index within the launch */
  int row;     /* this block produces a spectrum row */
  int row_idx;
  int rows_only;   /* 1 in t41rx_rows_kernel */
  int casc_warp;   /* rows kernel: the warp of the CTA that runs the ZoomFFT cascade */
  int dc_carried;  /* rows kernel: the slot holds the DC-block state the previous row block ended in (every block a row) */
};

T41RX_DEV float *Slot(const Cta &c, int g) { return c.smem + g * kSlot; }
/* receiver index of slot g of this CTA */
/* (a launch over an id list keeps the ids in the slots, PhCtaInit: with next to no L1 beside the shared memory a look-up
   in the list is a trip to L2, and the phases ask thousands of times per block; a contiguous range is plain arithmetic) */
T41RX_DEV int Sid(const Cta &c, int g) {
  return c.a.stream_ids ? reinterpret_cast<const int *>(c.smem + g * kSlot)[oMisc + mSid] : c.a.stream_base + c.s0 + g;
}
/* first thing every kernel of this file does, followed by a barrier */
T41RX_DEV void PhCtaInit(Cta &c, int tid) {
  if (tid < c.ng)
    reinterpret_cast<int *>(c.smem + tid * kSlot)[oMisc + mSid] =
        c.a.stream_ids ? LdgRO(c.a.stream_ids + c.s0 + tid) : c.a.stream_base + c.s0 + tid;
  /* the receivers' configuration records, beside the scalars: nearly every phase opens with a look at its receiver's
     mode or switches, and with next to no L1 beside the shared memory each such look was a trip to L2 (600 - 800 clocks
     in front of phases that take 1000 - 5000) */
  constexpr int kW = (int)(sizeof(StreamCfg) / sizeof(uint32_t));
  for (int i = tid; i < c.ng * kW; i += kNT) {
    const int g = i / kW, w = i - g * kW;
    const int sid = c.a.stream_ids ? LdgRO(c.a.stream_ids + c.s0 + g) : c.a.stream_base + c.s0 + g;
    reinterpret_cast<uint32_t *>(c.smem + g * kSlot + oMisc + mCfg)[w] = LdgRO(reinterpret_cast<const uint32_t *>(c.a.cfg + sid) + w);
  }
}
/* receiver g's configuration (the copy PhCtaInit left in the slot) */
T41RX_DEV const StreamCfg &CfgOf(const Cta &c, int g) {
  return *reinterpret_cast<const StreamCfg *>(c.smem + g * kSlot + oMisc + mCfg);
}

/* one word of a receiver's I/Q (float index `w` inside the [receiver][block][2048][2] array, block-relative base
   already applied): the float buffer, or arm_q15_to_float of the q15 one (x / 32768, exact) */
T41RX_DEV float IqWord(const LaunchArgs &a, size_t w) {
  return a.iq16 ? (float)LdgRO(a.iq16 + w) / 32768.0f : LdgRO(a.iq + w);
}
/* arm_float_to_q15 (Process.cpp:936): saturate((q31)(x * 32768)) to 16 bits, truncation toward zero; NaN -> 0 (the
   target's VCVT.S32.F32) */
T41RX_DEV int16_t FloatToQ15Word(float x) {
  const float v = x * 32768.0f;
  int q;
  if (v != v) q = 0;
  else if (!(v > -2147483648.0f)) q = INT32_MIN;
  else if (v >= 2147483648.0f) q = INT32_MAX;
  else q = (int)v;
  q = q > 32767 ? 32767 : q;
  q = q < -32768 ? -32768 : q;
  return (int16_t)q;
}

/* which receiver (or -1) thread `tid` serves in a one-lane-per-receiver serial phase:
 * T41RX_SERIAL_WPS = 1: lane 0 of the receiver's own first warp (no divergence between receivers
 * in different states, more issue slots); 0: lanes 0..G-1 of warp 0 (fewest issue slots). */
#ifndef T41RX_SERIAL_WPS
#define T41RX_SERIAL_WPS 1
#endif
T41RX_DEV int SerialStream(const Cta &c, int tid) {
#if T41RX_SERIAL_WPS
  return ((tid & 63) == 0 && (tid >> 6) < c.ng) ? (tid >> 6) : -1;
#else
  return (tid < c.ng) ? tid : -1;
#endif
}

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
  /* the reference's branches as selects around ONE division (the operands it would divide on the branch taken): the
     callers run this inside long serial chains, where a branch per sample keeps the compiler from overlapping the
     chain with anything else */
  const float pi = 3.1415926535897932384626433832795f;
  const float tpi = 6.283185307179586476925286766559f;
  const bool wide = fabsf(x) > fabsf(y);
  const float z = (wide ? y : x) / (wide ? x : y);        /* x != 0 and not wide: |y| >= |x| > 0 */
  const float p = AtanPoly(z);
  const float r_wide = (x > 0.0f) ? p : ((y >= 0.0f) ? p + pi : p - pi);
  const float r_tall = (y > 0.0f) ? -p + tpi : -p - tpi;
  const float r_axis = (y > 0.0f) ? tpi : ((y < 0.0f) ? -tpi : 0.0f);
  return (x != 0.0f) ? (wide ? r_wide : r_tall) : r_axis;
}

/* arm_sin_f32 / arm_cos_f32 tail: linear interpolation in the 513-entry table */
T41RX_DEV float TableTurns(const float *tab, float in) {
  /* arm_sin_f32: n = (int32_t) in; if (in < 0) n--; in -= (float) n.  floorf gives the same float except at the negative
     integers, where the reference is left with in = 1.0 -> findex = 512 -> wrapped to entry 0 with fraction 0: the same
     table entry and fraction as in = 0.  One rounding instruction instead of a conversion to integer and back on the
     serial chain of the SAM PLL; likewise for the fraction below (0 <= findex <= 512: floorf(findex) = (float) index) */
  in = in - floorf(in);
  float findex = 512.0f * in;
  uint32_t index = (uint32_t)findex & 0xFFFFu;
  float findex_floor = floorf(findex);
  if (index >= 512u) {
    index = 0;
    findex -= 512.0f;
    findex_floor = 0.0f;
  }
  const float fract = findex - findex_floor;
  const float a = tab[index];                     /* the table sits in shared memory (vSamSin) */
  const float b = tab[index + 1];
  const float wa = (1.0f - fract) * a;
  const float wb = fract * b;
  return wa + wb;
}

/* Process.cpp:165-174 + Utility.cpp:178-187 on one sample of the conditioned buffers.  The two
 * per-receiver values are fetched once (registers) so loops over samples do not re-read them. */
struct IqFix { bool mirrored; float phase; };
T41RX_DEV IqFix IqFixOf(const StreamCfg &cf) { return IqFix{cf.mirrored != 0, cf.iq_phase}; }
T41RX_DEV void IqCorr(const IqFix f, float &i, float &q) {
  if (f.mirrored) {
    if (f.phase < 0.0f) q = q + i * f.phase;
    else i = i + q * f.phase;
  }
}

/* FreqShift1 (Freq_Shift.cpp:42-65): multiply sample n by exp(+j*pi*n/2); branch-free
 * (n & 3 differs from lane to lane): 1: (-q, i)  2: (-i, -q)  3: (q, -i) */
T41RX_DEV void QuarterShift(int n, float &i, float &q) {
  const int k = n & 3;
  const float a = (k & 1) ? q : i;
  const float b = (k & 1) ? i : q;
  i = ((k + 1) & 2) ? -a : a;
  q = (k & 2) ? -b : b;
}

/* The DC-block biquad (arm_biquad_cascade_df2T_f32, 1 stage; coefficients FIR.cpp:87-89):
 *     y  = (b0*x) + d1;   d1 = ((b1*x) + (a1*y)) + d2;   d2 = (b2*x) + (a2*y)
 * with b2 = a2 = 0, so d2 is always +-0.  Adding +-0 changes nothing unless the other operand is
 * -0, and t = (b1*x) + (a1*y) can only be -0 if b1*x = -0 (x = +0, because b1 < 0) AND
 * a1*y = -0 (y = -0); but x = +0 gives y = (+0) + d1 which is never -0.  Hence d1 = t exactly
 * for every finite input and d2 is only materialised (from the last x and y) when a state is
 * stored.  Loop-carried chain per sample: FADD -> FMUL -> FADD. */
struct DcCoef { float b0, b1, a1; };
T41RX_DEV DcCoef DcCoefs() {
  return DcCoef{(float)0.927176191943378969, (float)-0.927176191943378969, (float)0.854352383886757938};
}
T41RX_DEV float DcStep(const DcCoef &k, float x, float &d1) {
  const float y = k.b0 * x + d1;
  d1 = k.b1 * x + k.a1 * y;
  return y;
}
/* d2 = (0*x) + (0*y) of the last processed sample */
T41RX_DEV float DcD2(float x, float y) { return 0.0f * x + 0.0f * y; }

/* raw sequence index (I block then Q block, B6) -> float offset inside the slot */
T41RX_DEV int SeqOff(int i) { return oRawI + 27 + i + ((i >= kBlock) ? (oRawQ - oRawI - kBlock) : 0); }

/* ------------------------------------------------------------------ */
/* launch prologue / epilogue: state <-> shared memory                 */
/* ------------------------------------------------------------------ */
T41RX_DEV void PhStateIn(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamState &st = c.a.st[Sid(c, g)];
    const StreamCfg &cf = CfgOf(c, g);
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
    const double *tab = c.a.nco_tab + (size_t)(Sid(c, g)) * 192;
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
    StreamState &st = c.a.st[Sid(c, g)];
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
      st.fast_native = 0;
    }
  }
}

/* ------------------------------------------------------------------ */
/* P0: HBM -> shared, de-interleave; restore dec1 history              */
/* ------------------------------------------------------------------ */
T41RX_DEV void PhLoad(Cta &c, int tid) {
  /* one complex sample per lane and load: a warp's 32 consecutive I (and Q) samples go to 32 different banks (two
     samples per lane would put every store on 16 banks) */
  constexpr int kPer = kBlock / kNT;             /* float2 loads per thread per receiver */
  static_assert(kPer * kNT == kBlock, "load tiling");
  for (int g0 = 0; g0 < c.ng; g0 += 2) {         /* two receivers' loads in flight at once */
    float2 v[2][kPer];
#pragma unroll
    for (int gg = 0; gg < 2; ++gg) {
      if (g0 + gg >= c.ng) continue;
      const size_t blk = ((size_t)(Sid(c, g0 + gg)) * c.a.t_stride + c.t) * (2 * kBlock);
      if (c.a.iq16) {
        const short2 *src = reinterpret_cast<const short2 *>(c.a.iq16 + blk);
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
          const short2 q = LdgRO(src + tid + kNT * k);
          v[gg][k] = float2{(float)q.x / 32768.0f, (float)q.y / 32768.0f};
        }
      } else {
        const float2 *src = reinterpret_cast<const float2 *>(c.a.iq + blk);
#pragma unroll
        for (int k = 0; k < kPer; ++k) v[gg][k] = LdgRO(src + tid + kNT * k);
      }
    }
#pragma unroll
    for (int gg = 0; gg < 2; ++gg) {
      if (g0 + gg >= c.ng) continue;
      float *s = Slot(c, g0 + gg);
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int n = tid + kNT * k;             /* sample index */
        s[oRawI + 27 + n] = v[gg][k].x;
        s[oRawQ + 27 + n] = v[gg][k].y;
        /* the DC bridge (vDcBridge) */
        if (n >= kDcBridgeLo) s[vDcBridge + n - kDcBridgeLo] = v[gg][k].x;
        if (n < kDcBridgeHi - kBlock) s[vDcBridge + kBlock - kDcBridgeLo + n] = v[gg][k].y;
      }
    }
  }
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    for (int h = tid; h < 54; h += kNT) {
      const int ch = h / 27, i = h % 27;
      s[(ch ? oRawQ : oRawI) + i] = s[oD1H + h];
    }
    if (c.t + 1 < c.a.n_blocks) {
      const size_t blk = ((size_t)(Sid(c, g)) * c.a.t_stride + c.t + 1) * (2 * kBlock);
      const char *nxt = c.a.iq16 ? reinterpret_cast<const char *>(c.a.iq16 + blk) : reinterpret_cast<const char *>(c.a.iq + blk);
      const int bytes = c.a.iq16 ? 2 * kBlock * 2 : 2 * kBlock * 4;
      for (int line = tid; line < bytes / 128; line += kNT) PrefetchL2(nxt + 128 * line);
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
/* Run steps [begin, end) of the conditioning chain from state (d1, d2).  kStore: write the
 * conditioned samples in place (x RFgain, then for I in mirrored modes x -IQAmp as a second,
 * separately rounded multiplication: Process.cpp:133,166).  Loads are issued in batches of 4
 * so that the shared-memory latency is off the recurrence's critical path. */
struct DcPost { float rfg; float rfgain; float neg_iq_amp; bool mirrored; };

#ifndef T41RX_DC_BATCH
#define T41RX_DC_BATCH 4   /* register batch of the chain: 4 keeps the front kernel (128-register cap) almost spill-free; a spill is an L2 access here */
#endif
/* n consecutive samples at x (one channel's part of the sequence: no address arithmetic per sample).  second: the
   I channel's extra factor -IQAmp in the mirrored modes */
template <bool kStore>
T41RX_DEV void DcSegment(float *x, int n, const DcCoef &k, const DcPost &p, bool second, float &d1, float &lx, float &ly) {
  constexpr int kB = T41RX_DC_BATCH;
  const int n_batches = n / kB;
  float cur[kB], nxt[kB];
  if (n_batches > 0) {
#pragma unroll
    for (int j = 0; j < kB; ++j) cur[j] = x[j];
  }
  int i = 0;
  for (int bt = 0; bt < n_batches; ++bt) {
    /* software pipeline: the next batch is in flight while this one runs from registers */
    if (bt + 1 < n_batches) {
#pragma unroll
      for (int j = 0; j < kB; ++j) nxt[j] = x[i + kB + j];
    }
#pragma unroll
    for (int j = 0; j < kB; ++j) {
      const float v = cur[j] * p.rfg;
      float y = DcStep(k, v, d1);
      lx = v;
      ly = y;
      if (kStore) {
        y = y * p.rfgain;
        if (second) y = y * p.neg_iq_amp;
        x[i + j] = y;
      }
    }
#pragma unroll
    for (int j = 0; j < kB; ++j) cur[j] = nxt[j];
    i += kB;
  }
  for (; i < n; ++i) {
    const float v = x[i] * p.rfg;
    float y = DcStep(k, v, d1);
    lx = v;
    ly = y;
    if (kStore) {
      y = y * p.rfgain;
      if (second) y = y * p.neg_iq_amp;
      x[i] = y;
    }
  }
}

/* chunk ch of the sequence: [DcChunkBegin, DcChunkEnd), all of it in one channel */
T41RX_DEV int DcChunkBegin(int ch) {
  return ch < kDcHalf ? ch * kDcChunkLen : kBlock + (ch - kDcHalf) * kDcChunkLen;
}
T41RX_DEV int DcChunkEnd(int ch) {
  const int e = DcChunkBegin(ch) + kDcChunkLen, lim = ch < kDcHalf ? kBlock : 2 * kBlock;
  return e < lim ? e : lim;
}
/* where sample i of the sequence lives in the slot (I block, then Q block) */
T41RX_DEV float *DcSample(float *s, int i) { return s + oRawI + 27 + i + (i >= kBlock ? oRawQ - oRawI - kBlock : 0); }

/* RFgain in force while block t of the launch is processed, from the values at launch start: Codec_gain
   (Process.cpp:979-1016 with the clip flags never set) raises it by one every 50 blocks up to 15 */
T41RX_DEV int RfGainAtBlock(int rf_gain0, uint32_t timer0, int t) {
  const int k = (int)((timer0 + (uint32_t)t) / 50u);          /* ticks of the 50-block timer so far */
  if (k == 0) return rf_gain0;
  const int rg = rf_gain0 + k;                                /* each tick: min(RFgain + 1, 15), also from above 15 */
  return rg > 15 ? 15 : rg;
}

T41RX_DEV DcPost DcPostOf(const Cta &c, int g) {
  const StreamCfg &cf = CfgOf(c, g);
  const StreamState &st = c.a.st[Sid(c, g)];
  /* rows_only: the state in HBM is the one at launch start (the throughput kernel advances it) */
  const int rg = c.rows_only ? RfGainAtBlock(st.rf_gain, st.codec_timer, c.t) : st.rf_gain;
  return DcPost{cf.rf_gain_value, (float)rg, cf.neg_iq_amp, cf.mirrored != 0};
}

T41RX_DEV void PhDcWarm(Cta &c, int tid) {
  if (tid >= c.ng * kDcChunks) return;
  const int g = tid / kDcChunks, ch = tid % kDcChunks;
  float *s = Slot(c, g);
  if (ch == 0) {
    s[oMisc + mDcBad] = 0.0f;
    return;
  }
  const DcPost p = DcPostOf(c, g);
  const int end = DcChunkBegin(ch);
  int start = end - kDcSpecWarm;
  float d1 = 0.0f;             /* (the state's second word is +-0 and never read: see DcStep) */
  if (start <= 0) {            /* the warm-up would reach before the block: start from the carried state instead */
    start = 0;
    d1 = s[oMisc + mDcD1];
  }
  /* one run of consecutive floats: in the tile, or (a warm-up across the seam) in the bridge */
  float *x = (start < kBlock && end > kBlock) ? s + vDcBridge + (start - kDcBridgeLo) : DcSample(s, start);
  const DcCoef k = DcCoefs();
  float lx = 0.0f, ly = 0.0f;
  DcSegment<false>(x, end - start, k, p, false, d1, lx, ly);
  s[vDcSpec + 2 * ch] = d1;
  s[vDcSpec + 2 * ch + 1] = DcD2(lx, ly);
}

T41RX_DEV void PhDcMain(Cta &c, int tid) {
  if (tid >= c.ng * kDcChunks) return;
  const int g = tid / kDcChunks, ch = tid % kDcChunks;
  float *s = Slot(c, g);
  const DcPost p = DcPostOf(c, g);
  float d1 = (ch == 0) ? s[oMisc + mDcD1] : s[vDcSpec + 2 * ch];
  const int begin = DcChunkBegin(ch), end = DcChunkEnd(ch);
  const DcCoef k = DcCoefs();
  float lx = 0.0f, ly = 0.0f;
  DcSegment<true>(DcSample(s, begin), end - begin, k, p, ch < kDcHalf && p.mirrored, d1, lx, ly);
  s[vDcEnd + 2 * ch] = d1;
  s[vDcEnd + 2 * ch + 1] = DcD2(lx, ly);
}

#ifndef T41RX_HOST_EMUL
/* how often a chunk's speculative start state missed (instrumentation: t41rx_dc_refilter_count) */
static __device__ unsigned long long g_dc_refilter_count;
#endif
T41RX_DEV bool SameBits(float a, float b) {
  union { float f; uint32_t u; } x, y;
  x.f = a;
  y.f = b;
  return x.u == y.u;
}

/* every chunk checks, bit for bit, that the state it started from is the state the previous
 * chunk ended in; the last chunk's end state becomes the receiver's carried state */
T41RX_DEV void PhDcVerify(Cta &c, int tid) {
  if (tid >= c.ng * kDcChunks) return;
  const int g = tid / kDcChunks, ch = tid % kDcChunks;
  float *s = Slot(c, g);
  if (ch > 0) {
    if (!SameBits(s[vDcSpec + 2 * ch], s[vDcEnd + 2 * (ch - 1)]) ||
        !SameBits(s[vDcSpec + 2 * ch + 1], s[vDcEnd + 2 * (ch - 1) + 1]))
      s[oMisc + mDcBad] = 1.0f;           /* several lanes may store the same value: benign */
  }
  if (ch == kDcChunks - 1) {
    s[oMisc + mDcD1] = s[vDcEnd + 2 * ch];
    s[oMisc + mDcD2] = s[vDcEnd + 2 * ch + 1];
  }
}

/* speculation missed (vanishingly rare): redo serially from the first bad chunk, re-reading the
 * raw samples from HBM/L2 because the chunks were filtered in place */
T41RX_DEV void PhDcFix(Cta &c, int tid) {
  const int g = SerialStream(c, tid);
  if (g < 0) return;
  float *s = Slot(c, g);
  if (s[oMisc + mDcBad] == 0.0f) return;
  const DcPost p = DcPostOf(c, g);
  int bad = 1;
  for (; bad < kDcChunks; ++bad) {
    if (!SameBits(s[vDcSpec + 2 * bad], s[vDcEnd + 2 * (bad - 1)]) ||
        !SameBits(s[vDcSpec + 2 * bad + 1], s[vDcEnd + 2 * (bad - 1) + 1]))
      break;
  }
  if (bad >= kDcChunks) return;
#ifndef T41RX_HOST_EMUL
  atomicAdd(&g_dc_refilter_count, 1ull);
#endif
  const size_t src = ((size_t)(Sid(c, g)) * c.a.t_stride + c.t) * (2 * kBlock);
  float d1 = s[vDcEnd + 2 * (bad - 1)];
  float d2 = s[vDcEnd + 2 * (bad - 1) + 1];
  const int from = DcChunkBegin(bad);
  for (int i = from; i < 2 * kBlock; ++i)
    s[SeqOff(i)] = IqWord(c.a, src + (i < kBlock ? 2 * i : 2 * (i - kBlock) + 1));
  const DcCoef k = DcCoefs();
  float lx = 0.0f, ly = 0.0f;
#ifndef T41RX_HOST_EMUL
#pragma unroll 1                          /* one copy of the loop body */
#endif
  for (int seg = (from < kBlock) ? 0 : 1; seg < 2; ++seg) {      /* the rest of the I block, then the Q block */
    const int b = seg ? (from > kBlock ? from : kBlock) : from, e = seg ? 2 * kBlock : kBlock;
    DcSegment<true>(DcSample(s, b), e - b, k, p, seg == 0 && p.mirrored, d1, lx, ly);
  }
  d2 = DcD2(lx, ly);
  s[oMisc + mDcD1] = d1;
  s[oMisc + mDcD2] = d2;
}

/* ------------------------------------------------------------------ */
/* P2: display spectrum on row-producing blocks (FFT.cpp:67-251)        */
/* ------------------------------------------------------------------ */
/* zoom index >= 1: 4-stage elliptic DF1 biquad cascade on the Fs/4-shifted I and Q, 4-tap
 * FIR decimation by 2^zoom, first zoom_samples outputs into the 512-deep ring.  One lane
 * per (receiver, channel); the cascade is evaluated sample by sample through all four
 * stages, which yields the same values as the reference's stage-by-stage order.           */
#ifdef T41RX_HOST_EMUL
/* host emulation: one lane per (receiver, channel) walks the four stages sample by sample */
T41RX_DEV void PhZoomIir(Cta &c, int tid) {
  if (!c.row || tid >= c.ng * 2) return;
  const int g = tid >> 1, chn = tid & 1;
  const StreamCfg &cf = CfgOf(c, g);
  if (cf.zoom == 0) return;
  StreamState &st = c.a.st[Sid(c, g)];
  const float *s = Slot(c, g);
  float k[20];
  for (int i = 0; i < 20; ++i) k[i] = LdgRO(c.a.zoom_iir + (cf.zoom - 1) * 20 + i);
  float zs[16];
  for (int i = 0; i < 16; ++i) zs[i] = st.zoom_iir[chn][i];
  float h0 = st.zoom_fir_hist[chn][0], h1 = st.zoom_fir_hist[chn][1], h2 = st.zoom_fir_hist[chn][2];
  const int M = 1 << cf.zoom;
  int ptr = st.zoom_ptr;
  int produced = 0;
  const IqFix fix = IqFixOf(cf);
  for (int n = 0; n < kBlock; ++n) {
    float vi = s[oRawI + 27 + n], vq = s[oRawQ + 27 + n];
    IqCorr(fix, vi, vq);
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

#else
/* device: eight lanes per receiver = (channel, biquad stage), software-pipelined with a skew of
 * two samples per stage; a stage's output travels to the next lane by warp shuffle one step
 * ahead of its use, so the shuffle latency hides behind the biquad's add chain.  Identical
 * arithmetic to the sample-by-sample form above. */
T41RX_DEV void PhZoomIir(Cta &c, int tid) {
  if (!c.row) return;
  const int g = tid >> 6, lane = tid & 63;
  if (g >= c.ng || lane >= 32) return;               /* first warp of the receiver's 64-thread group */
  const StreamCfg &cf = CfgOf(c, g);
  if (cf.zoom == 0) return;                          /* warp-uniform */
  StreamState &st = c.a.st[Sid(c, g)];
  const float *s = Slot(c, g);
  const bool active = lane < 8;
  const int chn = (lane >> 2) & 1, sg = lane & 3;
  float b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0, x1 = 0, x2 = 0, y1 = 0, y2 = 0;
  if (active) {
    const float *kk = c.a.zoom_iir + (cf.zoom - 1) * 20 + 5 * sg;
    b0 = LdgRO(kk); b1 = LdgRO(kk + 1); b2 = LdgRO(kk + 2); a1 = LdgRO(kk + 3); a2 = LdgRO(kk + 4);
    x1 = st.zoom_iir[chn][4 * sg]; x2 = st.zoom_iir[chn][4 * sg + 1];
    y1 = st.zoom_iir[chn][4 * sg + 2]; y2 = st.zoom_iir[chn][4 * sg + 3];
  }
  float h0 = 0, h1 = 0, h2 = 0;
  const bool last = active && sg == 3;
  if (last) { h0 = st.zoom_fir_hist[chn][0]; h1 = st.zoom_fir_hist[chn][1]; h2 = st.zoom_fir_hist[chn][2]; }
  const float f0 = cf.zoom_fir[0], f1 = cf.zoom_fir[1], f2 = cf.zoom_fir[2], f3 = cf.zoom_fir[3];
  const int M = 1 << cf.zoom;
  const int zs = cf.zoom_samples;
  int ptr = st.zoom_ptr;
  int produced = 0;
  float ylast = 0.0f, xcur = 0.0f;
  const IqFix fix = IqFixOf(cf);
  /* every lane runs the same instruction stream (selects instead of branches); the raw sample a
     stage-0 lane needs is fetched one step ahead */
  float ri = s[oRawI + 27], rq = s[oRawQ + 27];
  for (int k = 0; k < kBlock + 6; ++k) {
    /* my previous output is the next lane's input at step k+1 */
    const float xfetch = __shfl_up_sync(0xffffffffu, ylast, 1);
    const int n = k - 2 * sg;
    const int nn = (k + 1 < kBlock) ? k + 1 : kBlock - 1;
    const float ri_next = s[oRawI + 27 + nn], rq_next = s[oRawQ + 27 + nn];
    float vi = ri, vq = rq;                       /* sample k: what a stage-0 lane processes now */
    IqCorr(fix, vi, vq);
    QuarterShift(k, vi, vq);
    const float x0 = chn ? vq : vi;
    const float x = (sg == 0) ? x0 : xcur;
    const bool live = active && n >= 0 && n < kBlock;
    float acc = b0 * x;
    acc = acc + b1 * x1;
    acc = acc + b2 * x2;
    acc = acc + a1 * y1;
    acc = acc + a2 * y2;
    if (live) {
      x2 = x1; x1 = x;
      y2 = y1; y1 = acc;
      ylast = acc;
    }
    /* 4-tap FIR decimator on the last stage's output (arm_fir_decimate_f32) */
    float o = 0.0f;
    o = fmaf(h0, f0, o);
    o = fmaf(h1, f1, o);
    o = fmaf(h2, f2, o);
    o = fmaf(acc, f3, o);
    if (live && last) {
      if ((n & (M - 1)) == 0) {
        if (produced < zs) {
          st.zoom_ring[chn][ptr] = o;
          ptr = (ptr + 1 >= kSpecRes) ? 0 : ptr + 1;
        }
        ++produced;
      }
      h0 = h1; h1 = h2; h2 = acc;
    }
    xcur = xfetch;
    ri = ri_next;
    rq = rq_next;
  }
  if (active) {
    st.zoom_iir[chn][4 * sg] = x1; st.zoom_iir[chn][4 * sg + 1] = x2;
    st.zoom_iir[chn][4 * sg + 2] = y1; st.zoom_iir[chn][4 * sg + 3] = y2;
  }
  if (last) {
    st.zoom_fir_hist[chn][0] = h0; st.zoom_fir_hist[chn][1] = h1; st.zoom_fir_hist[chn][2] = h2;
    if (chn == 1) st.zoom_ptr = ptr;
  }
}
#endif

/* window the 512 samples into the spectrum FFT buffer */
/* part: 0 = every receiver; 1 = only those at zoom x1 (their input is the raw tile), 2 = only the zoomed ones (the
   rows-only kernel runs part 1 before the ZoomFFT cascade, whose idle lanes then have the x1 receivers' dead tiles to
   work on, and part 2 behind it) */
T41RX_DEV void PhSpecWindow(Cta &c, int tid, int part = 0) {
  if (!c.row) return;
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  float *s = Slot(c, g);
  const StreamCfg &cf = CfgOf(c, g);
  if ((part == 1 && cf.zoom != 0) || (part == 2 && cf.zoom == 0)) return;
  const StreamState &st = c.a.st[Sid(c, g)];
  float2 *buf = reinterpret_cast<float2 *>(s + vSpecFft);
  const IqFix fix = IqFixOf(cf);
  const int zoom = cf.zoom, zptr = st.zoom_ptr;
  const float zmult = cf.zoom_mult;
  for (int j = 0; j < 8; ++j) {
    const int i = u + 64 * j;
    const double w = LdgRO(c.a.hann + i);
    float re, im;
    if (zoom == 0) {             /* CalcZoom1Magn, FFT.cpp:220-223: raw (pre-shift) samples */
      float vi = s[oRawI + 27 + i], vq = s[oRawQ + 27 + i];
      IqCorr(fix, vi, vq);
      re = (float)((double)vi * w);
      im = (float)((double)vq * w);
    } else {                     /* ZoomFFTExe, FFT.cpp:109-116: ring, oldest first */
      const int p = (zptr + i) & (kSpecRes - 1);
      const float a = zmult * st.zoom_ring[0][p];
      const float b = zmult * st.zoom_ring[1][p];
      re = (float)((double)a * w);
      im = (float)((double)b * w);
    }
    buf[FftPhys(i)] = float2{re, im};
  }
}

T41RX_DEV void PhSpecFftPass(Cta &c, int tid, int pass) {
  if (!c.row) return;
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  Radix8Butterfly<true>(reinterpret_cast<float2 *>(Slot(c, g) + vSpecFft), c.a.twiddle, pass, u);
}

/* |X|^2 with half swap -> smoothing -> log -> pixel -> spectrum and waterfall rows */
T41RX_DEV void PhSpecRow(Cta &c, int tid) {
  if (!c.row) return;
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  float *s = Slot(c, g);
  const StreamCfg &cf = CfgOf(c, g);
  StreamState &st = c.a.st[Sid(c, g)];
  const float2 *buf = reinterpret_cast<const float2 *>(s + vSpecFft);
  const size_t row_base = ((size_t)(Sid(c, g)) * c.a.n_rows + c.row_idx) * kSpecRes;
  const float lpf = 0.7f;
  /* the smoothing state lives in HBM: all eight reads in flight before the first is used (the stores below would
     otherwise order them one trip to memory after the other) */
  float olds[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) olds[j] = st.spec_old[u + 64 * j];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int x = u + 64 * j;
    const int bin = (x + 256) & 511;
    const float2 v = buf[FftPhys((int)OctRev3((unsigned)bin))];
    const float pw = v.x * v.x + v.y * v.y;
    const float old = olds[j];
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
  const StreamCfg &cf = CfgOf(c, g);
  const StreamState &st = c.a.st[Sid(c, g)];
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
    const StreamCfg &cf = CfgOf(c, g);
    if (s[oMisc + mNcoMode] != 0.0f) {
      /* exact path: the FP64 oscillator recurrence, one lane per receiver */
      if (tid != g) continue;
      StreamState &st = c.a.st[Sid(c, g)];
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
      const IqFix fix = IqFixOf(cf);
      for (int n = 0; n < kBlock; ++n) {
        const double oq = (vq * oc) - (vi * os);
        const double oi = (vi * oc) + (vq * os);
        const double gain = 1.95 - ((vq * vq) + (vi * vi));
        vq = gain * oq;
        vi = gain * oi;
        float xi = s[oRawI + 27 + n], xq = s[oRawQ + 27 + n];
        IqCorr(fix, xi, xq);
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
      const IqFix fix = IqFixOf(cf);
      constexpr int kPer = kBlock / kNT;
      /* in-place: read this thread's samples first so the compiler may overlap their (long:
         two conversions each way) dependency chains */
      for (int k0 = 0; k0 < kPer; k0 += 8) {
        float xi[8], xq[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int n = tid + kNT * (k0 + k);
          xi[k] = s[oRawI + 27 + n];
          xq[k] = s[oRawQ + 27 + n];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int n = tid + kNT * (k0 + k);
          const D2 cb = tab[96 + (n >> 6)];
          const double oq = cb.x * w.x - cb.y * w.y;
          const double oi = cb.x * w.y + cb.y * w.x;
          IqCorr(fix, xi[k], xq[k]);
          QuarterShift(n, xi[k], xq[k]);
          MixStore(s, n, xi[k], xq[k], oq, oi);
        }
      }
    }
  }
}

/* closed form: advance the phase by one block */
T41RX_DEV void PhNcoAdvance(Cta &c, int tid) {
  if (tid >= c.ng) return;
  float *s = Slot(c, tid);
  if (s[oMisc + mNcoMode] != 0.0f) return;
  const StreamCfg &cf = CfgOf(c, tid);
  StreamState &st = c.a.st[Sid(c, tid)];
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
  /* 64 threads per receiver; a thread owns outputs u, u+64, ..., u+448 of both channels:
     16 independent accumulation chains keep the FMA pipe busy despite the in-order tap sums */
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  float *s = Slot(c, g);
  float taps[kDec1Taps];
#pragma unroll
  for (int i = 0; i < kDec1Taps; ++i) taps[i] = s[oTaps + kTapDec1 + i];
  for (int h = u; h < 90; h += 64) {     /* dec2 history back in front of the dec1 output */
    const int ch = h / 45, i = h % 45;
    s[(ch ? oD1Q : oD1I) + i] = s[oD2H + h];
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float acc[2][4];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[ch][r] = 0.0f;
#pragma unroll
    for (int q = 0; q < kDec1Taps / 4; ++q) {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const float *x = s + (ch ? oRawQ : oRawI);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int m = u + 64 * (4 * half + r);
          const float4 v = *reinterpret_cast<const float4 *>(x + 4 * m + 4 * q);
          acc[ch][r] = fmaf(v.x, taps[4 * q + 0], acc[ch][r]);
          acc[ch][r] = fmaf(v.y, taps[4 * q + 1], acc[ch][r]);
          acc[ch][r] = fmaf(v.z, taps[4 * q + 2], acc[ch][r]);
          acc[ch][r] = fmaf(v.w, taps[4 * q + 3], acc[ch][r]);
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
      for (int r = 0; r < 4; ++r) s[(ch ? oD1Q : oD1I) + 45 + u + 64 * (4 * half + r)] = acc[ch][r];
  }
}

/* dec2, level adjust (Process.cpp:482-492), overlap-save assembly (Process.cpp:498-522) */
T41RX_DEV void PhDec2(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  float *s = Slot(c, g);
  const StreamCfg &cf = CfgOf(c, g);
  const int mode = cf.mode;
  const float vol_scale = cf.vol_scale;
  const bool first_block = c.a.st[Sid(c, g)].first_block != 0;
  /* dec1's input region is about to be overlaid: keep its last 27 samples */
  for (int h = u; h < 54; h += 64) {
    const int ch = h / 27, i = h % 27;
    s[oD1H + h] = s[(ch ? oRawQ : oRawI) + kBlock + i];
  }
  float taps[kDec2Taps];
#pragma unroll
  for (int i = 0; i < kDec2Taps; ++i) taps[i] = s[oTaps + kTapDec2 + i];
  /* outputs u, u+64, u+128, u+192 of both channels: 8 independent chains */
  float acc[2][4];
#pragma unroll
  for (int ch = 0; ch < 2; ++ch)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[ch][r] = 0.0f;
#pragma unroll
  for (int q = 0; q < kDec2Taps / 2; ++q) {
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float *x = s + (ch ? oD1Q : oD1I);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int o = u + 64 * r;
        const float2 v = *reinterpret_cast<const float2 *>(x + 2 * o + 2 * q);
        acc[ch][r] = fmaf(v.x, taps[2 * q], acc[ch][r]);
        acc[ch][r] = fmaf(v.y, taps[2 * q + 1], acc[ch][r]);
      }
    }
  }
  float2 *fa = reinterpret_cast<float2 *>(s + vFftA);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int o = u + 64 * r;
    if (mode == kModePsk31) {
      s[vAud + 23 + o] = acc[0][r];                /* Process.cpp:376-387,745: raw decimated I */
    } else if (mode == kModeNfm) {
      fa[FftPhys(kDec + o)] = float2{acc[0][r], acc[1][r]}; /* Process.cpp:272-275 */
    } else {
      const float li = acc[0][r] * vol_scale, lq = acc[1][r] * vol_scale;
      float2 prev = float2{s[oOla + o], s[oOla + 256 + o]};
      if (first_block) prev = float2{0.0f, 0.0f};
      fa[FftPhys(o)] = prev;
      fa[FftPhys(kDec + o)] = float2{li, lq};
      s[oOla + o] = li;
      s[oOla + 256 + o] = lq;
    }
  }
}

/* dec2 history, NFM discriminator (Demod.cpp:220-235, Process.cpp:716-727,768-779) */
T41RX_DEV void PhPostDec2(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamCfg &cf = CfgOf(c, g);
    StreamState &st = c.a.st[Sid(c, g)];
    for (int h = tid; h < 90; h += kNT) {
      const int ch = h / 45, i = h % 45;
      s[oD2H + h] = s[(ch ? oD1Q : oD1I) + kDec1Out + i];
    }
    if (tid == 0 && cf.mode != kModePsk31 && cf.mode != kModeNfm) st.first_block = 0;
    if (cf.mode == kModeNfm) {
      const float2 *fa = reinterpret_cast<const float2 *>(s + vFftA);
      /* fmdemod_quadri_K is a double literal (Demod.h:7): K * num / den runs in double */
      const double kq = 0.340447550238101026565118445432744920253753662109375;
      for (int o = tid; o < kDec; o += kNT) {
      const float2 now = fa[FftPhys(kDec + o)];
      const float den = now.x * now.x + now.y * now.y;
      float out;
      if (o == 0) {
        const float li = st.nfm_last_i, lq = st.nfm_last_q;
        const float num = now.x * (now.y - lq) - now.y * (now.x - li);
        out = (float)(kq * (double)num / (double)den);
      } else {
        const float2 last = fa[FftPhys(kDec + o - 1)];
        const float num = now.y * last.x - now.x * last.y;
        out = (float)(kq * (double)num / (double)den);
        out = (1.0f < out) ? 1.0f : out;           /* limiter skips index 0 (B5) */
        out = (-1.0f > out) ? -1.0f : out;
      }
      s[vAmTmp + o] = out;
      }
    }
  }
}

/* NFM: build the real-input overlap-save buffer from the demodulated audio */
T41RX_DEV void PhNfmAssemble(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamCfg &cf = CfgOf(c, g);
    if (cf.mode != kModeNfm) continue;
    StreamState &st = c.a.st[Sid(c, g)];
    float2 *fa = reinterpret_cast<float2 *>(s + vFftA);
    if (tid == 0) {                                /* "last sample" = complex sample 127 (B4) */
      st.nfm_last_i = fa[FftPhys(kDec + 127)].x;
      st.nfm_last_q = fa[FftPhys(kDec + 127)].y;
    }
  }
}
T41RX_DEV void PhNfmAssemble2(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamCfg &cf = CfgOf(c, g);
    if (cf.mode != kModeNfm) continue;
    float2 *fa = reinterpret_cast<float2 *>(s + vFftA);
    for (int o = tid; o < kDec; o += kNT) {
      const float a = s[vAmTmp + o];
      fa[FftPhys(o)] = float2{s[oOla + o], 0.0f};
      fa[FftPhys(kDec + o)] = float2{a, 0.0f};
      s[oOla + o] = a;
    }
  }
}

/* ------------------------------------------------------------------ */
/* P6-P8: fast convolution (Process.cpp:535-595,787-808)                */
/* ------------------------------------------------------------------ */
T41RX_DEV bool UsesFilter(int mode) { return mode != kModePsk31; }

T41RX_DEV void PhFftPass(Cta &c, int tid, int which, int pass) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  if (!UsesFilter(CfgOf(c, g).mode)) return;
  Radix8Butterfly<true>(reinterpret_cast<float2 *>(Slot(c, g) + (which ? vFftB : vFftA)), c.a.twiddle, pass, u);
}

/* digit-reverse the forward result, multiply by the mask, conjugate for the inverse */
T41RX_DEV void PhMask(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (!UsesFilter(cf.mode)) return;
  float *s = Slot(c, g);
  const float2 *fa = reinterpret_cast<const float2 *>(s + vFftA);
  float2 *fb = reinterpret_cast<float2 *>(s + vFftB);
  const float2 *mask = reinterpret_cast<const float2 *>(c.a.fsets[cf.filter_id].mask);
  /* (the display rows of this block may come from the rows-only kernel: c.row is not the test here) */
  float2 *arow = nullptr;
  if (c.a.aspec && c.a.row_every > 0 && ((c.a.t0 + c.t) % c.a.row_every) == 0)
    arow = c.a.aspec + ((size_t)Sid(c, g) * c.a.n_rows + (c.a.t0 + c.t) / c.a.row_every) * kFft;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = u + 64 * j;
    const float2 x = fa[FftPhys((int)OctRev3((unsigned)k))];
    const float2 h = LdgRO(mask + k);
    const float rr = x.x * h.x, ii = x.y * h.y, ri = x.x * h.y, ir = x.y * h.x;
    fb[FftPhys(k)] = float2{rr - ii, -(ri + ir)};
    /* row-producing block: keep the masked spectrum for the audio-spectrum by-product (Process.cpp:550-553) */
    if (arow) arow[k] = float2{rr - ii, ri + ir};
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
  const StreamCfg &cf = CfgOf(c, g);
  if (!UsesFilter(cf.mode)) return;
  float *s = Slot(c, g);
  const float2 *fb = reinterpret_cast<const float2 *>(s + vFftB);
  float2 *zext = reinterpret_cast<float2 *>(s + vZext);
  const float inv = 1.0f / 512.0f;
  float2 z[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = u + 64 * j;
    const float2 v = fb[FftPhys((int)OctRev3((unsigned)(kDec + i)))];
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

/* Sliding maximum over 97 entries: rm[i] = max |z| over entries [e - 96, e], e = i + 97, of the 353-entry magnitude
 * array (97 delayed + 256 new).  Three steps instead of a doubling ladder of seven:
 *   A  chunks of 8: prefix maxima P[e] and suffix maxima S[e] inside every chunk, and the chunk maximum M[c];
 *   B  W[c] = max M[c - 11 .. c - 1]: the 11 whole chunks inside every window that ends in chunk c;
 *   C  with e = 8 c + o: rm = max(S[8 (c - 12) + o], W[c], P[e])  (97 = 12 x 8 + 1: the window starts at offset o of
 *      chunk c - 12 and ends at offset o of chunk c).
 * max is exact, so any order gives the reference's value (`if (abs > ring_max)`, DSP_Fn.cpp:509-519). */
constexpr int kMaxChunk = 8;
constexpr int kMaxChunks = (kAgcDelay + kDec + kMaxChunk - 1) / kMaxChunk;        /* 45 */
static_assert(kAgcDelay == 12 * kMaxChunk + 1 && kMaxChunks <= 64, "window = 12 chunks + 1 entry");
static_assert(vAbs % 4 == 0 && vLvlA % 4 == 0 && vLvlB % 4 == 0 && kSlot % 4 == 0, "16-byte accesses of the chunk arrays");
constexpr int vMaxW = oD1I, vMaxM = oD1I + 64;     /* the dec1 output region is dead between PhDec2 and the equaliser */
T41RX_DEV void PhAgcMaxA(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng || u >= kMaxChunks) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (!UsesFilter(cf.mode) || cf.agc_mode == 0) return;
  float *s = Slot(c, g);
  const int n = kAgcDelay + kDec;   /* 353 */
  float a[kMaxChunk], pm[kMaxChunk], sm[kMaxChunk];
  {
    /* (the last chunk reads past entry 352: still inside the slot, masked below) */
    const float4 lo = *reinterpret_cast<const float4 *>(s + vAbs + kMaxChunk * u);
    const float4 hi = *reinterpret_cast<const float4 *>(s + vAbs + kMaxChunk * u + 4);
    a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w;
    a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
  }
#pragma unroll
  for (int j = 0; j < kMaxChunk; ++j) {
    /* the reference's `if (abs > ring_max)` never lets a NaN magnitude (NFM discriminator on exact silence: 0/0)
       become the maximum: NaNs count as 0 here */
    a[j] = (kMaxChunk * u + j < n && a[j] == a[j]) ? a[j] : 0.0f;
  }
  pm[0] = a[0];
#pragma unroll
  for (int j = 1; j < kMaxChunk; ++j) pm[j] = fmaxf(pm[j - 1], a[j]);
  sm[kMaxChunk - 1] = a[kMaxChunk - 1];
#pragma unroll
  for (int j = kMaxChunk - 2; j >= 0; --j) sm[j] = fmaxf(sm[j + 1], a[j]);
  if (kMaxChunk * u + kMaxChunk <= n) {
    *reinterpret_cast<float4 *>(s + vLvlA + kMaxChunk * u) = float4{pm[0], pm[1], pm[2], pm[3]};
    *reinterpret_cast<float4 *>(s + vLvlA + kMaxChunk * u + 4) = float4{pm[4], pm[5], pm[6], pm[7]};
    *reinterpret_cast<float4 *>(s + vLvlB + kMaxChunk * u) = float4{sm[0], sm[1], sm[2], sm[3]};
    *reinterpret_cast<float4 *>(s + vLvlB + kMaxChunk * u + 4) = float4{sm[4], sm[5], sm[6], sm[7]};
  } else {                           /* the last, partial chunk: its prefix maxima only (no window starts in it) */
#pragma unroll
    for (int j = 0; j < kMaxChunk; ++j)
      if (kMaxChunk * u + j < n) s[vLvlA + kMaxChunk * u + j] = pm[j];
  }
  s[vMaxM + u] = pm[kMaxChunk - 1];
}
T41RX_DEV void PhAgcMaxB(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng || u < 12 || u >= kMaxChunks) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (!UsesFilter(cf.mode) || cf.agc_mode == 0) return;
  float *s = Slot(c, g);
  float w = s[vMaxM + u - 11];
#pragma unroll
  for (int j = 10; j >= 1; --j) w = fmaxf(w, s[vMaxM + u - j]);
  s[vMaxW + u] = w;
}
T41RX_DEV void PhAgcMaxC(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (!UsesFilter(cf.mode) || cf.agc_mode == 0) return;
  float *s = Slot(c, g);
  for (int i = u; i < kDec; i += 64) {
    const int e = i + kAgcDelay;
    s[vRm + i] = fmaxf(fmaxf(s[vLvlB + e - 12 * kMaxChunk], s[vMaxW + e / kMaxChunk]), s[vLvlA + e]);
  }
}

/* the serial envelope state machine (DSP_Fn.cpp:521-626): lane 0 of the receiver's first warp,
 * so that receivers in different AGC states do not serialise each other through divergence.
 * All constants and state live in registers; the per-sample inputs (delayed |z| and the
 * window maximum) are fetched four samples ahead of the recurrence. */
T41RX_DEV void PhAgcSerial(Cta &c, int tid) {
  const int g = SerialStream(c, tid);
  if (g < 0) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (!UsesFilter(cf.mode) || cf.agc_mode == 0) return;
  float *s = Slot(c, g);
  StreamState &st = c.a.st[Sid(c, g)];
  const float k_fbm = cf.agc.fast_backmult, k_omfbm = cf.agc.onemfast_backmult;
  const float k_hbm = cf.agc.hang_backmult, k_omhbm = cf.agc.onemhang_backmult;
  const float k_attack = cf.agc.attack_mult, k_decay = cf.agc.decay_mult, k_fdecay = cf.agc.fast_decay_mult;
  const float k_hdecay = cf.agc.hang_decay_mult, k_pop = cf.agc.pop_ratio, k_hlevel = cf.agc.hang_level;
  const float k_minv = cf.agc.min_volts;
  const int k_hload = cf.agc.hang_counter_load, k_henable = cf.agc.hang_enable;
  float fast = st.agc_fast_back, hang = st.agc_hang_back, v = st.agc_volts, save = st.agc_save_volts;
  int hc = st.agc_hang_counter, state = st.agc_state, dtype = st.agc_decay_type, action = st.agc_action;
  float rm = st.agc_ring_max;
  int i = 0;
  while (i < kDec) {
    if (state == 3) {
      /* run-length fast path for the steady state (slow decay, DSP_Fn.cpp:601-608): one loop-closing
         branch per sample while the window maximum stays below volts.  Identical arithmetic to the
         generic step below. */
      /* the next sample's window maximum and magnitude are fetched before this sample's branch: the branch waits
         for volts anyway, the loads need not wait for the branch (index kDec of either array is readable scratch) */
      float r_nx = s[vRm + i], a_nx = s[vAbs + i];
      while (i < kDec) {
        const float r = r_nx, abs_out = a_nx;
        r_nx = s[vRm + i + 1];
        a_nx = s[vAbs + i + 1];
        if (r >= v) break;
        fast = k_fbm * abs_out + k_omfbm * fast;
        hang = k_hbm * abs_out + k_omhbm * hang;
        rm = r;
        hc = (hc > 0) ? hc - 1 : hc;
        const float step = (r - v) * k_decay;
        v = (float)((double)v + (double)step * .05);   /* double product and sum */
        action = (v < k_minv) ? 0 : 1;
        v = (v < k_minv) ? k_minv : v;
        s[vVolt + i] = v;
        ++i;
      }
      if (i >= kDec) break;
    }
    /* generic step: the 5-state machine of DSP_Fn.cpp:521-626 */
    {
      const float abs_out = s[vAbs + i];
      fast = k_fbm * abs_out + k_omfbm * fast;
      hang = k_hbm * abs_out + k_omhbm * hang;
      rm = s[vRm + i];
      if (hc > 0) --hc;
      const float d = rm - v;
      if (rm >= v) {                       /* every state: attack; 2,3,4 remember the level they left */
        if (state >= 2) save = v;
        state = 0;
        v += d * k_attack;
      } else if (state == 3) {
        const float step = d * k_decay;
        v = (float)((double)v + (double)step * .05);
      } else if (state == 0) {
        if (v > k_pop * fast) {
          state = 1;
          v += d * k_fdecay;
        } else if (k_henable && (hang > k_hlevel)) {
          state = 2;
          hc = k_hload;
          dtype = 1;
        } else {
          state = 3;
          v += d * k_decay;
          dtype = 0;
        }
      } else if (state == 1) {
        if (v > save) {
          v += d * k_fdecay;
        } else if (hc > 0) {
          state = 2;
        } else if (dtype == 0) {
          state = 3;
          v += d * k_decay;
        } else {
          state = 4;
          v += d * k_hdecay;
        }
      } else if (state == 2) {
        if (hc == 0) {
          state = 4;
          v += d * k_hdecay;
        }
      } else {
        v += d * k_hdecay;
      }
      action = (v < k_minv) ? 0 : 1;
      v = (v < k_minv) ? k_minv : v;
      s[vVolt + i] = v;
      ++i;
    }
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
  const StreamCfg &cf = CfgOf(c, g);
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
  const StreamCfg &cf = CfgOf(c, g);
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
  if (cf.mode == kModeSam) {
    for (int i = u; i < 513; i += 64) s[vSamSin + i] = LdgRO(c.a.sin_table + i);
  }
}

T41RX_DEV void PhDemodSerial(Cta &c, int tid) {
  const int g = SerialStream(c, tid);
  if (g < 0) return;
  const StreamCfg &cf = CfgOf(c, g);
  float *s = Slot(c, g);
  StreamState &st = c.a.st[Sid(c, g)];
  const float2 *dem = reinterpret_cast<const float2 *>(s + vDem);
  if (cf.mode == kModeAm) {
    /* DC removal (Process.cpp:698-704) then 1-stage DF1 low-pass (Process.cpp:705) */
    float wold = st.am_wold;
    float x1 = st.am_lp_state[0], x2 = st.am_lp_state[1], y1 = st.am_lp_state[2], y2 = st.am_lp_state[3];
    const float b0 = cf.am_lp[0], b1 = cf.am_lp[1], b2 = cf.am_lp[2], a1 = cf.am_lp[3], a2 = cf.am_lp[4];
    for (int i0 = 0; i0 < kDec; i0 += 4) {
      float m4[4], o4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) m4[j] = s[vAmTmp + i0 + j];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w = m4[j] + wold * 0.99f;
        const float x = w - wold;
        wold = w;
        float acc = b0 * x;
        acc = acc + b1 * x1;
        acc = acc + b2 * x2;
        acc = acc + a1 * y1;
        acc = acc + a2 * y2;
        x2 = x1; x1 = x;
        y2 = y1; y1 = acc;
        o4[j] = acc;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) s[vAud + 23 + i0 + j] = o4[j];
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
    /* The phase of sample i + 1 is phz + the loop filter's output of sample i - 1 (del_out): it does not wait for
       sample i's detector.  Its sine / cosine are therefore formed beside sample i's arctangent - the same
       operations on the same values as the reference's loop, two dependent chains side by side instead of one. */
    float sn = TableTurns(s + vSamSin, phz * 0.159154943092f);
    float cs = TableTurns(s + vSamSin, phz * 0.159154943092f + 0.25f);
    for (int i = 0; i < kDec; ++i) {
      const float2 z = dem[i];
      float phz_n = phz + fil;                       /* del_out = fil_out before this sample's update */
      /* the reference's two while loops: phz is in [0, 2 pi) and |fil_out| <= g1 * 2 pi + omega_max < 1.2, so each
         runs at most once */
      phz_n = (phz_n >= tpi) ? phz_n - tpi : phz_n;
      phz_n = (phz_n < 0.0f) ? phz_n + tpi : phz_n;
      const float sn_n = TableTurns(s + vSamSin, phz_n * 0.159154943092f);
      const float cs_n = TableTurns(s + vSamSin, phz_n * 0.159154943092f + 0.25f);
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
      om2 = om2 + g2 * det;
      if (om2 < omega_min) om2 = omega_min;
      else if (om2 > omega_max) om2 = omega_max;
      fil = g1 * det + om2;
      phz = phz_n;
      sn = sn_n;
      cs = cs_n;
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
    const size_t o = (size_t)(Sid(c, g)) * c.a.t_stride + c.t;
    if (c.a.psk_bits) c.a.psk_bits[o] = bit_out;
    if (c.a.psk_chars) c.a.psk_chars[o] = char_out;
  } else {
    const size_t o = (size_t)(Sid(c, g)) * c.a.t_stride + c.t;
    if (c.a.psk_bits) c.a.psk_bits[o] = -1;
    if (c.a.psk_chars) c.a.psk_chars[o] = 0;
  }
}

/* ------------------------------------------------------------------ */
/* Split form of the chain: what PhAgcSerial / PhAgcPost / PhDemod* do   */
/* on ONE LANE PER RECEIVER (the envelope detector, the SAM PLL and the  */
/* AM detector are serial chains over the 256 samples of a block) runs   */
/* in a kernel of its own with THREAD = RECEIVER, 32 receivers per warp, */
/* over all blocks of the launch; the sample-parallel phases before and  */
/* after it keep their 64 threads per receiver in a front and a back     */
/* kernel.  Every float operation is the one the fused schedule does, in */
/* the same order: the three kernels are bit-identical to it.            */
/* ------------------------------------------------------------------ */
/* front kernel, after the AGC look-ahead: hand the block to the serial kernel and refresh the delay line
   (PhAgcPost's second half).  Thread (i, g) = (tid >> 2, tid & 3): four adjacent lanes write 64 contiguous bytes. */
T41RX_DEV void PhSerialStore(Cta &c, int tid) {
  const int g = tid % kG, u = tid / kG;           /* u: 0..63 */
  if (g < c.ng) {
    const StreamCfg &cf = CfgOf(c, g);
    const float *s = Slot(c, g);
    const size_t n = (size_t)c.a.n_streams;
    float4 *dst = c.a.ser_in + ((size_t)c.t * kDec) * n + (size_t)(c.s0 + g);
    const bool filt = UsesFilter(cf.mode);
    const bool agc = filt && cf.agc_mode != 0;
    const float2 *zext = reinterpret_cast<const float2 *>(s + vZext);
    const float2 *dem = reinterpret_cast<const float2 *>(s + vDem);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = u + 64 * j;
      float4 v;
      if (agc) {
        const float2 z = zext[i];
        v = float4{z.x, z.y, s[vAbs + i], s[vRm + i]};
      } else if (filt) {
        const float2 z = dem[i];                  /* AGC off: PhAgcPre applied the fixed gain */
        v = float4{z.x, z.y, 0.0f, 0.0f};
      } else {
        v = float4{s[vAud + 23 + i], 0.0f, 0.0f, 0.0f};   /* raw PSK31 mode: the decimated I channel is the audio */
      }
      dst[(size_t)i * n] = v;
    }
  }
}
T41RX_DEV void PhSerialRing(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (!UsesFilter(cf.mode) || cf.agc_mode == 0) return;
  float *s = Slot(c, g);
  const float2 *zext = reinterpret_cast<const float2 *>(s + vZext);
  /* ring index r holds new sample 128 + r of this block = zext[97 + 128 + r] */
  for (int r = u; r < kAgcRing; r += 64) {
    const float2 z = zext[kAgcDelay + 128 + r];
    s[oAgc + r] = z.x;
    s[oAgc + 128 + r] = z.y;
    s[oAgc + 256 + r] = s[vAbs + kAgcDelay + 128 + r];
  }
}

/* Codec_gain alone (the front kernel's share of PhBlockEnd) */
T41RX_DEV void PhCodecGain(Cta &c, int tid) {
  if (tid != kNT - 1) return;
  for (int g = 0; g < c.ng; ++g) {
    StreamState &st = c.a.st[Sid(c, g)];
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

/* ---- the serial stages as three small state machines (shared by the host emulation, which runs them one after the
   other per sample, and the device kernel, which runs them on three warps as a pipeline over chunks of samples) ---- */
/* AGC envelope (DSP_Fn.cpp:504-626; PhAgcSerial's generic step) */
struct SerAgc {
  float k_fbm, k_omfbm, k_hbm, k_omhbm, k_attack, k_decay, k_fdecay, k_hdecay, k_pop, k_hlevel, k_minv;
  int k_hload, k_henable;
  float fast, hang, v, save, rm;
  int hc, state, dtype, action;
  T41RX_DEV void Load(const StreamCfg &cf, const StreamState &st) {
    k_fbm = cf.agc.fast_backmult; k_omfbm = cf.agc.onemfast_backmult;
    k_hbm = cf.agc.hang_backmult; k_omhbm = cf.agc.onemhang_backmult;
    k_attack = cf.agc.attack_mult; k_decay = cf.agc.decay_mult; k_fdecay = cf.agc.fast_decay_mult;
    k_hdecay = cf.agc.hang_decay_mult; k_pop = cf.agc.pop_ratio; k_hlevel = cf.agc.hang_level;
    k_minv = cf.agc.min_volts;
    k_hload = cf.agc.hang_counter_load; k_henable = cf.agc.hang_enable;
    fast = st.agc_fast_back; hang = st.agc_hang_back; v = st.agc_volts; save = st.agc_save_volts;
    rm = st.agc_ring_max;
    hc = st.agc_hang_counter; state = st.agc_state; dtype = st.agc_decay_type; action = st.agc_action;
  }
  T41RX_DEV void Store(StreamState &st) const {
    st.agc_fast_back = fast; st.agc_hang_back = hang; st.agc_volts = v; st.agc_save_volts = save;
    st.agc_ring_max = rm;
    st.agc_hang_counter = hc; st.agc_state = state; st.agc_decay_type = dtype; st.agc_action = action;
  }
  /* one sample: delayed |z| and the window maximum in, volts out.
     The reference's ladder of ifs (DSP_Fn.cpp:521-626) written as selects: the lanes of the serial kernel's AGC warp are 32
     receivers in whatever states their signals put them, and as branches the ladder ran once per state present in the
     warp (tools/microbench/ser_stages.cu: 419 clocks per sample with the lanes in different states, the slowest stage of
     the pipeline).  Every candidate value is formed by the same operations on the same operands as in its branch. */
  T41RX_DEV float Step(float abs_out, float r) {
    fast = k_fbm * abs_out + k_omfbm * fast;
    hang = k_hbm * abs_out + k_omhbm * hang;
    rm = r;
    hc = (hc > 0) ? hc - 1 : hc;
    const float d = rm - v;
    const float v_att = v + d * k_attack;
    const float v_fd = v + d * k_fdecay;
    const float v_dk = v + d * k_decay;
    const float v_hd = v + d * k_hdecay;
    const float step3 = d * k_decay;
    const float v_d3 = (float)((double)v + (double)step3 * .05);   /* state 3: double product and sum */
    const bool att = rm >= v;                        /* every state: attack; 2,3,4 remember the level they left */
    const bool pop = v > k_pop * fast;
    const bool hangc = (k_henable != 0) && (hang > k_hlevel);
    const bool above = v > save;
    const bool hc_pos = hc > 0;
    const int s = state;
    /* state 0 without attack: 1 (fast decay) | 2 (hang) | 3 (decay) */
    const int n0 = pop ? 1 : (hangc ? 2 : 3);
    const float v0 = pop ? v_fd : (hangc ? v : v_dk);
    /* state 1 without attack: stay | 2 | 3 | 4 */
    const int n1 = above ? 1 : (hc_pos ? 2 : (dtype == 0 ? 3 : 4));
    const float v1 = above ? v_fd : (hc_pos ? v : (dtype == 0 ? v_dk : v_hd));
    /* state 2: until the hang counter runs out; state 4: hang decay */
    const int n2 = hc_pos ? 2 : 4;
    const float v2 = hc_pos ? v : v_hd;
    int ns = (s == 3) ? 3 : (s == 0) ? n0 : (s == 1) ? n1 : (s == 2) ? n2 : s;
    float nv = (s == 3) ? v_d3 : (s == 0) ? v0 : (s == 1) ? v1 : (s == 2) ? v2 : v_hd;
    const bool to_hang = !att && s == 0 && !pop && hangc;
    const bool to_decay = !att && s == 0 && !pop && !hangc;
    hc = to_hang ? k_hload : hc;
    dtype = to_hang ? 1 : (to_decay ? 0 : dtype);
    save = (att && s >= 2) ? v : save;
    ns = att ? 0 : ns;
    nv = att ? v_att : nv;
    state = ns;
    action = (nv < k_minv) ? 0 : 1;
    v = (nv < k_minv) ? k_minv : nv;
    return v;
  }
};

/* gain from volts (DSP_Fn.cpp:628; PhAgcPost) */
struct SerGain {
  float k_inv_in, k_target, k_slope;
  T41RX_DEV void Load(const StreamCfg &cf) {
    k_inv_in = cf.agc.inv_max_input; k_target = cf.agc.out_target; k_slope = cf.agc.slope_constant;
  }
  T41RX_DEV float Mult(float v) const {
    const double lg = (double)Log10Fast(k_inv_in * v);
    const double clipped = (0.0 < lg) ? 0.0 : lg;
    return (float)(((double)k_target - (double)k_slope * clipped) / (double)v);
  }
};

/* demodulators with a serial chain (Process.cpp:697-707 AM, Demod.cpp:40-139 SAM) and the pass-through modes */
struct SerDemod {
  int mode;
  /* AM */
  float wold, x1, x2, y1, y2, b0, b1, b2, a1, a2;
  /* SAM */
  float omega_min, omega_max, g1, g2, phz, fil, om2, sn, cs;
  const float *tab;
  T41RX_DEV void Load(const LaunchArgs &a, const StreamCfg &cf, const StreamState &st, const float *sin_tab) {
    mode = cf.mode;
    wold = st.am_wold;
    x1 = st.am_lp_state[0]; x2 = st.am_lp_state[1]; y1 = st.am_lp_state[2]; y2 = st.am_lp_state[3];
    b0 = cf.am_lp[0]; b1 = cf.am_lp[1]; b2 = cf.am_lp[2]; a1 = cf.am_lp[3]; a2 = cf.am_lp[4];
    omega_min = LdgRO(a.sam_consts + 0); omega_max = LdgRO(a.sam_consts + 1);
    g1 = LdgRO(a.sam_consts + 2); g2 = LdgRO(a.sam_consts + 3);
    phz = st.sam_phzerror; fil = st.sam_fil_out; om2 = st.sam_omega2;
    sn = 0.0f; cs = 0.0f;
    tab = sin_tab;
  }
  T41RX_DEV void Store(StreamState &st) const {
    if (mode == kModeAm) {
      st.am_wold = wold;
      st.am_lp_state[0] = x1; st.am_lp_state[1] = x2; st.am_lp_state[2] = y1; st.am_lp_state[3] = y2;
    } else if (mode == kModeSam) {
      st.sam_phzerror = phz;
      st.sam_fil_out = fil;
      st.sam_omega2 = om2;
    }
  }
  /* at the start of every block: the reference's loop takes the sine / cosine of the carried phase first */
  T41RX_DEV void BlockStart() {
    if (mode == kModeSam) {
      sn = TableTurns(tab, phz * 0.159154943092f);
      cs = TableTurns(tab, phz * 0.159154943092f + 0.25f);
    }
  }
  /* one gained sample in, one audio sample out */
  T41RX_DEV float Step(float zx, float zy) {
    if (mode == kModeSam) {
      const float tpi = 6.283185307179586476925286766559f;
      /* the phase of sample i + 1 is phz + the loop filter's output of sample i - 1: its sine / cosine are formed
         beside sample i's arctangent (same operations on the same values as the reference's loop) */
      float phz_n = phz + fil;
      phz_n = (phz_n >= tpi) ? phz_n - tpi : phz_n;
      phz_n = (phz_n < 0.0f) ? phz_n + tpi : phz_n;
      const float sn_n = TableTurns(tab, phz_n * 0.159154943092f);
      const float cs_n = TableTurns(tab, phz_n * 0.159154943092f + 0.25f);
      const float ai = cs * zx, bi = sn * zx, aq = cs * zy, bq = sn * zy;
      const float corr0 = +ai + bq;
      const float corr1 = -bi + aq;
      float audio = (ai - bi) + (aq + bq);
      audio = (audio + 0.0f) - 0.0f;             /* the fade leveller's identity (see PhDemodSerial) */
      const float det = Atan2Approx(corr1, corr0);
      om2 = om2 + g2 * det;
      if (om2 < omega_min) om2 = omega_min;
      else if (om2 > omega_max) om2 = omega_max;
      fil = g1 * det + om2;
      phz = phz_n;
      sn = sn_n;
      cs = cs_n;
      return audio;
    }
    if (mode == kModeAm) {
      const float m = AlphaBetaMag(zx, zy);
      const float w = m + wold * 0.99f;
      const float x = w - wold;
      wold = w;
      float acc = b0 * x;
      acc = acc + b1 * x1;
      acc = acc + b2 * x2;
      acc = acc + a1 * y1;
      acc = acc + a2 * y2;
      x2 = x1; x1 = x;
      y2 = y1; y1 = acc;
      return acc;
    }
    return zx;                                   /* USB / LSB / NFM: real part; raw PSK31 mode: the sample itself */
  }
};

/* PSK31 tap (psk31.cpp:235-310) on the first gained sample of a block; writes the block's bit / character outputs */
T41RX_DEV void SerPskTap(const LaunchArgs &a, const StreamCfg &cf, StreamState &st, int sid, int t, float2 dem0) {
  int8_t bit_out = -1;
  uint8_t char_out = 0;
  if (cf.psk31_enable && cf.mode != kModeNfm && cf.mode != kModePsk31) {
    if (st.psk_block_count % 3u == 0u) {
      const double pi_d = 3.1415926535897932384626433832795;
      const float phase = Atan2Approx(dem0.y, dem0.x);
      float dphase = phase - st.psk_last_phase;
      while ((double)dphase < -pi_d) dphase = (float)((double)dphase + 2 * pi_d);
      while ((double)dphase >= pi_d) dphase = (float)((double)dphase - 2 * pi_d);
      const uint8_t bit = (((double)dphase > (pi_d / 2)) || ((double)dphase < (-pi_d / 2))) ? 0 : 1;
      st.psk_last_phase = phase;
      bit_out = (int8_t)bit;
      unsigned long long shr = (st.psk_shr << 1) | (unsigned long long)bit;
      if ((shr & 0xFFFull) != 0) {
        for (int i = 0; i < 128; ++i) {
          const uint32_t e = LdgRO(a.varicode + i);
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
  }
  const size_t o = (size_t)sid * a.t_stride + t;
  if (a.psk_bits) a.psk_bits[o] = bit_out;
  if (a.psk_chars) a.psk_chars[o] = char_out;
}

/* the serial kernel's work for receiver number r of the launch, one stage after the other per sample (the host
   emulation's form; the device kernel runs the same three state machines as a warp pipeline, rx_api.cu) */
T41RX_DEV void SerialReceiver(const LaunchArgs &a, int r, const float *sin_tab) {
  const int sid = a.stream_ids ? LdgRO(a.stream_ids + r) : a.stream_base + r;
  const StreamCfg &cf = a.cfg[sid];
  StreamState &st = a.st[sid];
  const size_t n = (size_t)a.n_streams;
  const bool agc_on = UsesFilter(cf.mode) && cf.agc_mode != 0;
  SerAgc agc;
  SerGain gain;
  SerDemod dem;
  agc.Load(cf, st);
  gain.Load(cf);
  dem.Load(a, cf, st, sin_tab);
  for (int t = 0; t < a.n_blocks; ++t) {
    float2 dem0 = float2{0.0f, 0.0f};
    dem.BlockStart();
    for (int i = 0; i < kDec; ++i) {
      const size_t e = ((size_t)t * kDec + i) * n + r;
      float4 x = LdgRO(a.ser_in + e);
      if (agc_on) {
        const float m = gain.Mult(agc.Step(x.z, x.w));
        x.x = x.x * m;
        x.y = x.y * m;
      }
      if (i == 0) dem0 = float2{x.x, x.y};
      a.ser_out[e] = dem.Step(x.x, x.y);
    }
    SerPskTap(a, cf, st, sid, t, dem0);
  }
  if (agc_on) agc.Store(st);
  dem.Store(st);
}

/* back kernel: the demodulated block from the serial kernel into the audio buffer (+ PhInterp1's history restore) */
T41RX_DEV void PhBackLoad(Cta &c, int tid) {
  const int g = tid % kG, u = tid / kG;
  if (g < c.ng) {
    float *s = Slot(c, g);
    const size_t n = (size_t)c.a.n_streams;
    const float *src = c.a.ser_out + ((size_t)c.t * kDec) * n + (size_t)(c.s0 + g);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = u + 64 * j;
      s[vAud + 23 + i] = src[(size_t)i * n];
    }
  }
  for (int gg = 0; gg < c.ng; ++gg) {
    float *s = Slot(c, gg);
    for (int h = tid; h < 23; h += kNT) s[vAud + h] = s[oIntH + h];
  }
}
T41RX_DEV void PhBackStateIn(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    const StreamState &st = c.a.st[Sid(c, g)];
    const FilterSet &fs = c.a.fsets[CfgOf(c, g).filter_id];
    for (int i = tid; i < 154; i += kNT) {
      float v;
      if (i < kTapDec2) v = fs.dec1[i];
      else if (i < kTapInt1) v = fs.dec2[i - kTapDec2];
      else if (i < kTapInt2) v = fs.int1[i - kTapInt1];
      else v = fs.int2[i - kTapInt2];
      s[oTaps + i] = v;
    }
    for (int i = tid; i < 23; i += kNT) s[oIntH + i] = st.int1_hist[i];
    for (int i = tid; i < 7; i += kNT) s[oIntH + 24 + i] = st.int2_hist[i];
  }
}
T41RX_DEV void PhBackStateOut(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    StreamState &st = c.a.st[Sid(c, g)];
    for (int i = tid; i < 23; i += kNT) st.int1_hist[i] = s[oIntH + i];
    for (int i = tid; i < 7; i += kNT) st.int2_hist[i] = s[oIntH + 24 + i];
  }
}
/* the front kernel's state: everything PhStateOut stores except the interpolator histories (the back kernel's) */
T41RX_DEV void PhFrontStateOut(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    StreamState &st = c.a.st[Sid(c, g)];
    for (int i = tid; i < 512; i += kNT) st.ola_prev[i >> 8][i & 255] = s[oOla + i];
    for (int i = tid; i < 128; i += kNT) {
      st.agc_re[i] = s[oAgc + i];
      st.agc_im[i] = s[oAgc + 128 + i];
      st.agc_abs[i] = s[oAgc + 256 + i];
    }
    for (int i = tid; i < 54; i += kNT) st.dec1_hist[i / 27][i % 27] = s[oD1H + i];
    for (int i = tid; i < 90; i += kNT) st.dec2_hist[i / 45][i % 45] = s[oD2H + i];
    if (tid == 0) {
      st.dc_d1 = s[oMisc + mDcD1];
      st.dc_d2 = s[oMisc + mDcD2];
      st.fast_native = 0;
    }
  }
}
/* int2 history alone (the back kernel's share of PhBlockEnd) */
T41RX_DEV void PhBackBlockEnd(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    for (int h = tid; h < 7; h += kNT) s[oIntH + 24 + h] = s[vInt2 + 2 * kDec + h];
  }
}

/* ------------------------------------------------------------------ */
/* P14/P15: arm_fir_interpolate_f32 x2 (48 taps) and x4 (32 taps), volume */
/* (Process.cpp:917-931)                                                 */
/* ------------------------------------------------------------------ */
/* ------------------------------------------------------------------ */
/* receive equaliser (Filter.cpp:117-165, hook Process.cpp:827-831)      */
/* ------------------------------------------------------------------ */
T41RX_DEV float *EqBand(float *s, int band) { return s + (band < 12 ? vEqBand + kDec * band : oD1I + kDec * (band - 12)); }

/* one lane per (receiver, band): the band's 4-stage transposed-direct-form-II cascade over the 256 demodulated
   samples, arm_biquad_cascade_df2T_f32's operation order (y = b0 x + d1; d1 = (b1 x + a1 y) + d2; d2 = b2 x + a2 y),
   all four stages per sample in registers (each stage sees the same input sequence as stage-by-stage order) */
T41RX_DEV void PhEqBands(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng || u >= 14) return;
  const int sid = Sid(c, g);
  if (!CfgOf(c, g).eq_on) return;
  float *s = Slot(c, g);
  StreamState &st = c.a.st[sid];
  const float *k = c.a.eq_coeffs + 20 * u;
  float b0[4], b1[4], b2[4], a1[4], a2[4], d1[4], d2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    b0[j] = LdgRO(k + 5 * j); b1[j] = LdgRO(k + 5 * j + 1); b2[j] = LdgRO(k + 5 * j + 2);
    a1[j] = LdgRO(k + 5 * j + 3); a2[j] = LdgRO(k + 5 * j + 4);
    d1[j] = st.eq_state[u][2 * j]; d2[j] = st.eq_state[u][2 * j + 1];
  }
  const float *x = s + vAud + 23;
  float *out = EqBand(s, u);
  for (int n = 0; n < kDec; ++n) {
    float v = x[n];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float y = b0[j] * v + d1[j];
      const float t = b1[j] * v + a1[j] * y;
      d1[j] = t + d2[j];
      d2[j] = b2[j] * v + a2[j] * y;
      v = y;
    }
    out[n] = v;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { st.eq_state[u][2 * j] = d1[j]; st.eq_state[u][2 * j + 1] = d2[j]; }
}

/* bands scaled by -/+ level (arm_scale_f32) and added in band order (arm_add_f32 chain, Filter.cpp:151-164) */
T41RX_DEV void PhEqSum(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (!cf.eq_on) return;
  float *s = Slot(c, g);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = u + 64 * r;
    float acc = EqBand(s, 0)[n] * cf.eq_scale[0] + EqBand(s, 1)[n] * cf.eq_scale[1];
    for (int b = 2; b < 14; ++b) acc = acc + EqBand(s, b)[n] * cf.eq_scale[b];
    s[vAud + 23 + n] = acc;
  }
}

/* ------------------------------------------------------------------ */
/* LMS noise reduction / automatic notch (Xanr, Noise.cpp:322-369; hooks Process.cpp:841-865) */
/* ------------------------------------------------------------------ */
/* One pass of the variable-leak LMS over the 256 samples at vAud + 23, one lane per receiver, every operation in the
   reference's order and precision (its unsuffixed literals make part of the arithmetic FP64).  notch: the output
   is the error signal, written back in place; otherwise the reference leaves the filter output in float_buffer_R,
   which nothing reads: only the state advances.  The 64-term sums are the serial bottleneck of this exactness
   path (the throughput kernel has its own form). */
constexpr int vAnrD = vEqBand;                    /* 512: the delay line, staged in shared memory for the block */
constexpr int vAnrW = vEqBand + 512;              /* 64: the taps */

T41RX_DEV void XanrPass(float *aud, float *d, float *w, StreamState &st, bool notch) {
  const int kDelay = 16, kMask = 511, kTaps = 64;
  const float den_mult = 6.25e-10, gamma = 0.1, lidx_min = 120.0, lidx_max = 200.0, lincr = 1.0, ldecr = 3.0,
              two_mu = 0.0001;
  int in_idx = st.anr_in_idx;
  float lidx = st.anr_lidx, ngamma = st.anr_ngamma;
  for (int i = 0; i < kDec; ++i) {
    d[in_idx] = aud[i];
    float y = 0, sigma = 0;
    for (int j = 0; j < kTaps; ++j) {
      const float dv = d[(in_idx + j + kDelay) & kMask];
      y += w[j] * dv;
      sigma += dv * dv;
    }
    const float inv_sigp = 1.0 / (sigma + 1e-10);
    const float error = d[in_idx] - y;
    if (notch) aud[i] = error;
    float nel = error * (1.0 - two_mu * sigma * inv_sigp);
    if (nel < 0.0) nel = -nel;
    float nev = d[in_idx] - (1.0 - two_mu * ngamma) * y - two_mu * error * sigma * inv_sigp;
    if (nev < 0.0) nev = -nev;
    if (nev < nel) {
      if ((lidx += lincr) > lidx_max) lidx = lidx_max;
      else if ((lidx -= ldecr) < lidx_min) lidx = lidx_min;
    }
    ngamma = gamma * (lidx * lidx) * (lidx * lidx) * den_mult;
    const float c0 = 1.0 - two_mu * ngamma;
    const float c1 = two_mu * error * inv_sigp;
    for (int j = 0; j < kTaps; ++j) w[j] = c0 * w[j] + c1 * d[(in_idx + j + kDelay) & kMask];
    in_idx = (in_idx + kMask) & kMask;
  }
  st.anr_in_idx = in_idx;
  st.anr_lidx = lidx;
  st.anr_ngamma = ngamma;
}

/* delay line and taps between HBM and shared memory (dir 0: in, 1: out), 64 threads per receiver */
T41RX_DEV void PhNrStage(Cta &c, int tid, int dir) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (!cf.nr_lms && !cf.anr_notch) return;
  float *s = Slot(c, g);
  StreamState &st = c.a.st[Sid(c, g)];
  for (int i = u; i < 512; i += 64) {
    if (dir == 0) s[vAnrD + i] = st.anr_d[i];
    else st.anr_d[i] = s[vAnrD + i];
  }
  if (dir == 0) s[vAnrW + u] = st.anr_w[u];
  else st.anr_w[u] = s[vAnrW + u];
}

T41RX_DEV void PhNrNotch(Cta &c, int tid) {
  const int g = SerialStream(c, tid);
  if (g < 0) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (!cf.nr_lms && !cf.anr_notch) return;
  float *s = Slot(c, g);
  float *aud = s + vAud + 23;
  StreamState &st = c.a.st[Sid(c, g)];
  if (cf.nr_lms) {                         /* Process.cpp:852-856: Xanr's output is dropped, float_buffer_L x 1.5 */
    XanrPass(aud, s + vAnrD, s + vAnrW, st, false);
    for (int i = 0; i < kDec; ++i) aud[i] = aud[i] * 1.5f;
  }
  if (cf.anr_notch) XanrPass(aud, s + vAnrD, s + vAnrW, st, true);   /* Process.cpp:860-865 */
}

}  // namespace t41rx
#include "rx_nr.cuh"
namespace t41rx {

/* Kim / spectral noise reduction (Process.cpp:845-852), one lane per receiver; scratch over the equaliser's band
   buffers (consumed by then; the LMS stage that follows re-stages its own data there) */
constexpr int vNrBuf = vEqBand, vNrTmp = vNrBuf + 512, vNrOut = vNrTmp + 512, vNrPh = vNrOut + 256;
static_assert(vNrPh + 128 <= 2 * kRawLen, "noise-reduction scratch fits the raw region");
T41RX_DEV void PhNrSpectral(Cta &c, int tid) {
  const int g = SerialStream(c, tid);
  if (g < 0) return;
  const int sid = Sid(c, g);
  const StreamCfg &cf = CfgOf(c, g);
  if ((!cf.nr_kim && !cf.nr_spectral) || !c.a.nr) return;
  float *s = Slot(c, g);
  if (cf.nr_kim) KimNrLane(c.a.nr[sid], cf, c.a.nr_tab, c.a.twiddle, s + vAud + 23, s + vNrBuf, s + vNrTmp, s + vNrOut);
  else SpectralNrLane(c.a.nr[sid], cf, c.a.nr_tab, c.a.twiddle, s + vAud + 23, s + vNrBuf, s + vNrTmp, s + vNrPh);
}
/* noise blanker (Process.cpp:873-876), behind the notch */
T41RX_DEV void PhNoiseBlank(Cta &c, int tid) {
  const int g = SerialStream(c, tid);
  if (g < 0) return;
  const int sid = Sid(c, g);
  if (!CfgOf(c, g).nb_on || !c.a.nr) return;
  float *s = Slot(c, g);
  NoiseBlankLane(c.a.nr[sid], s + vAud + 23, s + vNrBuf, s + vNrTmp);
}

/* CW audio low-pass (Process.cpp:878-914: CW receive state, CWFilterIndex 0..4): 6 transposed-direct-form-II
   biquads over the 256 samples at aud, arm_biquad_cascade_df2T_f32's operation order, all stages per sample in
   registers; st: the selected filter's 12 state values.  One lane. */
T41RX_DEV void CwFilterLane(float *aud, const float *coef, float *st) {
  float b0[6], b1[6], b2[6], a1[6], a2[6], d1[6], d2[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    b0[j] = coef[5 * j]; b1[j] = coef[5 * j + 1]; b2[j] = coef[5 * j + 2]; a1[j] = coef[5 * j + 3]; a2[j] = coef[5 * j + 4];
    d1[j] = st[2 * j]; d2[j] = st[2 * j + 1];
  }
  for (int n = 0; n < kDec; ++n) {
    float v = aud[n];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const float y = b0[j] * v + d1[j];
      const float t = b1[j] * v + a1[j] * y;
      d1[j] = t + d2[j];
      d2[j] = b2[j] * v + a2[j] * y;
      v = y;
    }
    aud[n] = v;
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) { st[2 * j] = d1[j]; st[2 * j + 1] = d2[j]; }
}

T41RX_DEV void PhCwFilter(Cta &c, int tid) {
  const int g = SerialStream(c, tid);
  if (g < 0) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (cf.cw_filter < 0) return;
  CwFilterLane(Slot(c, g) + vAud + 23, c.a.cw_coeffs + 30 * cf.cw_filter, c.a.st[Sid(c, g)].cw_state[cf.cw_filter]);
}

T41RX_DEV void PhInterp1(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    for (int h = tid; h < 23; h += kNT) s[vAud + h] = s[oIntH + h];
  }
}
T41RX_DEV void PhInterp1b(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  float *s = Slot(c, g);
  float taps[kInt1Taps];
#pragma unroll
  for (int i = 0; i < kInt1Taps; ++i) taps[i] = s[oTaps + kTapInt1 + i];
  /* inputs u, u+64, u+128, u+192; two output phases each: 8 independent chains */
  float a0[4], a1[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { a0[r] = 0.0f; a1[r] = 0.0f; }
#pragma unroll
  for (int k = 0; k < 24; ++k) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float x = s[vAud + u + 64 * r + k];      /* oldest-first window of 24 ending at the input */
      a0[r] = fmaf(x, taps[2 * k + 1], a0[r]);        /* phase 0: c[(L-1) + kL] */
      a1[r] = fmaf(x, taps[2 * k], a1[r]);            /* phase 1: c[0 + kL]     */
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = u + 64 * r;
    s[vInt2 + 7 + 2 * n] = a0[r];
    s[vInt2 + 7 + 2 * n + 1] = a1[r];
  }
  for (int h = u; h < 7; h += 64) s[vInt2 + h] = s[oIntH + 24 + h];
}

T41RX_DEV void PhInterp2(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  float *s = Slot(c, g);
  const float volume = CfgOf(c, g).volume;
  float4 *dst = reinterpret_cast<float4 *>(c.a.audio + ((size_t)(Sid(c, g)) * c.a.t_stride + c.t) * kBlock);
  int16_t *dst16 = c.a.audio16 ? c.a.audio16 + ((size_t)(Sid(c, g)) * c.a.t_stride + c.t) * kBlock : nullptr;
  for (int h = u; h < 23; h += 64) s[oIntH + h] = s[vAud + kDec + h];   /* int1 history for the next block */
  float taps[kInt2Taps];
#pragma unroll
  for (int i = 0; i < kInt2Taps; ++i) taps[i] = s[oTaps + kTapInt2 + i];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    /* inputs n = u + 64 r: four outputs each (one float4 store), 16 independent chains */
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[r][p] = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float x = s[vInt2 + u + 64 * (4 * half + r) + k];
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[r][p] = fmaf(x, taps[4 * k + (3 - p)], acc[r][p]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = u + 64 * (4 * half + r);
      const float4 o4 = float4{acc[r][0] * volume, acc[r][1] * volume, acc[r][2] * volume, acc[r][3] * volume};
      if (dst16) {
        dst16[4 * n] = FloatToQ15Word(o4.x); dst16[4 * n + 1] = FloatToQ15Word(o4.y);
        dst16[4 * n + 2] = FloatToQ15Word(o4.z); dst16[4 * n + 3] = FloatToQ15Word(o4.w);
      } else {
        dst[n] = o4;
      }
    }
  }
}

/* int2 history + Codec_gain (Process.cpp:979-1016 with the clip flags never set) */
T41RX_DEV void PhBlockEnd(Cta &c, int tid) {
  for (int g = 0; g < c.ng; ++g) {
    float *s = Slot(c, g);
    for (int h = tid; h < 7; h += kNT) s[oIntH + 24 + h] = s[vInt2 + 2 * kDec + h];
    if (tid == kNT - 1) {
      StreamState &st = c.a.st[Sid(c, g)];
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
/* rows-only kernel: display spectrum of the row-producing blocks for the receivers the            */
/* throughput kernel serves (same phase functions as above; runs BEFORE that kernel in the stream)  */
/* ------------------------------------------------------------------ */
/* DC-block state at the start of block t: the carried state for t = 0, otherwise the state reached by
 * filtering the last 256 Q samples of block t-1 from zero (pole 0.854: converged to the last bit, the
 * same argument as P1's speculative chunks) */
T41RX_DEV void PhRowDcSeed(Cta &c, int tid) {
  const int g = SerialStream(c, tid);
  if (g < 0) return;
  float *s = Slot(c, g);
  const StreamState &st = c.a.st[Sid(c, g)];
  if (c.a.t0 + c.t == 0) {       /* first block of the call (a later launch of a call cut over time finds block t0 - 1 in the buffer) */
    s[oMisc + mDcD1] = st.dc_d1;
    s[oMisc + mDcD2] = st.dc_d2;
    return;
  }
  const DcCoef k = DcCoefs();
  const float rfg = CfgOf(c, g).rf_gain_value;
  const size_t q = ((size_t)Sid(c, g) * c.a.t_stride + (c.t - 1)) * (2 * kBlock) + 2 * (kBlock - kDcWarm) + 1;
  float d1 = 0.0f, lx = 0.0f, ly = 0.0f;
  for (int i = 0; i < kDcWarm; ++i) {
    const float x = IqWord(c.a, q + 2 * i) * rfg;
    ly = DcStep(k, x, d1);
    lx = x;
  }
  s[oMisc + mDcD1] = d1;
  s[oMisc + mDcD2] = DcD2(lx, ly);
}

#ifndef T41RX_HOST_EMUL
/* ---- device-only throughput forms of the two serial pieces of the rows-only kernel ---- */
#ifndef T41RX_ROWS_LAYOUT
constexpr int vRowTail = oOla;                    /* 256 floats: last Q samples of the previous block */
#else
constexpr int vRowTail = vDcSpec;                 /* (consumed by the seed before PhDcWarm writes the chunk states) */
static_assert(kDcWarm <= 4 * kDcChunks, "the tail fits the chunk-state area");
#endif

/* stage the 256 samples PhRowDcSeed filters (coalesced instead of 256 dependent global loads) */
T41RX_DEV void PhRowTailLoad(Cta &c, int tid) {
  if (c.a.t0 + c.t == 0 || c.dc_carried) return;
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  float *s = Slot(c, g);
  const size_t q = ((size_t)Sid(c, g) * c.a.t_stride + (c.t - 1)) * (2 * kBlock) + 2 * (kBlock - kDcWarm);
  for (int i = u; i < kDcWarm; i += 64) s[vRowTail + i] = IqWord(c.a, q + 2 * i + 1);
}
T41RX_DEV void PhRowDcSeedFast(Cta &c, int tid) {
  const int g = SerialStream(c, tid);
  if (g < 0) return;
  if (c.dc_carried) return;      /* the block before this one was a row block of this launch: PhDcVerify / PhDcFix left the state */
  float *s = Slot(c, g);
  const StreamState &st = c.a.st[Sid(c, g)];
  if (c.a.t0 + c.t == 0) {       /* first block of the call (a later launch of a call cut over time finds block t0 - 1 in the buffer) */
    s[oMisc + mDcD1] = st.dc_d1;
    s[oMisc + mDcD2] = st.dc_d2;
    return;
  }
  const DcCoef k = DcCoefs();
  const float rfg = CfgOf(c, g).rf_gain_value;
  float d1 = 0.0f, lx = 0.0f, ly = 0.0f;
  for (int i0 = 0; i0 < kDcWarm; i0 += 8) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = s[vRowTail + i0 + j] * rfg;
#pragma unroll
    for (int j = 0; j < 8; ++j) { ly = DcStep(k, x[j], d1); lx = x[j]; }
  }
  s[oMisc + mDcD1] = d1;
  s[oMisc + mDcD2] = DcD2(lx, ly);
}

/* ZoomFFTExe's input: IQ correction + Fs/4 shift, in place (FFT.cpp:83-84 run on the *_EX buffers) */
T41RX_DEV void PhZoomShift(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (cf.zoom == 0) return;
  float *s = Slot(c, g);
  const IqFix fix = IqFixOf(cf);
  for (int n = u; n < kBlock; n += 64) {
    float vi = s[oRawI + 27 + n], vq = s[oRawQ + 27 + n];
    IqCorr(fix, vi, vq);
    QuarterShift(n, vi, vq);
    s[oRawI + 27 + n] = vi;
    s[oRawQ + 27 + n] = vq;
  }
}

/* The 4-stage elliptic DF1 cascade (arm_biquad_cascade_df1_f32, FFT.cpp:83-84) as blocked linear scans:
 * one warp per channel, lane L owns samples 65 L .. 65 L + 64 (odd stride: conflict-free), every stage =
 * zero-state pass, 2x2 carry scan over the lanes, correction pass with the two homogeneous responses.
 * FP32 results differ from the sample-by-sample form in the last bits (rows tolerance: 1 LSB). */
T41RX_DEV void PhZoomIirScan(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (cf.zoom == 0) return;                          /* warp-uniform */
  StreamState &st = c.a.st[Sid(c, g)];
  const int chn = u >> 5, lane = u & 31;
  float *x = Slot(c, g) + (chn ? oRawQ : oRawI) + 27;
  constexpr int kLen = 65;
  const int n0 = kLen * lane;
  const int n1 = (n0 + kLen < kBlock) ? n0 + kLen : kBlock;
  const unsigned full = 0xffffffffu;
  for (int sg = 0; sg < 4; ++sg) {
    const float *kk = c.a.zoom_iir + (cf.zoom - 1) * 20 + 5 * sg;
    const float b0 = LdgRO(kk), b1 = LdgRO(kk + 1), b2 = LdgRO(kk + 2), a1 = LdgRO(kk + 3), a2 = LdgRO(kk + 4);
    /* inputs just before the chunk (previous lane's last two, or the carried state) */
    float x1 = (lane == 0) ? st.zoom_iir[chn][4 * sg] : x[n0 - 1];
    float x2 = (lane == 0) ? st.zoom_iir[chn][4 * sg + 1] : ((n0 >= 2) ? x[n0 - 2] : 0.0f);
    const float sx1 = x[kBlock - 1], sx2 = x[kBlock - 2];       /* this stage's input history for the next block */
    __syncwarp(full);
    /* zero-state pass, in place */
    float y1 = 0.0f, y2 = 0.0f;
    for (int n = n0; n < n1; ++n) {
      const float xn = x[n];
      float f = b0 * xn;
      f = fmaf(b1, x1, f);
      f = fmaf(b2, x2, f);
      const float y = fmaf(a1, y1, fmaf(a2, y2, f));
      x[n] = y;
      x2 = x1; x1 = xn;
      y2 = y1; y1 = y;
    }
    /* chunk transfer matrix M = A^65 from the homogeneous responses */
    float u1 = 1.0f, u2 = 0.0f, v1 = 0.0f, v2 = 1.0f;
    for (int j = 0; j < kLen; ++j) {
      const float un = fmaf(a1, u1, a2 * u2), vn = fmaf(a1, v1, a2 * v2);
      u2 = u1; u1 = un;
      v2 = v1; v1 = vn;
    }
    float m00 = u1, m01 = v1, m10 = u2, m11 = v2;
    float s1 = y1, s2 = y2;
    if (lane == 0) {
      const float c1 = st.zoom_iir[chn][4 * sg + 2], c2 = st.zoom_iir[chn][4 * sg + 3];
      s1 += m00 * c1 + m01 * c2;
      s2 += m10 * c1 + m11 * c2;
    }
    for (int d = 1; d < 32; d <<= 1) {
      const float t1 = __shfl_up_sync(full, s1, d), t2 = __shfl_up_sync(full, s2, d);
      if (lane >= d) {
        s1 += m00 * t1 + m01 * t2;
        s2 += m10 * t1 + m11 * t2;
      }
      const float q00 = m00 * m00 + m01 * m10, q01 = m00 * m01 + m01 * m11;
      const float q10 = m10 * m00 + m11 * m10, q11 = m10 * m01 + m11 * m11;
      m00 = q00; m01 = q01; m10 = q10; m11 = q11;
    }
    float c1 = __shfl_up_sync(full, s1, 1), c2 = __shfl_up_sync(full, s2, 1);
    if (lane == 0) { c1 = st.zoom_iir[chn][4 * sg + 2]; c2 = st.zoom_iir[chn][4 * sg + 3]; }
    /* correction pass: y[n] += h1[j] c1 + h2[j] c2 */
    u1 = 1.0f; u2 = 0.0f; v1 = 0.0f; v2 = 1.0f;
    float yl1 = 0.0f, yl2 = 0.0f;
    for (int n = n0; n < n1; ++n) {
      const float un = fmaf(a1, u1, a2 * u2), vn = fmaf(a1, v1, a2 * v2);
      u2 = u1; u1 = un;
      v2 = v1; v1 = vn;
      const float y = x[n] + (un * c1 + vn * c2);
      x[n] = y;
      yl2 = yl1; yl1 = y;
    }
    if (lane == 31) {
      st.zoom_iir[chn][4 * sg] = sx1;
      st.zoom_iir[chn][4 * sg + 1] = sx2;
      st.zoom_iir[chn][4 * sg + 2] = yl1;
      st.zoom_iir[chn][4 * sg + 3] = yl2;
    }
    __syncwarp(full);
  }
}

/* The same cascade, sample by sample and rounded exactly like the reference: eight lanes per receiver =
 * (channel, biquad stage), the four stages of a channel working IN PLACE on the same array (as CMSIS cascades do),
 * stage s a fixed kSkew samples behind stage s - 1.  Nothing crosses lanes inside a step (a stage finds its input in
 * shared memory, written kSkew steps earlier by its predecessor and fetched four steps ahead of use), so a step is
 * as long as one biquad's own recurrence: acc3 = acc2 + a1 y1, acc = acc3 + a2 y2.  Input: the shifted samples
 * (PhZoomShift); output: the last stage's, in place. */
T41RX_DEV void PhZoomIirPipe(Cta &c, int tid) {
  /* ONE warp serves the CTA's receivers, eight lanes each.  Lane (receiver g, channel c, stage s) touches bank
     8 g + 28 c - kSkew s (mod 32) at every step: with an ODD skew these are 32 different banks (an even one leaves
     only the multiples of 4: a 4-way conflict on every load and store of the loop).
     Where the time of a step goes (tools/microbench/casc_loop.cu, one warp, clocks per step): the y-chain alone 12.2, with
     the three input terms 14.4, with the fetch 14.5, with the store of the step's own result 18.0, and 20.8 if that
     store is predicated (the form of the first two rounds).  The result of a step is the last link of the chain: a
     store of it sits in the issue order right where the next step's multiply wants to go.  So the main loop stores
     the result of two steps ago (a register that has been ready for 24 clocks) and never predicates: lanes without a
     receiver, or at zoom x1, run on a dead stretch of their slot instead.  15.5 clocks per step. */
  static_assert(kG * 8 <= 32, "eight lanes per receiver in one warp");
  constexpr int kSkew = 15, kAhead = 4, kGroup = 8, kLate = 2, kEdgeSteps = 48;
  static_assert(kGroup <= kSkew - kAhead - kLate, "a stage never fetches what its predecessor has not written yet");
  static_assert(kSlot % 32 == 8 && (oRawQ - oRawI) % 32 == 28 && kSkew % 2 == 1, "bank map of the cascade lanes");
  /* the stretch inactive lanes work on: their own slot's raw tile (no receiver in the slot, or one at zoom x1 whose
     spectrum input has been taken from the tile already: PhSpecWindow part 1 runs before this phase) */
  constexpr int kDummy = oRawI + 27 + 3 * kSkew;
  static_assert(kDummy + kBlock + 3 * kSkew + kAhead + 4 <= 2 * kRawLen, "dummy stretch inside the raw tile");
  if ((tid >> 5) != c.casc_warp) return;
  const int lane = tid & 31, g = lane >> 3;
  const bool have = g < c.ng;
  const int gg = have ? g : 0;
  const StreamCfg &cf = CfgOf(c, gg);
  StreamState &st = c.a.st[Sid(c, gg)];
  const bool active = have && cf.zoom != 0;
  if (!__any_sync(0xffffffffu, active)) return;       /* every receiver of the CTA at zoom x1: no cascade (CalcZoom1Magn) */
  const int chn = (lane >> 2) & 1, sg = lane & 3;
  float *x = Slot(c, g) + (active ? (chn ? oRawQ : oRawI) + 27 : kDummy);    /* (slot g exists even without a receiver) */
  float b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0, x1 = 0, x2 = 0, y1 = 0, y2 = 0;
  if (active) {
    const float *kk = c.a.zoom_iir + (cf.zoom - 1) * 20 + 5 * sg;
    b0 = LdgRO(kk); b1 = LdgRO(kk + 1); b2 = LdgRO(kk + 2); a1 = LdgRO(kk + 3); a2 = LdgRO(kk + 4);
    x1 = st.zoom_iir[chn][4 * sg]; x2 = st.zoom_iir[chn][4 * sg + 1];
    y1 = st.zoom_iir[chn][4 * sg + 2]; y2 = st.zoom_iir[chn][4 * sg + 3];
  }
  float *xm = x - kSkew * sg;                         /* this stage's sample at step k is xm[k] */
  const int lo = kSkew * sg - 27;                     /* xm[lo] is the first float of the raw region: early fetches stop there */
  float xn0 = xm[max(0, lo)], xn1 = xm[max(1, lo)], xn2 = xm[max(2, lo)], xn3 = xm[max(3, lo)];
  /* one step; kEdge: some stages are outside the block (the first and the last 3 kSkew steps), the result is stored at
     once; else: every stage is inside the block (the state moves are plain register renames) and the store is the
     result of kLate = 2 steps ago.
     The sum is ((((b0 x + b1 x1) + b2 x2) + a1 y1) + a2 y2) in this order (arm_biquad_cascade_df1_f32); its first
     three terms do not involve the recurrence and are formed one step ahead (pre), beside the previous step's
     y-chain */
#define T41RX_ZOOM_STEP(kEdge)                                                     \
  {                                                                                \
    const float xin = xn0;                                                         \
    float acc = pre + a1 * y1;                                                     \
    acc = acc + a2 * y2;                                                           \
    if (kEdge) {                                                                   \
      const int n = k - kSkew * sg;                                                \
      if (n >= 0 && n < kBlock) {                                                  \
        x2 = x1;                                                                   \
        x1 = xin;                                                                  \
        y2 = y1;                                                                   \
        y1 = acc;                                                                  \
        xm[k] = acc;                                                               \
      }                                                                            \
    } else {                                                                       \
      xm[k - kLate] = y2;                                                          \
      x2 = x1;                                                                     \
      x1 = xin;                                                                    \
      y2 = y1;                                                                     \
      y1 = acc;                                                                    \
    }                                                                              \
    pre = b0 * xn1;                                                                \
    pre = pre + b1 * x1;                                                           \
    pre = pre + b2 * x2;                                                           \
    xn0 = xn1; xn1 = xn2; xn2 = xn3;                                               \
    xn3 = xm[kEdge ? max(k + kAhead, lo) : k + kAhead];                            \
  }
  float pre = b0 * xn0;
  pre = pre + b1 * x1;
  pre = pre + b2 * x2;
  /* a warp-level barrier every kGroup steps keeps the compiler from moving a fetch above the predecessor's store it
     has to see (that store is kSkew - kAhead - kLate steps older than the fetch); inside such a group any order is
     fine */
  static_assert(kEdgeSteps >= 3 * kSkew + kLate && kEdgeSteps % 4 == 0 && (kBlock - kEdgeSteps) % kGroup == 0, "groups");
  int k = 0;
  for (; k < kEdgeSteps;) {                           /* until every stage has been inside the block for kLate steps */
    for (int u_ = 0; u_ < 4; ++u_, ++k) T41RX_ZOOM_STEP(true)
    __syncwarp();
  }
#pragma unroll 1
  for (; k < kBlock;) {                               /* (its first two stores repeat what the edge steps stored) */
#pragma unroll
    for (int u_ = 0; u_ < kGroup; ++u_, ++k) T41RX_ZOOM_STEP(false)
    __syncwarp();
  }
  xm[kBlock - 2] = y2;                                /* the two results the main loop still owes */
  xm[kBlock - 1] = y1;
  __syncwarp();
  for (; k < kBlock + 3 * kSkew;) {
    for (int u_ = 0; u_ < 4; ++u_, ++k) T41RX_ZOOM_STEP(true)
    __syncwarp();
  }
#undef T41RX_ZOOM_STEP
  if (active) {
    st.zoom_iir[chn][4 * sg] = x1; st.zoom_iir[chn][4 * sg + 1] = x2;
    st.zoom_iir[chn][4 * sg + 2] = y1; st.zoom_iir[chn][4 * sg + 3] = y2;
  }
}

/* 4-tap FIR decimation by 2^zoom of the cascade's output, first zoom_samples outputs into the 512-deep ring
 * (arm_fir_decimate_f32, FFT.cpp:86-101) */
T41RX_DEV void PhZoomDecimate(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (cf.zoom == 0) return;
  StreamState &st = c.a.st[Sid(c, g)];
  const float *s = Slot(c, g);
  const int M = 1 << cf.zoom, zs = cf.zoom_samples, ptr = st.zoom_ptr;
  const float f0 = cf.zoom_fir[0], f1 = cf.zoom_fir[1], f2 = cf.zoom_fir[2], f3 = cf.zoom_fir[3];
  for (int e = u; e < 2 * zs; e += 64) {
    const int chn = e >= zs, k = e - (chn ? zs : 0);        /* e < 2 zs */
    const float *y = s + (chn ? oRawQ : oRawI) + 27;
    const int n = k * M;
    const float h0 = (n >= 3) ? y[n - 3] : st.zoom_fir_hist[chn][n], h1 = (n >= 2) ? y[n - 2] : st.zoom_fir_hist[chn][n + 1],
                h2 = (n >= 1) ? y[n - 1] : st.zoom_fir_hist[chn][n + 2];
    float o = 0.0f;
    o = fmaf(h0, f0, o);
    o = fmaf(h1, f1, o);
    o = fmaf(h2, f2, o);
    o = fmaf(y[n], f3, o);
    st.zoom_ring[chn][(ptr + k) & (kSpecRes - 1)] = o;
  }
}
/* history of the decimator and ring pointer (after every lane has read the old values) */
T41RX_DEV void PhZoomDecimateEnd(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const StreamCfg &cf = CfgOf(c, g);
  if (cf.zoom == 0) return;
  StreamState &st = c.a.st[Sid(c, g)];
  const float *s = Slot(c, g);
  if (u < 6) {
    const int chn = u / 3, i = u % 3;
    st.zoom_fir_hist[chn][i] = s[(chn ? oRawQ : oRawI) + 27 + kBlock - 3 + i];
  }
  if (u == 63) st.zoom_ptr = (st.zoom_ptr + cf.zoom_samples) & (kSpecRes - 1);
}

#define T41RX_ROWS_SCHEDULE_FAST(RX_PHASE)                               \
  RX_PHASE(PhLoad(c, tid); PhRowTailLoad(c, tid));                       \
  RX_PHASE(PhRowDcSeedFast(c, tid));                                     \
  RX_PHASE(PhDcWarm(c, tid));                                            \
  RX_PHASE(PhDcMain(c, tid));                                            \
  RX_PHASE(PhDcVerify(c, tid));                                          \
  RX_PHASE(PhDcFix(c, tid));                                             \
  RX_PHASE(PhZoomShift(c, tid); PhSpecWindow(c, tid, 1));                \
  if (c.a.flags & 4u) {                                                  \
    RX_PHASE(PhZoomIirScan(c, tid));                                     \
  } else {                                                               \
    RX_PHASE(PhZoomIirPipe(c, tid));                                     \
  }                                                                      \
  RX_PHASE(PhZoomDecimate(c, tid));                                      \
  RX_PHASE(PhZoomDecimateEnd(c, tid));                                   \
  RX_PHASE(PhSpecWindow(c, tid, 2));                                     \
  RX_PHASE(PhSpecFftPass(c, tid, 0));                                    \
  RX_PHASE(PhSpecFftPass(c, tid, 1));                                    \
  RX_PHASE(PhSpecFftPass(c, tid, 2));                                    \
  RX_PHASE(PhSpecRow(c, tid));
#endif

/* ------------------------------------------------------------------ */
/* audio-spectrum + S-meter by-product (Process.cpp:550-570, NFM :791-805) */
/* ------------------------------------------------------------------ */
/* One row of one receiver per 64 threads, from the masked spectrum the chain kernels left in a.aspec:
 *   audioSpectBuffer[1023 - k] = iFFT_buffer[k]^2 (each float of the 512 complex bins on its own);
 *   audioYPixel[k] = offset + map(15 log10f(3-point average), 0, 100, 0, 120), offset 50 (20 for NFM), read from
 *   the top of the reversed array except for LSB; clamped at 0; k < 270;
 *   audioMaxSquaredAve = .5 max + .5 audioMaxSquaredAve (FP64 sum).
 * Arduino map() with a float argument computes in float (Teensyduino core, wiring.h).  Modes without the
 * filter (PSK31 raw mode) leave both untouched: their rows repeat the state.
 * Shared memory of slot g: words 0..1023 the reversed squares, 1024..1087 partial maxima. */
T41RX_DEV bool AudioSpecUpdates(int mode) { return UsesFilter(mode); }

T41RX_DEV void PhAudioSquares(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const int sid = Sid(c, g);
  if (!AudioSpecUpdates(CfgOf(c, g).mode)) return;
  float *s = Slot(c, g);
  const float *src = reinterpret_cast<const float *>(c.a.aspec + ((size_t)sid * c.a.n_rows + c.row_idx) * kFft);
  float m = 0.0f;
  for (int j = 0; j < 16; ++j) {
    const int k = u + 64 * j;
    const float v = src[k];
    const float q = v * v;
    s[2 * kFft - 1 - k] = q;
    m = (q > m) ? q : m;                    /* squares are >= 0: arm_max_f32's result */
  }
  s[2 * kFft + u] = m;
}

T41RX_DEV void PhAudioPixels(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const int sid = Sid(c, g);
  const int mode = CfgOf(c, g).mode;
  StreamState &st = c.a.st[sid];
  const float *sq = Slot(c, g);
  int32_t *out = c.a.audio_ypixel ? c.a.audio_ypixel + ((size_t)sid * c.a.n_rows + c.row_idx) * kAudioSpecPixels : nullptr;
  const bool upd = AudioSpecUpdates(mode);
  const float offset = (mode == kModeNfm) ? 20.0f : 50.0f;
  for (int k = u; k < kAudioSpecPixels; k += 64) {
    int pix;
    if (upd) {
      float sum;
      if (mode == kModeLsb) sum = (sq[k] + sq[k + 1]) + sq[k + 2];
      else sum = (sq[1021 - k] + sq[1022 - k]) + sq[1023 - k];
      const float x = 15.0f * log10f(sum / 3.0f);
      const float v = offset + ((x - 0.0f) * (120.0f - 0.0f) / (100.0f - 0.0f) + 0.0f);
      pix = (v >= 0.0f) ? (int)v : 0;       /* -inf / NaN (empty spectrum) end up 0 like on the targets */
      st.audio_ypixel[k] = (int16_t)pix;
    } else {
      pix = st.audio_ypixel[k];
    }
    if (out) out[k] = pix;
    /* Process.cpp:818-825: the same pixels, limited to a byte, go to the control app */
    if (c.a.audio_frames)
      c.a.audio_frames[((size_t)sid * c.a.n_rows + c.row_idx) * kAudioSpecPixels + k] = (uint8_t)(pix > 255 ? 255 : pix);
  }
  if (u == 0) {
    float ave = st.audio_max_sq_ave;
    if (upd) {
      const float *pm = sq + 2 * kFft;
      float mx = pm[0];
      for (int i = 1; i < 64; ++i) mx = (pm[i] > mx) ? pm[i] : mx;
      /* arm_max_f32 starts from element 0 and replaces it only by larger values: a NaN there (NFM discriminator
         on exact silence) is the result, NaNs elsewhere are skipped like above */
      if (sq[0] != sq[0]) mx = sq[0];
      ave = (float)(.5 * (double)mx + .5 * (double)ave);
      st.audio_max_sq_ave = ave;
    }
    if (c.a.audio_max_ave) c.a.audio_max_ave[(size_t)sid * c.a.n_rows + c.row_idx] = ave;
  }
}

#define T41RX_AUDIO_SPEC_SCHEDULE(RX_PHASE) \
  RX_PHASE(PhAudioSquares(c, tid));         \
  RX_PHASE(PhAudioPixels(c, tid));

/* ------------------------------------------------------------------ */
/* spectrum frame for the PC control app (FFT.cpp:142-194)              */
/* ------------------------------------------------------------------ */
/* "FD" + "%03d" of (255 - max) + 512 bytes + ';' from the pixelnew row the spectrum phases wrote: data = pixelnew +
 * currentNF (int16), max starts at 0, byte = max(data + 255 - max, 0).  Only ZoomFFTExe (zoom != x1) sends it:
 * zoom x1 rows are all zeros.  Shared memory of slot g: words 0..63 partial maxima. */
T41RX_DEV void PhSpecFrameMax(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const int sid = Sid(c, g);
  const StreamCfg &cf = CfgOf(c, g);
  if (cf.zoom == 0) return;
  const int16_t *row = c.a.spec_rows + ((size_t)sid * c.a.n_rows + c.row_idx) * kSpecRes;
  int m = 0;
  for (int j = 0; j < 8; ++j) {
    const int d = (int16_t)((int)row[u + 64 * j] + cf.current_nf);
    m = d > m ? d : m;
  }
  reinterpret_cast<int *>(Slot(c, g))[u] = m;
}

T41RX_DEV void PhSpecFrameWrite(Cta &c, int tid) {
  const int g = tid >> 6, u = tid & 63;
  if (g >= c.ng) return;
  const int sid = Sid(c, g);
  const StreamCfg &cf = CfgOf(c, g);
  uint8_t *f = c.a.spec_frames + ((size_t)sid * c.a.n_rows + c.row_idx) * kSpecFrameBytes;
  if (cf.zoom == 0) {                       /* CalcZoom1Magn sends nothing */
    for (int i = u; i < kSpecFrameBytes; i += 64) f[i] = 0;
    return;
  }
  const int *pm = reinterpret_cast<const int *>(Slot(c, g));
  int mx = 0;
  for (int i = 0; i < 64; ++i) mx = pm[i] > mx ? pm[i] : mx;
  const int16_t *row = c.a.spec_rows + ((size_t)sid * c.a.n_rows + c.row_idx) * kSpecRes;
  for (int j = 0; j < 8; ++j) {
    const int x = u + 64 * j;
    int t = (int16_t)((int)row[x] + cf.current_nf) + 255 - mx;
    if (t < 0) t = 0;
    f[5 + x] = (uint8_t)t;
  }
  if (u == 0) {
    /* the first three characters sprintf("%03d") gives for 255 - max (a fourth one would be overwritten) */
    const int v = 255 - mx;
    char h0, h1, h2;
    if (v >= 0) {
      h0 = (char)('0' + (v / 100) % 10); h1 = (char)('0' + (v / 10) % 10); h2 = (char)('0' + v % 10);
    } else {
      int n = -v;
      while (n >= 100) n /= 10;
      h0 = '-'; h1 = (char)('0' + n / 10); h2 = (char)('0' + n % 10);
    }
    f[0] = 'F'; f[1] = 'D'; f[2] = (uint8_t)h0; f[3] = (uint8_t)h1; f[4] = (uint8_t)h2;
    f[kSpecFrameBytes - 1] = ';';
  }
}

#define T41RX_SPEC_FRAME_SCHEDULE(RX_PHASE) \
  RX_PHASE(PhSpecFrameMax(c, tid));         \
  RX_PHASE(PhSpecFrameWrite(c, tid));

#define T41RX_ROWS_SCHEDULE(RX_PHASE)                                    \
  RX_PHASE(PhLoad(c, tid); PhRowDcSeed(c, tid));                         \
  RX_PHASE(PhDcWarm(c, tid));                                            \
  RX_PHASE(PhDcMain(c, tid));                                            \
  RX_PHASE(PhDcVerify(c, tid));                                          \
  RX_PHASE(PhDcFix(c, tid));                                             \
  RX_PHASE(PhZoomIir(c, tid));                                           \
  RX_PHASE(PhSpecWindow(c, tid));                                        \
  RX_PHASE(PhSpecFftPass(c, tid, 0));                                    \
  RX_PHASE(PhSpecFftPass(c, tid, 1));                                    \
  RX_PHASE(PhSpecFftPass(c, tid, 2));                                    \
  RX_PHASE(PhSpecRow(c, tid));

/* ------------------------------------------------------------------ */
/* the block schedule; RX_PHASE(stmt) runs stmt for every tid then syncs */
/* ------------------------------------------------------------------ */
#define T41RX_BLOCK_SCHEDULE(RX_PHASE)                                   \
  RX_PHASE(PhLoad(c, tid));                                              \
  RX_PHASE(PhDcWarm(c, tid));                                            \
  RX_PHASE(PhDcMain(c, tid));                                            \
  RX_PHASE(PhDcVerify(c, tid));                                          \
  RX_PHASE(PhDcFix(c, tid));                                             \
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
  RX_PHASE(PhAgcMaxA(c, tid));                                           \
  RX_PHASE(PhAgcMaxB(c, tid));                                           \
  RX_PHASE(PhAgcMaxC(c, tid));                                           \
  RX_PHASE(PhAgcSerial(c, tid));                                         \
  RX_PHASE(PhAgcPost(c, tid));                                           \
  RX_PHASE(PhDemodParallel(c, tid); PhInterp1(c, tid));                  \
  RX_PHASE(PhDemodSerial(c, tid));                                       \
  RX_PHASE(PhEqBands(c, tid));                                           \
  RX_PHASE(PhEqSum(c, tid));                                             \
  RX_PHASE(PhNrSpectral(c, tid));                                        \
  RX_PHASE(PhNrStage(c, tid, 0));                                        \
  RX_PHASE(PhNrNotch(c, tid));                                           \
  RX_PHASE(PhNrStage(c, tid, 1));                                        \
  RX_PHASE(PhNoiseBlank(c, tid));                                        \
  RX_PHASE(PhCwFilter(c, tid));                                          \
  RX_PHASE(PhInterp1b(c, tid));                                          \
  RX_PHASE(PhInterp2(c, tid));                                           \
  RX_PHASE(PhBlockEnd(c, tid));

/* the same block as three kernels (see "Split form of the chain") */
#define T41RX_FRONT_SCHEDULE(RX_PHASE)                                   \
  RX_PHASE(PhLoad(c, tid));                                              \
  RX_PHASE(PhDcWarm(c, tid));                                            \
  RX_PHASE(PhDcMain(c, tid));                                            \
  RX_PHASE(PhDcVerify(c, tid));                                          \
  RX_PHASE(PhDcFix(c, tid));                                             \
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
  RX_PHASE(PhAgcMaxA(c, tid));                                           \
  RX_PHASE(PhAgcMaxB(c, tid));                                           \
  RX_PHASE(PhAgcMaxC(c, tid));                                           \
  RX_PHASE(PhSerialStore(c, tid));                                       \
  RX_PHASE(PhSerialRing(c, tid); PhCodecGain(c, tid));

#define T41RX_BACK_SCHEDULE(RX_PHASE)                                    \
  RX_PHASE(PhBackLoad(c, tid));                                          \
  RX_PHASE(PhEqBands(c, tid));                                           \
  RX_PHASE(PhEqSum(c, tid));                                             \
  RX_PHASE(PhNrSpectral(c, tid));                                        \
  RX_PHASE(PhNrStage(c, tid, 0));                                        \
  RX_PHASE(PhNrNotch(c, tid));                                           \
  RX_PHASE(PhNrStage(c, tid, 1));                                        \
  RX_PHASE(PhNoiseBlank(c, tid));                                        \
  RX_PHASE(PhCwFilter(c, tid));                                          \
  RX_PHASE(PhInterp1b(c, tid));                                          \
  RX_PHASE(PhInterp2(c, tid));                                           \
  RX_PHASE(PhBackBlockEnd(c, tid));

}  // namespace t41rx
#endif
