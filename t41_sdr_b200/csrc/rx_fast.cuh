/*
 * rx_fast.cuh — t41rx_stream_rx_kernel: the throughput form of the fused T41 receive chain.
 *
 * Same chain as rx_phases.cuh (ProcessIQData(), reference Process.cpp:70-944), same HBM state
 * (StreamState) and tables, different mapping:
 *   - one CTA owns G <= kFastMaxG receivers for all n_blocks blocks of a launch;
 *   - ONE WARP PER RECEIVER runs every sample-parallel stage (only __syncwarp inside a block);
 *   - one extra warp runs the AGC envelope state machine with lane = receiver;
 *   - blocks are software-pipelined with one __syncthreads per block: in superstep k a receiver
 *     warp does back end(k-2) then front end(k), the AGC warp does AGC(k-1).
 * Linear recurrences (DC block, AM detector filters) are blocked scans, the NCO is a closed-form
 * FP32 phasor driven by an FP64 block phasor, the fast-convolution filter fuses the last forward
 * FFT pass, the mask multiply and the first inverse pass in registers.  FP32 with FMA contraction:
 * results agree with the oracle to the stated tolerance (audio SNR >= 90 dB; in practice > 110 dB),
 * not bit for bit; the bit-exact kernel is t41rx_fused_rx_kernel.
 *
 * Row-producing blocks (display spectrum) and amplitude-transient blocks of the oscillator take
 * slower side paths inside the same kernel.
 */
#ifndef T41RX_FAST_CUH
#define T41RX_FAST_CUH

#include "rx_phases.cuh"

namespace t41rx {
namespace fast {

constexpr int kFastMaxG = 7;          /* receivers per CTA (shared memory: 7 x 31.5 KB) */
constexpr unsigned kFull = 0xffffffffu;

/* ---- per-receiver shared-memory slot, in 4-byte words ---- */
constexpr int kRawChunkWords = 36;                    /* 16 samples (32 words) + 4 pad: conflict-free LDS.128 */
constexpr int kRawBufWords = 32 * kRawChunkWords;     /* one 512-sample quarter: 1152 */
constexpr int oRaw = 0;                               /* 2 quarter buffers */
constexpr int oMix = oRaw + 2 * kRawBufWords;         /* 2304: mixed samples, 8 planes (ch, phase) x 136; FFT buffer overlay */
constexpr int kMixPlane = 136;                        /* 8 history + 128 */
constexpr int kMixWords = 1152;                       /* >= 8 * 136 = 1088 and >= 2 * 576 (padded FFT buffer) */
constexpr int oD1 = oMix + kMixWords;                 /* 3456: dec1 output, 4 planes (ch, parity) x 280; overlays below */
constexpr int kD1Plane = 280;                         /* 24 history + 256 */
constexpr int kD1Words = 1152;
constexpr int oOlaF = oD1 + kD1Words;                 /* 4608: previous 256 complex filter inputs (interleaved) */
constexpr int oStA = oOlaF + 512;                     /* 5120: 2 x (|z| delayed [256], window max -> volts [256]) */
constexpr int oStZ = oStA + 1024;                     /* 6144: 2 x 256 complex: delayed filter output */
constexpr int oZH = oStZ + 1024;                      /* 7168: last 97 complex filter outputs */
constexpr int oAH = oZH + 196;                        /* 7364: last 97 |z| */
constexpr int oTapsF = oAH + 100;                     /* 7464: dec1 28 | dec2 46 | int1 48 | int2 32 (+2) */
constexpr int oNcoW = oTapsF + 156;                   /* 7620: W[16] float2, Q[4] float2 */
constexpr int oIH = oNcoW + 40;                       /* 7660: int1 history 23 (24) | int2 history 7 (8) */
constexpr int oMH = oIH + 32;                         /* 7692: dec1 history, 8 planes x 8 */
constexpr int oDH = oMH + 64;                         /* 7756: dec2 history, 4 planes x 24 */
constexpr int oMiscF = oDH + 96;                      /* 7852 */
constexpr int kSlotF = oMiscF + 16;                   /* 7868 == 4 (mod 8): the AGC warp's LDS.128 hit distinct banks */
static_assert(kSlotF % 8 == 4, "slot stride");
static_assert((oStA % 4) == 0 && (oStZ % 4) == 0 && (oRaw % 4) == 0 && (oMix % 4) == 0 && (oD1 % 4) == 0, "16-byte alignment");

/* overlays on the dec1 region once dec2 has consumed it (front end) */
constexpr int vE = oD1;                               /* |z| extended: 97 history + 256 new (360) */
constexpr int vSfx = vE + 360;                        /* suffix maxima inside chunks of 8 */
constexpr int vPfx = vSfx + 360;                      /* prefix maxima */
constexpr int vCM = vPfx + 360;                       /* chunk maxima (45 -> 48) then 11-chunk window maxima */
static_assert(vCM + 48 <= oD1 + kD1Words, "sliding max scratch");
/* overlays on the dec1 region in the back end */
constexpr int vAudF = oD1;                            /* 24 (1 pad + 23 history) + 256 demodulated samples */
constexpr int vI1 = oD1 + 288;                        /* 8 (1 pad + 7 history) + 512 */
static_assert(vI1 + 520 <= oD1 + kD1Words, "back-end scratch");

/* the 512-point FFT buffer: element i lives at float2 index i + (i >> 3) */
__device__ __forceinline__ int FPos(int i) { return i + (i >> 3); }

constexpr float kDcA1 = 0.854352383886757938f;        /* FIR.cpp:87-89 */
constexpr float kDcB0 = 0.927176191943378969f;

__device__ __forceinline__ float PowA1(int n) {       /* a1^n, n >= 0, by squaring (set-up only) */
  float r = 1.0f, b = kDcA1;
  while (n) {
    if (n & 1) r *= b;
    b *= b;
    n >>= 1;
  }
  return r;
}

#ifdef T41RX_FAST_TIMING
/* developer build: cycles per section, CTA 0, receiver warp 0 (slots 0..15) and the AGC warp (16..19) */
__device__ unsigned long long g_fast_cycles[32];
struct SectionTimer {
  long long mark;
  bool on;
  int base;
  __device__ __forceinline__ void Start(bool enable, int b) { on = enable; base = b; mark = clock64(); }
  __device__ __forceinline__ void Lap(int slot) {
    if (on) {
      const long long now = clock64();
      g_fast_cycles[base + slot] += (unsigned long long)(now - mark);
      mark = now;
    }
  }
};
#define T41RX_LAP(tm, slot) (tm).Lap(slot)
#else
struct SectionTimer {
  __device__ __forceinline__ void Start(bool, int) {}
};
#define T41RX_LAP(tm, slot) ((void)0)
#endif

struct F2 { float x, y; };
__device__ __forceinline__ F2 CMul(F2 a, F2 b) { return F2{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

__device__ __forceinline__ void CpAsync16(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void CpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

/* everything a receiver warp keeps in registers across blocks (uniform over the lanes unless noted) */
struct RxRegs {
  /* configuration */
  int mode, agc_mode, mirrored;
  float in_gain;         /* rfGainValue * b0 * 1.1 (DC-block numerator and freqAdjFactor folded) */
  float neg_iq_amp, iq_phase, vol_scale, volume, fixed_gain;
  /* state */
  float dc_w;            /* DC-block recurrence value after the last Q sample (B6: feeds the next block's I) */
  int rf_gain;
  unsigned codec_timer;
  int first_block;
  double ph_re, ph_im;   /* unit block phasor exp(j * phase) of the oscillator */
  int nco_closed;
  double osc_q, osc_i;   /* Osc_Vect while the amplitude loop is still settling (lane 0) */
  /* lane constants */
  F2 lane_rot;           /* nco_amp * exp(-j * delta * (16 * lane + 1)) */
  float tail_w;          /* a1^(4 * lane): weight of this lane's partial sum in the I-tail pre-read */
};

/* ------------------------------------------------------------------ */
/* radix-8 passes on the padded buffer                                  */
/* ------------------------------------------------------------------ */
template <int PASS>
__device__ __forceinline__ void FwdPass(float2 *buf, const float2 *tw, int b) {
  int n2, j, i0, stride;
  if (PASS == 0) { n2 = 64; j = b; i0 = b; stride = 1; }
  else { n2 = 8; j = b & 7; i0 = (b >> 3) * 64 + j; stride = 8; }
  float r[8], im[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float2 x = buf[FPos(i0 + m * n2)];
    r[m] = x.x;
    im[m] = x.y;
  }
  Dft8(r, im);
  buf[FPos(i0)] = float2{r[0], im[0]};
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    const float2 w = __ldg(tw + j * k * stride);
    buf[FPos(i0 + k * n2)] = float2{r[k] * w.x + im[k] * w.y, im[k] * w.x - r[k] * w.y};
  }
}

/* inverse of FwdPass<PASS> without the 1/8: conjugate twiddles on the inputs, inverse 8-point DFT.
 * kUpperHalf: store only outputs 256..511 (the valid half of the overlap-save result). */
template <int PASS, bool kUpperHalf>
__device__ __forceinline__ void InvPass(float2 *buf, const float2 *tw, int b) {
  int n2, j, i0, stride;
  if (PASS == 0) { n2 = 64; j = b; i0 = b; stride = 1; }
  else { n2 = 8; j = b & 7; i0 = (b >> 3) * 64 + j; stride = 8; }
  float r[8], im[8];
  {
    const float2 x = buf[FPos(i0)];
    r[0] = x.x;
    im[0] = x.y;
  }
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    const float2 x = buf[FPos(i0 + k * n2)];
    const float2 w = __ldg(tw + j * k * stride);
    r[k] = x.x * w.x - x.y * w.y;
    im[k] = x.y * w.x + x.x * w.y;
  }
  Dft8(im, r);   /* swapping the roles of re and im turns the forward DFT into the inverse */
#pragma unroll
  for (int m = (kUpperHalf ? 4 : 0); m < 8; ++m) buf[FPos(i0 + m * n2)] = float2{r[m], im[m]};
}

/* forward pass 2 (8 contiguous elements, no twiddles), multiply by the filter mask (bins sit in
 * octal-digit-reversed positions), inverse pass 2: all in registers */
__device__ __forceinline__ void MidPass(float2 *buf, const float2 *mask, int b) {
  float r[8], im[8];
  float2 *p = buf + 9 * b;   /* FPos(8 b) */
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float2 x = p[m];
    r[m] = x.x;
    im[m] = x.y;
  }
  Dft8(r, im);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 h = __ldg(mask + OctRev3((unsigned)(8 * b + k)));
    const float xr = r[k], xi = im[k];
    r[k] = xr * h.x - xi * h.y;
    im[k] = xr * h.y + xi * h.x;
  }
  Dft8(im, r);
#pragma unroll
  for (int m = 0; m < 8; ++m) p[m] = float2{r[m], im[m]};
}

/* ------------------------------------------------------------------ */
/* receiver warp                                                        */
/* ------------------------------------------------------------------ */
struct RxWarp {
  const LaunchArgs &a;
  float *s;          /* this receiver's slot */
  int sid;           /* receiver index */
  int lane;
  RxRegs r;
  SectionTimer tm;
  float tap1[kDec1Taps];   /* dec1 taps live in registers for the whole launch */

  __device__ __forceinline__ RxWarp(const LaunchArgs &a_, float *s_, int sid_, int lane_) : a(a_), s(s_), sid(sid_), lane(lane_) {}

  __device__ __forceinline__ const float *BlockIq(int t) const {
    return a.iq + ((size_t)sid * a.n_blocks + t) * (2 * kBlock);
  }

  /* issue the asynchronous copy of quarter q of block t into raw buffer (q & 1) */
  __device__ __forceinline__ void IssueQuarter(int t, int q) {
    const char *src = reinterpret_cast<const char *>(BlockIq(t)) + q * 4096;
    char *dst = reinterpret_cast<char *>(s + oRaw + (q & 1) * kRawBufWords);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = 32 * i + lane;              /* 16-byte piece: coalesced 512 B per instruction */
      CpAsync16(dst + (p >> 3) * (kRawChunkWords * 4) + (p & 7) * 16, src + 16 * p);
    }
    CpAsyncCommit();
  }

  /* ---------------- launch prologue ---------------- */
  __device__ void LoadState() {
    const StreamState &st = a.st[sid];
    const StreamCfg &cf = a.cfg[sid];
    const FilterSet &fs = a.fsets[cf.filter_id];
    r.mode = cf.mode;
    r.agc_mode = cf.agc_mode;
    r.mirrored = cf.mirrored;
    r.in_gain = cf.rf_gain_value * kDcB0 * 1.1f;
    r.neg_iq_amp = cf.neg_iq_amp;
    r.iq_phase = cf.iq_phase;
    r.vol_scale = cf.vol_scale;
    r.volume = cf.volume;
    r.fixed_gain = cf.agc.fixed_gain;
    r.dc_w = st.dc_d1 / (kDcB0 * (kDcA1 - 1.0f) * cf.rf_gain_value);
    r.rf_gain = st.rf_gain;
    r.codec_timer = st.codec_timer;
    r.first_block = st.first_block;
    r.nco_closed = (st.nco_closed && st.nco_epoch_seen == cf.nco_epoch) ? 1 : 0;
    {
      double sn, cs;
      sincos(st.nco_phase, &sn, &cs);
      r.ph_re = cs;
      r.ph_im = sn;
      if (st.nco_closed) {       /* leaving closed form (retune): rebuild the vector at the settled radius */
        const double rr = sqrt(st.osc_q * st.osc_q + st.osc_i * st.osc_i);
        r.osc_q = rr * cs;
        r.osc_i = rr * sn;
      } else {
        r.osc_q = st.osc_q;
        r.osc_i = st.osc_i;
      }
    }
    {
      double sn, cs;
      sincos(-cf.nco_delta * (double)(16 * lane + 1), &sn, &cs);
      r.lane_rot = F2{(float)(cf.nco_amp * cs), (float)(cf.nco_amp * sn)};
    }
    r.tail_w = PowA1(4 * lane);
    /* taps */
    for (int i = lane; i < 154; i += 32) {
      float v;
      if (i < 28) v = fs.dec1[i];
      else if (i < 74) v = fs.dec2[i - 28];
      else if (i < 122) v = fs.int1[i - 74];
      else v = fs.int2[i - 122];
      s[oTapsF + i] = v;
    }
    /* oscillator tables: W[j] = j^j * exp(-j delta j) (Fs/4 shift folded), Q[q] = exp(-j delta 512 q) */
    if (lane < 20) {
      const int n = lane < 16 ? lane : 512 * (lane - 16);
      double sn, cs;
      sincos(-cf.nco_delta * (double)n, &sn, &cs);
      float wr = (float)cs, wi = (float)sn;
      if (lane < 16) {
        const int k = lane & 3;              /* multiply by j^k */
        const float a0 = wr, b0 = wi;
        if (k == 1) { wr = -b0; wi = a0; }
        else if (k == 2) { wr = -a0; wi = -b0; }
        else if (k == 3) { wr = b0; wi = -a0; }
      }
      s[oNcoW + 2 * lane] = wr;
      s[oNcoW + 2 * lane + 1] = wi;
    }
    /* histories */
    for (int i = lane; i < 64; i += 32) {       /* dec1: plane (ch, p) entry e = -8..-1 holds sample 4 e + p */
      const int pl = i >> 3, e = (i & 7) - 8;
      const int ch = pl >> 2, p = pl & 3;
      const int n = 4 * e + p;                  /* -32 .. -1 */
      s[oMH + i] = (n >= -(kDec1Taps - 1)) ? st.dec1_hist[ch][n + (kDec1Taps - 1)] : 0.0f;
    }
    for (int i = lane; i < 96; i += 32) {       /* dec2: plane (ch, par) entry e = -24..-1 holds sample 2 e + par */
      const int pl = i / 24, e = (i % 24) - 24;
      const int ch = pl >> 1, par = pl & 1;
      const int n = 2 * e + par;
      s[oDH + i] = (n >= -(kDec2Taps - 1)) ? st.dec2_hist[ch][n + (kDec2Taps - 1)] : 0.0f;
    }
    for (int i = lane; i < 256; i += 32) {
      s[oOlaF + 2 * i] = st.ola_prev[0][i];
      s[oOlaF + 2 * i + 1] = st.ola_prev[1][i];
    }
    for (int i = lane; i < kAgcDelay; i += 32) {     /* sample -(97 - i) sits at ring index 31 + i */
      s[oZH + 2 * i] = st.agc_re[31 + i];
      s[oZH + 2 * i + 1] = st.agc_im[31 + i];
      s[oAH + i] = st.agc_abs[31 + i];
    }
    if (lane < 23) s[oIH + lane] = st.int1_hist[lane];
    if (lane < 7) s[oIH + 24 + lane] = st.int2_hist[lane];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kDec1Taps; ++i) tap1[i] = s[oTapsF + i];
  }

  __device__ void StoreState() {
    StreamState &st = a.st[sid];
    const StreamCfg &cf = a.cfg[sid];
    __syncwarp();
    for (int i = lane; i < 2 * (kDec1Taps - 1); i += 32) {
      const int ch = i / (kDec1Taps - 1), n = (i % (kDec1Taps - 1)) - (kDec1Taps - 1);   /* -27..-1 */
      const int p = n & 3, e = (n - p) / 4;                                              /* e = -7..-1 */
      st.dec1_hist[ch][n + (kDec1Taps - 1)] = s[oMH + (ch * 4 + p) * 8 + (e + 8)];
    }
    for (int i = lane; i < 2 * (kDec2Taps - 1); i += 32) {
      const int ch = i / (kDec2Taps - 1), n = (i % (kDec2Taps - 1)) - (kDec2Taps - 1);   /* -45..-1 */
      const int par = n & 1, e = (n - par) / 2;                                          /* e = -23..-1 */
      st.dec2_hist[ch][n + (kDec2Taps - 1)] = s[oDH + (ch * 2 + par) * 24 + (e + 24)];
    }
    for (int i = lane; i < 256; i += 32) {
      st.ola_prev[0][i] = s[oOlaF + 2 * i];
      st.ola_prev[1][i] = s[oOlaF + 2 * i + 1];
    }
    for (int i = lane; i < kAgcDelay; i += 32) {
      st.agc_re[31 + i] = s[oZH + 2 * i];
      st.agc_im[31 + i] = s[oZH + 2 * i + 1];
      st.agc_abs[31 + i] = s[oAH + i];
    }
    if (lane < 23) st.int1_hist[lane] = s[oIH + lane];
    if (lane < 7) st.int2_hist[lane] = s[oIH + 24 + lane];
    if (lane == 0) {
      st.dc_d1 = r.dc_w * (kDcB0 * (kDcA1 - 1.0f) * cf.rf_gain_value);
      st.dc_d2 = 0.0f;
      st.rf_gain = r.rf_gain;
      st.codec_timer = r.codec_timer;
      st.first_block = r.first_block;
      if (r.nco_closed) {
        double ph = atan2(r.ph_im, r.ph_re);
        if (ph < 0) ph += 6.283185307179586476925286766559;
        st.nco_phase = ph;
        st.nco_closed = 1;
        st.nco_epoch_seen = cf.nco_epoch;
        /* keep (osc_q, osc_i) at the settled radius so that a later exact block restarts correctly */
        const double rr = sqrt(cf.nco_r2_fix);
        st.osc_q = rr * r.ph_re;
        st.osc_i = rr * r.ph_im;
      } else {
        st.osc_q = r.osc_q;
        st.osc_i = r.osc_i;
        st.nco_closed = 0;
        st.nco_epoch_seen = cf.nco_epoch;
      }
    }
  }

  /* ---------------- front end ---------------- */
  /* blocked inclusive scan over the lanes of chunk-end values of the recurrence w <- a1 w + x
     (16 samples per lane): after it, lane L holds the true w at the end of its chunk */
  __device__ __forceinline__ float ScanDc(float e) const {
    const float m1 = PowConst16(), m2 = m1 * m1, m4 = m2 * m2, m8 = m4 * m4;
    float v = e, t;
    t = __shfl_up_sync(kFull, v, 1); if (lane >= 1) v = fmaf(m1, t, v);
    t = __shfl_up_sync(kFull, v, 2); if (lane >= 2) v = fmaf(m2, t, v);
    t = __shfl_up_sync(kFull, v, 4); if (lane >= 4) v = fmaf(m4, t, v);
    t = __shfl_up_sync(kFull, v, 8); if (lane >= 8) v = fmaf(m8, t, v);
    /* a1^256 ~ 3e-18: the 16-lane step is below any float's resolution */
    return v;
  }
  static __device__ __forceinline__ float PowConst16() {
    float p = kDcA1;        /* a1^16 by four squarings (compile-time folded) */
    p *= p; p *= p; p *= p; p *= p;
    return p;
  }

  /* one 512-sample quarter: DC block + IQ correction + Fs/4 + NCO mix -> phase planes; dec1 -> d1 planes.
     cI / cQ: recurrence values entering the quarter (updated).  base: conj(block phasor) * Q[q] * gain. */
  template <bool kTable>
  __device__ __forceinline__ void Quarter(int q, float &cI, float &cQ, F2 base, const float2 *osc) {
    const float *raw = s + oRaw + (q & 1) * kRawBufWords + lane * kRawChunkWords;
    float xi[16], xq[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = *reinterpret_cast<const float4 *>(raw + 4 * k);
      xi[2 * k] = v.x; xq[2 * k] = v.y; xi[2 * k + 1] = v.z; xq[2 * k + 1] = v.w;
    }
    /* zero-state recurrences (two independent chains) */
    float wi[16], wq[16];
    {
      float ai = 0.0f, aq = 0.0f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        ai = fmaf(kDcA1, ai, xi[j]);
        aq = fmaf(kDcA1, aq, xq[j]);
        wi[j] = ai;
        wq[j] = aq;
      }
    }
    /* carries: the value entering the quarter is folded into lane 0's chunk end */
    const float p16 = PowConst16();
    float ei = wi[15], eq = wq[15];
    if (lane == 0) { ei = fmaf(p16, cI, ei); eq = fmaf(p16, cQ, eq); }
    const float si = ScanDc(ei), sq = ScanDc(eq);
    float ci = __shfl_up_sync(kFull, si, 1), cq = __shfl_up_sync(kFull, sq, 1);
    if (lane == 0) { ci = cI; cq = cQ; }
    cI = __shfl_sync(kFull, si, 31);
    cQ = __shfl_sync(kFull, sq, 31);
    /* true recurrence values, first difference (DC-block numerator 1 - z^-1) */
    float yi[16], yq[16];
    {
      float pw = kDcA1, pi_ = ci, pq_ = cq;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float ti = fmaf(pw, ci, wi[j]);
        const float tq = fmaf(pw, cq, wq[j]);
        yi[j] = ti - pi_;
        yq[j] = tq - pq_;
        pi_ = ti;
        pq_ = tq;
        pw *= kDcA1;    /* compile-time constant after unrolling */
      }
    }
    /* I *= -IQAmp, phase correction (Process.cpp:165-174, Utility.cpp:178-187) */
    if (r.mirrored) {
#pragma unroll
      for (int j = 0; j < 16; ++j) yi[j] *= r.neg_iq_amp;
      if (r.iq_phase != 0.0f) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (r.iq_phase < 0.0f) yq[j] = fmaf(yi[j], r.iq_phase, yq[j]);
          else yi[j] = fmaf(yq[j], r.iq_phase, yi[j]);
        }
      }
    }
    /* mix: multiplier of sample j = base * lane_rot * W[j] */
    const F2 m = CMul(base, r.lane_rot);
    const float4 *wt = reinterpret_cast<const float4 *>(s + oNcoW);
    float oi[16], oq[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      F2 m0, m1;
      if (kTable) {
        /* settling oscillator: multiplier = gain * conj(osc[n]) * j^n, osc from the step-by-step table */
        const float4 o2 = *reinterpret_cast<const float4 *>(osc + 16 * lane + 2 * k);
        const F2 c0 = F2{o2.x * base.x, -o2.y * base.x}, c1 = F2{o2.z * base.x, -o2.w * base.x};
        m0 = (k & 1) ? F2{-c0.x, -c0.y} : c0;                    /* sample 2k:   j^(2k)   = (-1)^k   */
        m1 = (k & 1) ? F2{c1.y, -c1.x} : F2{-c1.y, c1.x};        /* sample 2k+1: j^(2k+1) = j (-1)^k */
      } else {
        const float4 w2 = wt[k];
        m0 = CMul(m, F2{w2.x, w2.y});
        m1 = CMul(m, F2{w2.z, w2.w});
      }
      oi[2 * k] = yi[2 * k] * m0.x - yq[2 * k] * m0.y;
      oq[2 * k] = yi[2 * k] * m0.y + yq[2 * k] * m0.x;
      oi[2 * k + 1] = yi[2 * k + 1] * m1.x - yq[2 * k + 1] * m1.y;
      oq[2 * k + 1] = yi[2 * k + 1] * m1.y + yq[2 * k + 1] * m1.x;
    }
    /* phase planes: sample 16 L + j -> plane (j & 3), entry 4 L + (j >> 2) */
    float *mix = s + oMix;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      *reinterpret_cast<float4 *>(mix + p * kMixPlane + 8 + 4 * lane) = float4{oi[p], oi[4 + p], oi[8 + p], oi[12 + p]};
      *reinterpret_cast<float4 *>(mix + (4 + p) * kMixPlane + 8 + 4 * lane) = float4{oq[p], oq[4 + p], oq[8 + p], oq[12 + p]};
    }
    __syncwarp();
    T41RX_LAP(tm, 2);
    Dec1Quarter(q);
    T41RX_LAP(tm, 3);
  }

  /* arm_fir_decimate_f32, M = 4, 28 taps (Process.cpp:474-475): 4 outputs per lane per quarter.
     Output m (quarter-local) = sum_t h[t] x[4 m - 27 + t]; sample 4 m - 27 + t = plane (1 + t) & 3,
     entry m - 7 + ((1 + t) >> 2). */
  __device__ __forceinline__ void Dec1Quarter(int q) {
    const float *mix = s + oMix;
    float *d1 = s + oD1;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      float w[4][12];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float4 *src = reinterpret_cast<const float4 *>(mix + (ch * 4 + p) * kMixPlane + 4 * lane);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 v = src[k];
          w[p][4 * k] = v.x; w[p][4 * k + 1] = v.y; w[p][4 * k + 2] = v.z; w[p][4 * k + 3] = v.w;
        }
      }
#pragma unroll
      for (int t = 0; t < kDec1Taps; ++t) {
        const int p = (1 + t) & 3, off = 1 + ((1 + t) >> 2);
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = fmaf(w[p][o + off], tap1[t], acc[o]);
      }
      /* d1 sample n = 128 q + 4 L + o -> plane (n & 1), entry n >> 1 */
      const int e = 64 * q + 2 * lane;
      *reinterpret_cast<float2 *>(d1 + (ch * 2 + 0) * kD1Plane + 24 + e) = float2{acc[0], acc[2]};
      *reinterpret_cast<float2 *>(d1 + (ch * 2 + 1) * kD1Plane + 24 + e) = float2{acc[1], acc[3]};
    }
    __syncwarp();
    /* slide the plane histories: entries 120..127 become -8..-1 */
    for (int i = lane; i < 64; i += 32) {
      const int pl = i >> 3, e = i & 7;
      const float v = s[oMix + pl * kMixPlane + 8 + 120 + e];
      s[oMix + pl * kMixPlane + e] = v;
    }
    __syncwarp();
  }

  /* arm_fir_decimate_f32, M = 2, 46 taps (Process.cpp:478-479): 8 outputs per lane.
     Output o = sum_t h[t] d[2 o - 45 + t]; sample 2 o - 45 + t = plane (1 + t) & 1, entry o - 23 + ((1 + t) >> 1). */
  __device__ __forceinline__ void Dec2(float (&out)[2][8]) {
    const float *d1 = s + oD1;
    const float *tp = s + oTapsF + 28;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      float w[2][32];
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const float4 *src = reinterpret_cast<const float4 *>(d1 + (ch * 2 + par) * kD1Plane + 8 * lane);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 v = src[k];
          w[par][4 * k] = v.x; w[par][4 * k + 1] = v.y; w[par][4 * k + 2] = v.z; w[par][4 * k + 3] = v.w;
        }
      }
      float acc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = 0.0f;
#pragma unroll
      for (int t = 0; t < kDec2Taps; ++t) {
        const float h = tp[t];
        const int par = (1 + t) & 1, off = 1 + ((1 + t) >> 1);
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = fmaf(w[par][o + off], h, acc[o]);
      }
#pragma unroll
      for (int o = 0; o < 8; ++o) out[ch][o] = acc[o];
    }
  }

  /* FE for block t.  row blocks and oscillator transients are handled by the caller. */
  __device__ void FrontEnd(int t, int buf) {
    const StreamCfg &cf = a.cfg[sid];
    /* the recurrence value entering the Q chain is the one leaving the I chain (B6): pre-read the last
       128 I samples of the block (a1^128 ~ 2e-9) */
    float tail;
    {
      const float4 *p = reinterpret_cast<const float4 *>(BlockIq(t) + 2 * (kBlock - 4 * (lane + 1)));
      const float4 u = __ldg(p), v = __ldg(p + 1);    /* samples n0 .. n0+3, n0 = 2044 - 4 lane */
      float acc = u.x;                                 /* oldest first */
      acc = fmaf(kDcA1, acc, u.z);
      acc = fmaf(kDcA1, acc, v.x);
      acc = fmaf(kDcA1, acc, v.z);
      tail = acc * r.tail_w;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) tail += __shfl_xor_sync(kFull, tail, d);
    }
    /* restore the decimator histories in front of the planes */
    for (int i = lane; i < 64; i += 32) s[oMix + (i >> 3) * kMixPlane + (i & 7)] = s[oMH + i];
    for (int i = lane; i < 96; i += 32) s[oD1 + (i / 24) * kD1Plane + (i % 24)] = s[oDH + i];
    /* gains of this block (Process.cpp:117,133): rfGainValue and RFgain are folded into the phasor */
    const float gain = r.in_gain * (float)r.rf_gain;
    T41RX_LAP(tm, 0);
    const F2 pb = F2{(float)r.ph_re * gain, -(float)r.ph_im * gain};
    float cI = r.dc_w, cQ = tail;
    /* NB: tail misses a1^2048 * (state entering I), which is exactly 0 in float */
    if (r.nco_closed) {
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        CpAsyncWaitAll();
        __syncwarp();
        T41RX_LAP(tm, 1);
        if (q < 3) IssueQuarter(t, q + 1);
        else if (t + 1 < a.n_blocks) IssueQuarter(t + 1, 0);
        const F2 qq = F2{s[oNcoW + 32 + 2 * q], s[oNcoW + 32 + 2 * q + 1]};
        Quarter<false>(q, cI, cQ, CMul(pb, qq), nullptr);
      }
      /* advance the block phasor by 2048 samples */
      double sn, cs;
      sincos(cf.nco_block_delta, &sn, &cs);
      const double nr = r.ph_re * cs - r.ph_im * sn, ni = r.ph_re * sn + r.ph_im * cs;
      r.ph_re = nr;
      r.ph_im = ni;
    } else {
      /* the oscillator's amplitude loop has not settled (first block of a receiver, block after a
         retune): FreqShift2's FP64 recurrence step by step (Freq_Shift.cpp:126-140) on lane 0, one
         quarter at a time into the idle raw buffer (so no copy is in flight during such a block) */
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        CpAsyncWaitAll();
        __syncwarp();
        float2 *tab = reinterpret_cast<float2 *>(s + oRaw + ((q + 1) & 1) * kRawBufWords);
        if (lane == 0) {
          double vq = r.osc_q, vi = r.osc_i;
          const double oc = cf.osc_cos, os = cf.osc_sin;
          for (int n = 0; n < 512; ++n) {
            const double oq = (vq * oc) - (vi * os);
            const double oi = (vi * oc) + (vq * os);
            const double gn = 1.95 - ((vq * vq) + (vi * vi));
            vq = gn * oq;
            vi = gn * oi;
            tab[n] = float2{(float)oq, (float)oi};
          }
          r.osc_q = vq;
          r.osc_i = vi;
        }
        __syncwarp();
        Quarter<true>(q, cI, cQ, F2{gain, 0.0f}, tab);
        if (q < 3) IssueQuarter(t, q + 1);
        else if (t + 1 < a.n_blocks) IssueQuarter(t + 1, 0);
      }
      /* settled?  then continue in closed form from the vector's angle */
      int settled = 0;
      double pr = 0.0, pi2 = 0.0;
      if (lane == 0) {
        const double r2 = r.osc_q * r.osc_q + r.osc_i * r.osc_i;
        settled = fabs(r2 - cf.nco_r2_fix) < 4.0e-15;
        const double inv = rsqrt(r2);
        pr = r.osc_q * inv;
        pi2 = r.osc_i * inv;
      }
      settled = __shfl_sync(kFull, settled, 0);
      if (settled) {
        r.nco_closed = 1;
        r.ph_re = __shfl_sync(kFull, pr, 0);
        r.ph_im = __shfl_sync(kFull, pi2, 0);
      }
    }
    r.dc_w = cQ;
    /* save the dec1 plane histories (the FFT buffer overlays the planes) */
    for (int i = lane; i < 64; i += 32) s[oMH + i] = s[oMix + (i >> 3) * kMixPlane + (i & 7)];
    __syncwarp();
    float dq[2][8];
    T41RX_LAP(tm, 4);
    Dec2(dq);
    __syncwarp();
    T41RX_LAP(tm, 5);
    /* dec2 history for the next block: entries 232..255 of each plane */
    for (int i = lane; i < 96; i += 32) s[oDH + i] = s[oD1 + (i / 24) * kD1Plane + 24 + 232 + (i % 24)];
    __syncwarp();
    AfterDec2(dq, buf);
    /* Codec_gain (Process.cpp:979-1016 with the clip flags never set) */
    {
      unsigned timer = r.codec_timer + 1;
      if (timer > 10000) timer = 10000;
      if (timer >= 50) {
        r.rf_gain = min(r.rf_gain + 1, 15);
        timer = 0;
      }
      r.codec_timer = timer;
    }
  }

  /* level adjust + overlap-save + fast convolution + |z| + window maximum -> staging */
  __device__ void AfterDec2(float (&dq)[2][8], int buf) {
    float2 *fb = reinterpret_cast<float2 *>(s + oMix);
    float2 *ola = reinterpret_cast<float2 *>(s + oOlaF);
    float2 *stz = reinterpret_cast<float2 *>(s + oStZ + buf * 512);
    const int o0 = 8 * lane;
    if (r.mode == kModePsk31) {           /* Process.cpp:376-387,745: raw decimated I, no filter, no AGC */
#pragma unroll
      for (int o = 0; o < 8; ++o) stz[o0 + o] = float2{dq[0][o], 0.0f};
      return;
    }
    if (r.mode == kModeNfm) {
      NfmDiscriminator(dq, fb, ola);
    } else {
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const float2 cur = float2{dq[0][o] * r.vol_scale, dq[1][o] * r.vol_scale};   /* Process.cpp:482-492 */
        float2 prev = ola[o0 + o];
        if (r.first_block) prev = float2{0.0f, 0.0f};                                /* Process.cpp:498-504 */
        fb[FPos(o0 + o)] = prev;
        fb[FPos(256 + o0 + o)] = cur;
        ola[o0 + o] = cur;
      }
      r.first_block = 0;
    }
    __syncwarp();
    const float2 *tw = a.twiddle;
    const float2 *mask = reinterpret_cast<const float2 *>(a.fsets[a.cfg[sid].filter_id].mask);
    FwdPass<0>(fb, tw, lane); FwdPass<0>(fb, tw, lane + 32);
    __syncwarp();
    FwdPass<1>(fb, tw, lane); FwdPass<1>(fb, tw, lane + 32);
    __syncwarp();
    MidPass(fb, mask, lane); MidPass(fb, mask, lane + 32);
    __syncwarp();
    InvPass<1, false>(fb, tw, lane); InvPass<1, false>(fb, tw, lane + 32);
    __syncwarp();
    InvPass<0, true>(fb, tw, lane); InvPass<0, true>(fb, tw, lane + 32);
    __syncwarp();
    T41RX_LAP(tm, 6);
    /* valid outputs 256..511, scaled by 1/512 */
    float2 z[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float2 v = fb[FPos(256 + o0 + o)];
      z[o] = float2{v.x * (1.0f / 512.0f), v.y * (1.0f / 512.0f)};
    }
    if (r.agc_mode == 0) {                /* DSP_Fn.cpp:494-502: fixed gain, no delay line */
#pragma unroll
      for (int o = 0; o < 8; ++o) stz[o0 + o] = z[o];
      return;
    }
    /* delayed output: zd[i] = z[i - 97]; keep the last 97 for the next block */
    float2 *zh = reinterpret_cast<float2 *>(s + oZH);
    float *E = s + vE;
    for (int i = lane; i < kAgcDelay; i += 32) {
      stz[i] = zh[i];
      E[i] = s[oAH + i];
    }
    __syncwarp();
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int i = o0 + o;
      if (i + kAgcDelay < kDec) stz[i + kAgcDelay] = z[o];
      else zh[i + kAgcDelay - kDec] = z[o];
      E[kAgcDelay + i] = __fsqrt_rn(z[o].x * z[o].x + z[o].y * z[o].y);
    }
    if (lane < 7) E[353 + lane] = 0.0f;
    __syncwarp();
    /* chunk pass: prefix / suffix maxima inside chunks of 8 (NaN magnitudes count as 0) */
    for (int c = lane; c < 45; c += 32) {
      float v[8];
      const float4 u0 = *reinterpret_cast<const float4 *>(E + 8 * c), u1 = *reinterpret_cast<const float4 *>(E + 8 * c + 4);
      v[0] = u0.x; v[1] = u0.y; v[2] = u0.z; v[3] = u0.w; v[4] = u1.x; v[5] = u1.y; v[6] = u1.z; v[7] = u1.w;
      float p[8], sf[8];
      p[0] = fmaxf(v[0], 0.0f);
#pragma unroll
      for (int k = 1; k < 8; ++k) p[k] = fmaxf(p[k - 1], v[k]);
      sf[7] = fmaxf(v[7], 0.0f);
#pragma unroll
      for (int k = 6; k >= 0; --k) sf[k] = fmaxf(sf[k + 1], v[k]);
      *reinterpret_cast<float4 *>(s + vPfx + 8 * c) = float4{p[0], p[1], p[2], p[3]};
      *reinterpret_cast<float4 *>(s + vPfx + 8 * c + 4) = float4{p[4], p[5], p[6], p[7]};
      *reinterpret_cast<float4 *>(s + vSfx + 8 * c) = float4{sf[0], sf[1], sf[2], sf[3]};
      *reinterpret_cast<float4 *>(s + vSfx + 8 * c + 4) = float4{sf[4], sf[5], sf[6], sf[7]};
      s[vCM + c] = p[7];
    }
    __syncwarp();
    /* F[c] = max(CM[c .. c+10]) for c = 1 .. 33 (kept in registers: lane L needs F[L+1] and F[L+2]) */
    float f1 = 0.0f, f2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      f1 = fmaxf(f1, s[vCM + lane + 1 + k]);
      f2 = fmaxf(f2, s[vCM + min(lane + 2 + k, 44)]);
    }
    /* rm[i] = max(E[i+1 .. i+97]) = max(Sfx[i+1], F[((i+1) >> 3) + 1], Pfx[i+97]); |z| delayed = E[i] */
    float *sta = s + oStA + buf * 512;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int i = o0 + o;
      const float f = (o == 7) ? f2 : f1;
      sta[256 + i] = fmaxf(fmaxf(s[vSfx + i + 1], f), s[vPfx + i + 97]);
      sta[i] = E[i];
    }
    __syncwarp();
    for (int i = lane; i < kAgcDelay; i += 32) s[oAH + i] = E[kDec + i];
    T41RX_LAP(tm, 7);
  }

  /* NFM: discriminator on the decimated samples, then the audio goes through the filter as a real
     signal (Demod.cpp:220-235, Process.cpp:716-727,765-779) */
  __device__ __forceinline__ void NfmDiscriminator(float (&dq)[2][8], float2 *fb, float2 *ola) {
    StreamState &st = a.st[sid];
    const float kq = 0.340447550238101026565118445432744920253753662109375f;
    const int o0 = 8 * lane;
    /* previous sample of the lane's first output comes from the lane below */
    float pi_ = __shfl_up_sync(kFull, dq[0][7], 1), pq_ = __shfl_up_sync(kFull, dq[1][7], 1);
    const float li = st.nfm_last_i, lq = st.nfm_last_q;
    float outv[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float I = dq[0][o], Q = dq[1][o];
      const float den = I * I + Q * Q;
      float out;
      if (o0 + o == 0) {
        const float num = I * (Q - lq) - Q * (I - li);
        out = kq * num / den;
      } else {
        const float num = Q * pi_ - I * pq_;
        out = kq * num / den;
        out = (1.0f < out) ? 1.0f : out;          /* limiter skips index 0 (B5) */
        out = (-1.0f > out) ? -1.0f : out;
      }
      outv[o] = out;
      pi_ = I;
      pq_ = Q;
    }
    __syncwarp();
    if (lane == 15) {                              /* "last sample" = complex sample 127 (B4) */
      st.nfm_last_i = dq[0][7];
      st.nfm_last_q = dq[1][7];
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      fb[FPos(o0 + o)] = float2{ola[o0 + o].x, 0.0f};
      fb[FPos(256 + o0 + o)] = float2{outv[o], 0.0f};
      ola[o0 + o].x = outv[o];
    }
  }

  /* ---------------- back end ---------------- */
  __device__ void BackEnd(int t, int buf) {
    const StreamCfg &cf = a.cfg[sid];
    StreamState &st = a.st[sid];
    const float2 *stz = reinterpret_cast<const float2 *>(s + oStZ + buf * 512);
    const float *volts = s + oStA + buf * 512 + 256;
    float *aud = s + vAudF;
    const int o0 = 8 * lane;
    float2 dem[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) dem[o] = stz[o0 + o];
    if (r.mode != kModePsk31) {
      if (r.agc_mode == 0) {
#pragma unroll
        for (int o = 0; o < 8; ++o) dem[o] = float2{dem[o].x * r.fixed_gain, dem[o].y * r.fixed_gain};
      } else {
        const AgcConsts &ag = cf.agc;
        const float inv_in = ag.inv_max_input, tgt = ag.out_target, slope = ag.slope_constant;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const float v = volts[o0 + o];
          const float lg = Log10Fast(inv_in * v);             /* DSP_Fn.cpp:628 */
          const float clipped = (0.0f < lg) ? 0.0f : lg;
          const float mult = (tgt - slope * clipped) / v;
          dem[o] = float2{dem[o].x * mult, dem[o].y * mult};
        }
      }
    }
    /* demodulators (Process.cpp:615-761) */
    float au[8];
    if (r.mode == kModeAm) {
      AmDetect(dem, au, st);
    } else if (r.mode == kModeSam) {
      SamDetect(dem, au, st);
    } else {
#pragma unroll
      for (int o = 0; o < 8; ++o) au[o] = dem[o].x;           /* USB / LSB / NFM / PSK31: real part */
    }
    /* PSK31 tap (psk31.cpp:235-310): first filtered sample of every third block */
    if (a.psk_bits || a.psk_chars || cf.psk31_enable) PskTap(t, dem[0], cf, st);
    /* int1 input: 1 pad + 23 history + 256 */
    if (lane < 23) aud[1 + lane] = s[oIH + lane];
    *reinterpret_cast<float4 *>(aud + 24 + o0) = float4{au[0], au[1], au[2], au[3]};
    *reinterpret_cast<float4 *>(aud + 24 + o0 + 4) = float4{au[4], au[5], au[6], au[7]};
    __syncwarp();
    T41RX_LAP(tm, 8);
    Interp1();
    __syncwarp();
    T41RX_LAP(tm, 9);
    if (lane < 23) s[oIH + lane] = aud[24 + 233 + lane];
    Interp2(t);
    __syncwarp();
    if (lane < 7) s[oIH + 24 + lane] = s[vI1 + 8 + 505 + lane];
    __syncwarp();
    T41RX_LAP(tm, 10);
  }

  /* arm_fir_interpolate_f32, L = 2, 48 taps (Process.cpp:917): 8 inputs per lane */
  __device__ __forceinline__ void Interp1() {
    const float *aud = s + vAudF;
    const float *tp = s + oTapsF + 74;
    float w[32];
    const float4 *src = reinterpret_cast<const float4 *>(aud + 8 * lane);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = src[k];
      w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
    }
    float a0[8], a1[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) { a0[o] = 0.0f; a1[o] = 0.0f; }
#pragma unroll
    for (int k = 0; k < 24; ++k) {
      const float c1 = tp[2 * k + 1], c0 = tp[2 * k];
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const float x = w[o + 1 + k];          /* input n - 23 + k, n = 8 L + o, lives at word 8 L + o + 1 + k */
        a0[o] = fmaf(x, c1, a0[o]);            /* phase 0: c[(L-1) + k L] */
        a1[o] = fmaf(x, c0, a1[o]);            /* phase 1: c[0 + k L]     */
      }
    }
    float *i1 = s + vI1 + 8 + 16 * lane;
#pragma unroll
    for (int o = 0; o < 8; o += 2)
      *reinterpret_cast<float4 *>(i1 + 2 * o) = float4{a0[o], a1[o], a0[o + 1], a1[o + 1]};
    if (lane < 7) s[vI1 + 1 + lane] = s[oIH + 24 + lane];
  }

  /* arm_fir_interpolate_f32, L = 4, 32 taps + volume (Process.cpp:919-931): 4 x 4 inputs per lane */
  __device__ __forceinline__ void Interp2(int t) {
    const float *tp = s + oTapsF + 122;
    float c[kInt2Taps];
#pragma unroll
    for (int i = 0; i < kInt2Taps; ++i) c[i] = tp[i];
    float4 *dst = reinterpret_cast<float4 *>(a.audio + ((size_t)sid * a.n_blocks + t) * kBlock);
    const float vol = r.volume;
#pragma unroll 1
    for (int rr = 0; rr < 4; ++rr) {
      const int n0 = 4 * lane + 128 * rr;
      float w[12];
      const float4 *src = reinterpret_cast<const float4 *>(s + vI1 + n0);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float4 v = src[k];
        w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float x = w[o + 1 + k];        /* input n - 7 + k at word n + 1 + k */
#pragma unroll
          for (int p = 0; p < 4; ++p) acc[p] = fmaf(x, c[4 * k + (3 - p)], acc[p]);
        }
        dst[n0 + o] = float4{acc[0] * vol, acc[1] * vol, acc[2] * vol, acc[3] * vol};
      }
    }
  }

  /* AM: alpha-beta magnitude, 1-pole DC removal, 1-stage DF1 low-pass (Process.cpp:697-707) as blocked
     linear scans over the lanes (8 samples per lane) */
  __device__ __forceinline__ void AmDetect(const float2 (&dem)[8], float (&au)[8], StreamState &st) {
    const StreamCfg &cf = a.cfg[sid];
    float m[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) m[o] = AlphaBetaMag(dem[o].x, dem[o].y);
    /* w[i] = m[i] + 0.99 w[i-1] */
    const float g = 0.99f;
    float w[8];
    {
      float acc = 0.0f;
#pragma unroll
      for (int o = 0; o < 8; ++o) { acc = fmaf(g, acc, m[o]); w[o] = acc; }
    }
    const float wold = st.am_wold;
    float g8 = g * g; g8 *= g8; g8 *= g8;       /* 0.99^8 */
    float e = w[7];
    if (lane == 0) e = fmaf(g8, wold, e);
    {
      float mult = g8, v = e, tt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        tt = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v = fmaf(mult, tt, v);
        mult *= mult;
      }
      e = v;
    }
    float cin = __shfl_up_sync(kFull, e, 1);
    if (lane == 0) cin = wold;
    const float wend = __shfl_sync(kFull, e, 31);
    float x[8];
    {
      float pw = g, prev = cin;
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const float tw_ = fmaf(pw, cin, w[o]);
        x[o] = tw_ - prev;
        prev = tw_;
        pw *= g;
      }
    }
    /* biquad: y[i] = b0 x[i] + b1 x[i-1] + b2 x[i-2] + a1 y[i-1] + a2 y[i-2] */
    const float b0 = cf.am_lp[0], b1 = cf.am_lp[1], b2 = cf.am_lp[2], a1 = cf.am_lp[3], a2 = cf.am_lp[4];
    float xm1 = __shfl_up_sync(kFull, x[7], 1), xm2 = __shfl_up_sync(kFull, x[6], 1);
    if (lane == 0) { xm1 = st.am_lp_state[0]; xm2 = st.am_lp_state[1]; }
    float y[8];
    {
      float p1 = 0.0f, p2 = 0.0f, q1 = xm1, q2 = xm2;
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        float f = b0 * x[o];
        f = fmaf(b1, q1, f);
        f = fmaf(b2, q2, f);
        const float yy = fmaf(a1, p1, fmaf(a2, p2, f));
        y[o] = yy;
        p2 = p1; p1 = yy;
        q2 = q1; q1 = x[o];
      }
    }
    /* homogeneous responses h1 (y[-1] = 1, y[-2] = 0) and h2 (y[-1] = 0, y[-2] = 1) over 8 steps */
    float h1[8], h2[8];
    {
      float u1 = 1.0f, u2 = 0.0f, v1 = 0.0f, v2 = 1.0f;
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const float un = a1 * u1 + a2 * u2, vn = a1 * v1 + a2 * v2;
        h1[o] = un; h2[o] = vn;
        u2 = u1; u1 = un;
        v2 = v1; v1 = vn;
      }
    }
    /* chunk transfer matrix M = [[h1[7], h2[7]], [h1[6], h2[6]]]; scan of (y[7], y[6]) */
    float m00 = h1[7], m01 = h2[7], m10 = h1[6], m11 = h2[6];
    float s1 = y[7], s2 = y[6];
    if (lane == 0) {
      const float y1 = st.am_lp_state[2], y2 = st.am_lp_state[3];
      s1 += m00 * y1 + m01 * y2;
      s2 += m10 * y1 + m11 * y2;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float t1 = __shfl_up_sync(kFull, s1, d), t2 = __shfl_up_sync(kFull, s2, d);
      if (lane >= d) {
        s1 += m00 * t1 + m01 * t2;
        s2 += m10 * t1 + m11 * t2;
      }
      const float n00 = m00 * m00 + m01 * m10, n01 = m00 * m01 + m01 * m11;
      const float n10 = m10 * m00 + m11 * m10, n11 = m10 * m01 + m11 * m11;
      m00 = n00; m01 = n01; m10 = n10; m11 = n11;
    }
    float c1 = __shfl_up_sync(kFull, s1, 1), c2 = __shfl_up_sync(kFull, s2, 1);
    if (lane == 0) { c1 = st.am_lp_state[2]; c2 = st.am_lp_state[3]; }
#pragma unroll
    for (int o = 0; o < 8; ++o) au[o] = y[o] + h1[o] * c1 + h2[o] * c2;
    __syncwarp();
    if (lane == 31) {
      st.am_wold = wend;
      st.am_lp_state[0] = x[7];
      st.am_lp_state[1] = x[6];
      st.am_lp_state[2] = au[7];
      st.am_lp_state[3] = au[6];
    }
  }

  /* SAM PLL (Demod.cpp:40-139): serial, lane 0, through shared memory */
  __device__ void SamDetect(const float2 (&dem)[8], float (&au)[8], StreamState &st) {
    float2 *tmp = reinterpret_cast<float2 *>(s + vI1);     /* 256 complex in, audio written over .x */
#pragma unroll
    for (int o = 0; o < 8; ++o) tmp[8 * lane + o] = dem[o];
    __syncwarp();
    if (lane == 0) {
      const float tpi = 6.283185307179586476925286766559f;
      const float omega_min = __ldg(a.sam_consts + 0), omega_max = __ldg(a.sam_consts + 1);
      const float g1 = __ldg(a.sam_consts + 2), g2 = __ldg(a.sam_consts + 3);
      float phz = st.sam_phzerror, fil = st.sam_fil_out, om2 = st.sam_omega2;
      for (int i = 0; i < kDec; ++i) {
        const float2 z = tmp[i];
        const float sn = TableTurns(a.sin_table, phz * 0.159154943092f);
        const float cs = TableTurns(a.sin_table, phz * 0.159154943092f + 0.25f);
        const float ai = cs * z.x, bi = sn * z.x, aq = cs * z.y, bq = sn * z.y;
        const float corr0 = +ai + bq;
        const float corr1 = -bi + aq;
        tmp[i].x = (ai - bi) + (aq + bq);
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
    __syncwarp();
#pragma unroll
    for (int o = 0; o < 8; ++o) au[o] = tmp[8 * lane + o].x;
    __syncwarp();
  }

  __device__ void PskTap(int t, float2 d0, const StreamCfg &cf, StreamState &st) {
    if (lane != 0) return;
    int8_t bit_out = -1;
    uint8_t char_out = 0;
    if (cf.psk31_enable && r.mode != kModeNfm && r.mode != kModePsk31) {
      if (st.psk_block_count % 3u == 0u) {
        const double pi_d = 3.1415926535897932384626433832795;
        const float phase = Atan2Approx(d0.y, d0.x);
        float dphase = phase - st.psk_last_phase;
        while ((double)dphase < -pi_d) dphase = (float)((double)dphase + 2 * pi_d);
        while ((double)dphase >= pi_d) dphase = (float)((double)dphase - 2 * pi_d);
        const uint8_t bit = (((double)dphase > (pi_d / 2)) || ((double)dphase < (-pi_d / 2))) ? 0 : 1;
        st.psk_last_phase = phase;
        bit_out = (int8_t)bit;
        unsigned long long shr = (st.psk_shr << 1) | (unsigned long long)bit;
        if ((shr & 0xFFFull) != 0) {
          for (int i = 0; i < 128; ++i) {
            const uint32_t e = __ldg(a.varicode + i);
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
    const size_t o = (size_t)sid * a.n_blocks + t;
    if (a.psk_bits) a.psk_bits[o] = bit_out;
    if (a.psk_chars) a.psk_chars[o] = char_out;
  }
};

/* ------------------------------------------------------------------ */
/* AGC warp: lane = receiver (DSP_Fn.cpp:504-631)                       */
/* ------------------------------------------------------------------ */
struct AgcLane {
  /* constants */
  float fbm, omfbm, hbm, omhbm, attack, decay, fdecay, hdecay, pop, hlevel, minv;
  int hload, henable;
  /* state */
  float fast, hang, v, save, rm;
  int hc, state, dtype, action;

  __device__ void Load(const StreamCfg &cf, const StreamState &st) {
    fbm = cf.agc.fast_backmult; omfbm = cf.agc.onemfast_backmult;
    hbm = cf.agc.hang_backmult; omhbm = cf.agc.onemhang_backmult;
    attack = cf.agc.attack_mult; decay = cf.agc.decay_mult; fdecay = cf.agc.fast_decay_mult;
    hdecay = cf.agc.hang_decay_mult; pop = cf.agc.pop_ratio; hlevel = cf.agc.hang_level;
    minv = cf.agc.min_volts;
    hload = cf.agc.hang_counter_load; henable = cf.agc.hang_enable;
    fast = st.agc_fast_back; hang = st.agc_hang_back; v = st.agc_volts; save = st.agc_save_volts;
    rm = st.agc_ring_max;
    hc = st.agc_hang_counter; state = st.agc_state; dtype = st.agc_decay_type; action = st.agc_action;
  }
  __device__ void Store(StreamState &st) const {
    st.agc_fast_back = fast; st.agc_hang_back = hang; st.agc_volts = v; st.agc_save_volts = save;
    st.agc_ring_max = rm;
    st.agc_hang_counter = hc; st.agc_state = state; st.agc_decay_type = dtype; st.agc_action = action;
  }

  /* one sample of the envelope state machine; returns volts */
  __device__ __forceinline__ float Step(float abs_out, float r) {
    fast = fbm * abs_out + omfbm * fast;
    hang = hbm * abs_out + omhbm * hang;
    rm = r;
    if (hc > 0) --hc;
    const float d = r - v;
    if (r >= v) {
      if (state >= 2) save = v;
      state = 0;
      v += d * attack;
    } else if (state == 3) {
      v = fmaf(d * decay, 0.05f, v);      /* reference: double product and sum, then float */
    } else if (state == 0) {
      if (v > pop * fast) {
        state = 1;
        v += d * fdecay;
      } else if (henable && (hang > hlevel)) {
        state = 2;
        hc = hload;
        dtype = 1;
      } else {
        state = 3;
        v += d * decay;
        dtype = 0;
      }
    } else if (state == 1) {
      if (v > save) {
        v += d * fdecay;
      } else if (hc > 0) {
        state = 2;
      } else if (dtype == 0) {
        state = 3;
        v += d * decay;
      } else {
        state = 4;
        v += d * hdecay;
      }
    } else if (state == 2) {
      if (hc == 0) {
        state = 4;
        v += d * hdecay;
      }
    } else {
      v += d * hdecay;
    }
    action = (v < minv) ? 0 : 1;
    v = (v < minv) ? minv : v;
    return v;
  }
};

/* one block of one receiver per lane.  sta: |z| delayed [256] then window max [256] (overwritten by volts).
 *
 * Almost every sample leaves the envelope detector in a "quiet" state: slow decay (3), hang decay (4) or
 * hang (2) with the window maximum below volts, where the update is  v <- max(v + ((r - v) k1) k2, min_volts)
 * with per-state constants and no branch.  The warp therefore runs 8 samples at a time speculatively on that
 * branch-free form (all lanes), votes once, and only re-runs the chunk through the exact per-sample state
 * machine (Step) when some lane saw a transition (attack, end of hang, states 0 / 1). */
__device__ __forceinline__ void AgcBlock(AgcLane &g, float *sta, bool active) {
  constexpr int kC = 8;
#pragma unroll 1
  for (int i0 = 0; i0 < kDec; i0 += kC) {
    float ab[kC], rm[kC];
    if (active) {
#pragma unroll
      for (int k = 0; k < kC; k += 4) {
        const float4 a4 = *reinterpret_cast<const float4 *>(sta + i0 + k);
        const float4 r4 = *reinterpret_cast<const float4 *>(sta + 256 + i0 + k);
        ab[k] = a4.x; ab[k + 1] = a4.y; ab[k + 2] = a4.z; ab[k + 3] = a4.w;
        rm[k] = r4.x; rm[k + 1] = r4.y; rm[k + 2] = r4.z; rm[k + 3] = r4.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < kC; ++k) { ab[k] = 0.0f; rm[k] = 0.0f; }
    }
    /* speculative quiet path */
    float k1 = 0.0f, k2 = 0.0f;
    bool ok = true;
    if (g.state == 3) { k1 = g.decay; k2 = 0.05f; }
    else if (g.state == 4) { k1 = g.hdecay; k2 = 1.0f; }
    else if (g.state == 2 && g.hc > kC) { k1 = 0.0f; k2 = 0.0f; }
    else ok = false;
    float v = g.v, fast = g.fast, hang = g.hang, last = g.v;
    float vo[kC];
#pragma unroll
    for (int k = 0; k < kC; ++k) {
      ok = ok && !(rm[k] >= v);
      fast = g.fbm * ab[k] + g.omfbm * fast;
      hang = g.hbm * ab[k] + g.omhbm * hang;
      const float t = (rm[k] - v) * k1;
      last = fmaf(t, k2, v);
      v = fmaxf(last, g.minv);
      vo[k] = v;
    }
    if (__all_sync(kFull, ok || !active)) {
      if (active) {
        g.v = v; g.fast = fast; g.hang = hang;
        g.rm = rm[kC - 1];
        g.hc = max(g.hc - kC, 0);
        g.action = (last < g.minv) ? 0 : 1;
      }
    } else if (active) {
#pragma unroll 1
      for (int k = 0; k < kC; ++k) {
        /* select ab[k], rm[k] without dynamic register indexing */
        float a_ = ab[0], r_ = rm[0];
#pragma unroll
        for (int j = 1; j < kC; ++j) { if (k == j) { a_ = ab[j]; r_ = rm[j]; } }
        const float vv = g.Step(a_, r_);
#pragma unroll
        for (int j = 0; j < kC; ++j) { if (k == j) vo[j] = vv; }
      }
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < kC; k += 4)
        *reinterpret_cast<float4 *>(sta + 256 + i0 + k) = float4{vo[k], vo[k + 1], vo[k + 2], vo[k + 3]};
    }
  }
}

/* ------------------------------------------------------------------ */
/* kernel body                                                          */
/* ------------------------------------------------------------------ */
__device__ __forceinline__ void StreamKernelBody(const LaunchArgs &a, int G, float *smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s0 = blockIdx.x * G;
  const int ng = min(G, a.n_streams - s0);
  const int T = a.n_blocks;
  if (warp < G) {
    /* ---- receiver warp ---- */
    const bool live = warp < ng;
    RxWarp w(a, smem + warp * kSlotF, s0 + warp, lane);
    if (live) {
      w.LoadState();
      w.IssueQuarter(0, 0);
    }
    w.tm.Start(blockIdx.x == 0 && warp == 0 && lane == 0, 0);
    for (int k = 0; k < T + 2; ++k) {
      if (live) {
        if (k >= 2) w.BackEnd(k - 2, k & 1);
        if (k < T) w.FrontEnd(k, k & 1);
      }
      T41RX_LAP(w.tm, 11);
      __syncthreads();
      T41RX_LAP(w.tm, 12);
    }
    if (live) w.StoreState();
  } else {
    /* ---- AGC warp ---- */
    AgcLane g;
    const bool mine = lane < ng;
    bool active = false;
    if (mine) {
      const StreamCfg &cf = a.cfg[s0 + lane];
      g.Load(cf, a.st[s0 + lane]);
      active = (cf.mode != kModePsk31) && (cf.agc_mode != 0);
    }
    float *slot = smem + (mine ? lane : 0) * kSlotF;
    SectionTimer tm;
    tm.Start(blockIdx.x == 0 && lane == 0, 16);
    for (int k = 0; k < T + 2; ++k) {
      if (k >= 1 && k <= T) AgcBlock(g, slot + oStA + ((k - 1) & 1) * 512, active);
      T41RX_LAP(tm, 0);
      __syncthreads();
      T41RX_LAP(tm, 1);
    }
    if (mine && active) g.Store(a.st[s0 + lane]);
  }
}

}  // namespace fast
}  // namespace t41rx
#endif
