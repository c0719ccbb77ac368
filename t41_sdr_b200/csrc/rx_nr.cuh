/*
 * rx_nr.cuh — the default-off audio stages between the demodulator and the interpolators that work on 256-point
 * spectra or on linear prediction (rest):
 *
 *   Kim1_NR                  Noise.cpp:108-311   Kim & Ruwisch 2002 noise estimator, then x 30    (Process.cpp:845-849)
 *   SpectralNoiseReduction   Noise.cpp:379-655   Ephraim-Malah style spectral weighting            (Process.cpp:850-852)
 *   NoiseBlanker             DSP_Fn.cpp:105-362  LPC impulse detection and two-sided prediction    (Process.cpp:873-876)
 *
 * Form: ONE LANE PER RECEIVER runs a block's stage from start to end (phases PhNrSpectral / PhNoiseBlank of the
 * bit-exact chain), every float operation in the reference's order and precision, so the results are bit-identical
 * to the reference build.  These stages are off by default and their receivers are routed to the bit-exact chain
 * (like the LMS stages); nothing here is on the throughput kernel's path.  State: NrState (rx_types.h), one per
 * receiver, allocated when the first receiver switches one of the stages on.  Scratch: the receiver's shared-memory
 * slot.  The same functions compile for the host emulation (tests/devtools).
 *
 * Tables (a.nr_tab, host-computed with the host's libm exactly like the reference's expressions, rx_design.cpp):
 *   [0..4]     ax, ap, xih1r, pfac, snr_prio_min       (Noise.cpp:404-413)
 *   [8..263]   Kim1_NR's Hann window                    (Noise.cpp:198-201)
 *   [264..519] sqrtHann                                 (Noise.cpp:55-88)
 */
#ifndef T41RX_NR_CUH
#define T41RX_NR_CUH

namespace t41rx {

constexpr int kNrL = 256, kNrHalf = 128, kNrLFrames = 3, kNrNFrames = 15;
constexpr int kNrTabConsts = 0, kNrTabKimWin = 8, kNrTabSqrtHann = 264, kNrTabLen = 520;
constexpr float kNrPsi = 0.0, kNrAlpha = 0.95, kNrBeta = 0.85;     /* EEPROM.cpp:71-73 */

/* 256-point complex FFT of buf[512] (interleaved), natural-order result, arm_cfft_f32(len256, ifft, 1) as the oracle
   defines it (oracle/cmsis_port.c: one radix-4 stage over the quarters, two radix-8 stages inside each quarter, mixed
   digit reversal; inverse = conjugate, forward, conjugate and scale).  tw512: the 512-point twiddle table, whose even
   entries are the 256-point twiddles to the bit.  tmp: 512 floats of scratch.  One thread. */
T41RX_DEV void NrTwiddleMul(const float2 *tw512, int t, float &re, float &im) {
  const float2 w = tw512[2 * t];
  const float rc = re * w.x, is = im * w.y;
  const float ic = im * w.x, rs = re * w.y;
  re = rc + is;
  im = ic - rs;
}

T41RX_DEV_NOINLINE void NrFft256(float *buf, float *tmp, const float2 *tw512, bool inverse) {
  if (inverse) {
    for (int i = 0; i < kNrL; ++i) buf[2 * i + 1] = -buf[2 * i + 1];
  }
  for (int j = 0; j < 64; ++j) {
    const float ar = buf[2 * j], ai = buf[2 * j + 1];
    const float br = buf[2 * (j + 64)], bi = buf[2 * (j + 64) + 1];
    const float cr = buf[2 * (j + 128)], ci = buf[2 * (j + 128) + 1];
    const float dr = buf[2 * (j + 192)], di = buf[2 * (j + 192) + 1];
    const float s0r = ar + cr, s0i = ai + ci;
    const float d0r = ar - cr, d0i = ai - ci;
    const float s1r = br + dr, s1i = bi + di;
    const float d1r = br - dr, d1i = bi - di;
    float y0r = s0r + s1r, y0i = s0i + s1i;
    float y2r = s0r - s1r, y2i = s0i - s1i;
    float y1r = d0r + d1i, y1i = d0i - d1r;
    float y3r = d0r - d1i, y3i = d0i + d1r;
    if (j != 0) {
      NrTwiddleMul(tw512, j, y1r, y1i);
      NrTwiddleMul(tw512, 2 * j, y2r, y2i);
      NrTwiddleMul(tw512, 3 * j, y3r, y3i);
    }
    buf[2 * j] = y0r;           buf[2 * j + 1] = y0i;
    buf[2 * (j + 64)] = y1r;    buf[2 * (j + 64) + 1] = y1i;
    buf[2 * (j + 128)] = y2r;   buf[2 * (j + 128) + 1] = y2i;
    buf[2 * (j + 192)] = y3r;   buf[2 * (j + 192) + 1] = y3i;
  }
  for (int q = 0; q < 4; ++q) {
    float *b = buf + 2 * 64 * q;
    for (int pass = 0; pass < 2; ++pass) {
      const int n1 = pass ? 8 : 64, n2 = n1 >> 3, stride = pass ? 32 : 4;
      for (int j = 0; j < n2; ++j) {
        for (int i0 = j; i0 < 64; i0 += n1) {
          float r[8], im[8];
          for (int m = 0; m < 8; ++m) {
            r[m] = b[2 * (i0 + m * n2)];
            im[m] = b[2 * (i0 + m * n2) + 1];
          }
          Dft8(r, im);
          b[2 * i0] = r[0];
          b[2 * i0 + 1] = im[0];
          for (int k = 1; k < 8; ++k) {
            float re = r[k], ie = im[k];
            if (j != 0) NrTwiddleMul(tw512, j * k * stride, re, ie);
            b[2 * (i0 + k * n2)] = re;
            b[2 * (i0 + k * n2) + 1] = ie;
          }
        }
      }
    }
  }
  for (int i = 0; i < 2 * kNrL; ++i) tmp[i] = buf[i];
  for (int p = 0; p < kNrL; ++p) {          /* position 64 q + 8 d1 + d0 holds bin 4 (8 d0 + d1) + q */
    const int k = 4 * (8 * (p & 7) + ((p >> 3) & 7)) + (p >> 6);
    buf[2 * k] = tmp[2 * p];
    buf[2 * k + 1] = tmp[2 * p + 1];
  }
  if (inverse) {
    const float inv = 1.0f / (float)kNrL;
    for (int i = 0; i < kNrL; ++i) {
      buf[2 * i] = buf[2 * i] * inv;
      buf[2 * i + 1] = -buf[2 * i + 1] * inv;
    }
  }
}

/* frame k of a block: the previous 128 samples, then samples 128 k .. 128 k + 127 of the block; imaginary parts zero */
T41RX_DEV void NrLoadFrame(NrState &ns, const float *aud, int k, float *buf) {
  for (int i = 0; i < kNrHalf; ++i) {
    buf[2 * i] = ns.last_sample[i];
    buf[2 * i + 1] = 0.0f;
  }
  for (int i = 0; i < kNrHalf; ++i) ns.last_sample[i] = aud[i + k * kNrHalf];
  for (int i = 0; i < kNrHalf; ++i) {
    buf[kNrL + 2 * i] = aud[i + k * kNrHalf];
    buf[kNrL + 2 * i + 1] = 0.0f;
  }
}

/* Kim1_NR over the 256 samples at aud, in place; scratch: buf[512], tmp[512], out[256] */
T41RX_DEV_NOINLINE void KimNrLane(NrState &ns, const StreamCfg &cf, const float *tab, const float2 *tw512, float *aud, float *buf,
                         float *tmp, float *out) {
  const float onemalpha = (1.0 - kNrAlpha);
  const float onemtwobeta = (1.0 - (2.0 * kNrBeta));
  const int lo = cf.nr_vad_lo, hi = cf.nr_vad_hi;
  const float *win = tab + kNrTabKimWin;
  for (int k = 0; k < 2; ++k) {
    NrLoadFrame(ns, aud, k, buf);
    for (int i = 0; i < kNrL; ++i) buf[2 * i] *= win[i];
    NrFft256(buf, tmp, tw512, false);
    const int xp = (int)ns.x_ptr, ep = (int)ns.e_ptr;
    for (int i = 0; i < kNrHalf; ++i) ns.X[i][xp] = (buf[2 * i] * buf[2 * i] + buf[2 * i + 1] * buf[2 * i + 1]);
    for (int i = lo; i < hi; ++i) {
      float sum = 0.0f;
      for (int j = 0; j < kNrLFrames; ++j) sum = sum + ns.X[i][j];
      ns.E[i][ep] = sum / (float)kNrLFrames;
    }
    for (int i = lo; i < hi; ++i) {
      float m = ns.E[i][0];
      for (int j = 1; j < kNrNFrames; ++j)
        if (ns.E[i][j] < m) m = ns.E[i][j];
      ns.M[i] = m;
    }
    for (int i = lo; i < hi; ++i) {
      const float t = ns.X[i][xp] / ns.M[i];
      ns.lambda[i] = (t > kNrPsi) ? ns.M[i] : ns.E[i][ep];
    }
    for (int i = lo; i < hi; ++i) {
      float g = 1.0 - (ns.lambda[i] / ns.E[i][ep]);            /* NR_KIM_K = 1, NR_use_X = 0 */
      if (g < 0.0) g = 0.0;
      ns.G[i] = g;
      ns.Gts[i][0] = kNrAlpha * ns.Gts[i][1] + onemalpha * g;
      ns.Gts[i][1] = ns.Gts[i][0];
    }
    for (int i = 1; i < kNrHalf - 1; ++i)
      ns.G[i] = kNrBeta * ns.Gts[i - 1][0] + onemtwobeta * ns.Gts[i][0] + kNrBeta * ns.Gts[i + 1][0];
    ns.G[0] = (onemtwobeta + kNrBeta) * ns.Gts[0][0] + kNrBeta * ns.Gts[1][0];
    ns.G[kNrHalf - 1] = kNrBeta * ns.Gts[kNrHalf - 2][0] + (onemtwobeta + kNrBeta) * ns.Gts[kNrHalf - 1][0];
    for (int i = 0; i < kNrHalf; ++i) {         /* the upper half pairs bin 255 - i with bin i, as the reference writes it */
      const float g = ns.G[i];
      buf[2 * i] = buf[2 * i] * g;
      buf[2 * i + 1] = buf[2 * i + 1] * g;
      buf[2 * kNrL - 2 * i - 2] = buf[2 * kNrL - 2 * i - 2] * g;
      buf[2 * kNrL - 2 * i - 1] = buf[2 * kNrL - 2 * i - 1] * g;
    }
    ns.x_ptr = (xp + 1 >= kNrLFrames) ? 0 : xp + 1;
    ns.e_ptr = (ep + 1 >= kNrNFrames) ? 0 : ep + 1;
    NrFft256(buf, tmp, tw512, true);
    for (int i = 0; i < kNrHalf; ++i) out[i + k * kNrHalf] = buf[2 * i] + ns.last_ifft[i];
    for (int i = 0; i < kNrHalf; ++i) ns.last_ifft[i] = buf[kNrL + 2 * i];
  }
  for (int i = 0; i < kNrL; ++i) aud[i] = out[i] * 30.0f;       /* Process.cpp:847 */
}

/* SpectralNoiseReduction over the 256 samples at aud, in place.  Reference quirks kept (see the oracle's restatement):
   the first 20 frames only train the noise estimate and leave the audio untouched; everything from the weighting to the
   overlap-add sits in the `trained` branch; the musical-noise smoothing runs once per bin of the gain loop; the
   long-tone gain is never written by the reference (zero): the trained stage's output is (signed) zero, which also
   makes the device's expf (a different last bit than the host's libm at times) invisible outside NrState. */
T41RX_DEV_NOINLINE void SpectralNrLane(NrState &ns, const StreamCfg &cf, const float *tab, const float2 *tw512, float *aud, float *buf,
                              float *tmp, float *ph1y) {
  const float psthr = 0.99, pnsaf = 0.01, psini = 0.5, power_threshold = 0.4;
  const int nr_width = 4;
  const float ax = tab[0], ap = tab[1], xih1r = tab[2], pfac = tab[3], snr_prio_min = tab[4];
  const float *sqrt_hann = tab + kNrTabSqrtHann;
  const int lo = cf.nr_vad_lo, hi = cf.nr_vad_hi;
  if (ns.spectral_stage == 0) {
    for (int i = 0; i < kNrHalf; ++i) {
      ns.last_sample[i] = 0.0f;
      ns.G[i] = 1.0f;
      ns.Hk_old[i] = 1.0f;
      ns.Nest[i][0] = 0.0f;
      ns.Nest[i][1] = 1.0f;
      ns.pslp[i] = 0.5f;
    }
    ns.spectral_stage = 1;
  }
  for (int k = 0; k < 2; ++k) {
    NrLoadFrame(ns, aud, k, buf);
    for (int i = 0; i < kNrL; ++i) buf[2 * i] *= sqrt_hann[i];
    NrFft256(buf, tmp, tw512, false);
    for (int i = 0; i < kNrHalf; ++i) ns.X[i][0] = (buf[2 * i] * buf[2 * i] + buf[2 * i + 1] * buf[2 * i + 1]);
    if (ns.spectral_stage == 1) {
      for (int i = 0; i < kNrHalf; ++i) {
        ns.Nest[i][0] = ns.Nest[i][0] + 0.05 * ns.X[i][0];
        ns.xt[i] = psini * ns.Nest[i][0];
      }
      ns.init_counter = (ns.init_counter + 1) & 255;
      if (ns.init_counter > 19) {
        ns.init_counter = 0;
        ns.spectral_stage = 2;
      }
    }
    if (ns.spectral_stage == 2) {
      for (int i = 0; i < kNrHalf; ++i) {
        float p = 1.0 / (1.0 + pfac * expf(xih1r * ns.X[i][0] / ns.xt[i]));
        ns.pslp[i] = ap * ns.pslp[i] + (1.0 - ap) * p;
        if (ns.pslp[i] > psthr) p = 1.0 - pnsaf;
        else p = fmin((double)p, 1.0);
        ph1y[i] = p;
        const float xtr = (1.0 - p) * ns.X[i][0] + p * ns.xt[i];
        ns.xt[i] = ax * ns.xt[i] + (1.0 - ax) * xtr;
      }
      for (int i = 0; i < kNrHalf; ++i) {
        ns.SNR_post[i] = fmax(fmin((double)(ns.X[i][0] / ns.xt[i]), 1000.0), (double)snr_prio_min);
        ns.SNR_prio[i] = fmax(kNrAlpha * ns.Hk_old[i] + (1.0 - kNrAlpha) * fmax(ns.SNR_post[i] - 1.0, 0.0), 0.0);   /* all-double fmax */
      }
      for (int i = lo; i < hi; ++i) {
        const float v = ns.SNR_prio[i] * ns.SNR_post[i] / (1.0 + ns.SNR_prio[i]);
        ns.G[i] = 1.0 / ns.SNR_post[i] * sqrtf((0.7212 * v + v * v));
        ns.Hk_old[i] = ns.SNR_post[i] * ns.G[i] * ns.G[i];
        float pre_power = 0.0f, post_power = 0.0f;
        for (int m = lo; m < hi; ++m) {
          pre_power += ns.X[m][0];
          post_power += ns.G[m] * ns.G[m] * ns.X[m][0];
        }
        float power_ratio = post_power / pre_power;
        int nn;
        if (power_ratio > power_threshold) {
          power_ratio = 1.0f;
          nn = 1;
        } else {
          nn = (int16_t)(1 + 2 * (int)(0.5 + nr_width * (1.0 - power_ratio / power_threshold)));
        }
        for (int b = lo + nn / 2; b < hi - nn / 2; ++b) {
          float acc = 0.0f;
          for (int m = b - nn / 2; m <= b + nn / 2; ++m) acc += ns.G[m];
          ns.Nest[b][0] = acc / (float)nn;
        }
        for (int b = lo; b < lo + nn / 2; ++b) {
          float acc = 0.0f;
          for (int m = b; m < b + nn; ++m) acc += ns.G[m];
          ns.Nest[b][0] = acc / (float)nn;
        }
        for (int b = hi - nn; b < hi; ++b) {
          float acc = 0.0f;
          for (int m = b; m > b - nn; --m) acc += ns.G[m];
          ns.Nest[b][0] = acc / (float)nn;
        }
        for (int b = lo + nn / 2; b < hi - nn / 2; ++b) ns.G[b] = ns.Nest[b][0];
      }
      for (int i = 0; i < kNrHalf; ++i) {
        const float g = ns.G[i], lt = ns.long_tone_gain[i];
        buf[2 * i] = buf[2 * i] * g * lt;
        buf[2 * i + 1] = buf[2 * i + 1] * g * lt;
        buf[2 * kNrL - 2 * i - 2] = buf[2 * kNrL - 2 * i - 2] * g * lt;
        buf[2 * kNrL - 2 * i - 1] = buf[2 * kNrL - 2 * i - 1] * g * lt;
      }
      NrFft256(buf, tmp, tw512, true);
      for (int i = 0; i < kNrL; ++i) buf[2 * i] *= sqrt_hann[i];
      for (int i = 0; i < kNrHalf; ++i) aud[i + k * kNrHalf] = buf[2 * i] + ns.last_ifft[i];
      for (int i = 0; i < kNrHalf; ++i) ns.last_ifft[i] = buf[kNrL + 2 * i];
    }
  }
}

/* NoiseBlanker / AltNoiseBlanking over the 256 samples at x, in place; scratch: fs[266] (FIR state), ts[256] */
T41RX_DEV_NOINLINE void NoiseBlankLane(NrState &ns, float *x, float *fs, float *ts) {
  constexpr int kN = 256, kOrder = 10, kImp = 7, kPl = 3, kBoundary = 14;
  const float nb_thresh = 2.5;
  int pos[20];
  float lpcs[kOrder + 1], rev[kOrder + 1], any[kOrder + 1], R[kOrder + 1];
  float rfw[kImp + kOrder], rbw[kImp + kOrder], wfw[kImp], wbw[kImp];
  for (int i = 0; i < kImp; ++i) {
    wbw[i] = 1.0 * i / (kImp - 1);
    wfw[kImp - i - 1] = wbw[i];
  }
  for (int i = 0; i <= kOrder; ++i) {                 /* autocorrelation (arm_dot_prod_f32: one fused multiply-add per term) */
    float acc = 0.0f;
    for (int n = 0; n < kN - i; ++n) acc = fmaf(x[n], x[n + i], acc);
    R[i] = acc;
  }
  R[0] = R[0] * (1.0 + 1.0e-9);
  lpcs[0] = 1.0f;
  for (int i = 1; i <= kOrder; ++i) lpcs[i] = 0.0f;
  float alfa = R[0];
  for (int m = 1; m <= kOrder; ++m) {                 /* Levinson-Durbin */
    float s = 0.0f;
    for (int u = 1; u < m; ++u) s = s + lpcs[u] * R[m - u];
    const float k = -(R[m] + s) / alfa;
    for (int v = 1; v < m; ++v) any[v] = lpcs[v] + k * lpcs[m - v];
    for (int w = 1; w < m; ++w) lpcs[w] = any[w];
    lpcs[m] = k;
    alfa = alfa * (1 - k * k);
  }
  for (int o = 0; o <= kOrder; ++o) rev[kOrder - o] = lpcs[o];
  /* arm_fir_f32 twice, each from a cleared state: y[n] = sum_i c[i] w[n + i] over the window that ends at sample n */
  for (int pass = 0; pass < 2; ++pass) {
    const float *c = pass ? lpcs : rev;
    const float *src = pass ? ts : x;
    for (int i = 0; i < kOrder; ++i) fs[i] = 0.0f;
    for (int n = 0; n < kN; ++n) {
      fs[kOrder + n] = src[n];
      float acc = 0.0f;
      for (int i = 0; i <= kOrder; ++i) acc = fmaf(fs[n + i], c[i], acc);
      ts[n] = acc;
    }
  }
  float sum = 0.0f, sumsq = 0.0f;                     /* arm_var_f32 */
  for (int n = 0; n < kN; ++n) {
    sum += ts[n];
    sumsq = fmaf(ts[n], ts[n], sumsq);
  }
  const float sigma2 = sumsq / (float)(kN - 1) - (sum * sum) / ((float)kN * (float)(kN - 1));
  float lpc_power = 0.0f;                             /* arm_power_f32 over lpcs[0 .. order - 1] */
  for (int i = 0; i < kOrder; ++i) lpc_power = fmaf(lpcs[i], lpcs[i], lpc_power);
  const float thr = nb_thresh * sqrtf(sigma2 * lpc_power);
  int search = kOrder + kPl, count = 0;
  do {
    if ((ts[search] > thr) || (ts[search] < (-thr))) {
      pos[count] = search - kOrder;
      count++;
      search += kPl;
    }
    search++;
  } while (((unsigned)search < (unsigned)(kN - kBoundary)) && ((unsigned)count < 20u));
  for (int i = 1; i <= kOrder; ++i) lpcs[i] = -lpcs[i];
  for (int i = 0; i < kOrder; ++i) rev[i] = -rev[i];
  for (int j = 0; j < count; ++j) {
    for (int q = 0; q < kOrder; ++q) {
      if ((pos[j] - kPl - kOrder + q) < 0) rfw[q] = ns.nb_last_frame_end[pos[j] + q];
      else rfw[q] = x[pos[j] - kPl - kOrder + q];
      rbw[kImp + q] = x[pos[j] + kPl + q + 1];
    }
    for (int i = 0; i < kImp; ++i) {
      float f = 0.0f, b = 0.0f;
      for (int q = 0; q < kOrder; ++q) f = fmaf(rev[q], rfw[i + q], f);
      rfw[i + kOrder] = f;
      for (int q = 0; q < kOrder; ++q) b = fmaf(lpcs[1 + q], rbw[kImp - i + q], b);
      rbw[kImp - i - 1] = b;
    }
    for (int i = 0; i < kImp; ++i) x[pos[j] - kPl + i] = wfw[i] * rfw[kOrder + i] + wbw[i] * rbw[i];
  }
  for (int p = 0; p < kOrder + kPl; ++p) ns.nb_last_frame_end[p] = x[kN - 1 - kOrder - kPl + p];
}

}  // namespace t41rx
#endif
