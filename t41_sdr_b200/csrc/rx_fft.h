/*
 * rx_fft.h — the 512-point complex FFT building block shared by host (mask design) and
 * device (fast-convolution filter, display spectrum).
 *
 * 512 = 8^3: three radix-8 decimation-in-frequency passes over an in-place buffer followed
 * by base-8 digit reversal, the structure of CMSIS-DSP's arm_cfft_f32 for length 512
 * (reference call sites Process.cpp:535,595,787,808; Filter.cpp:282; FFT.cpp:134,226).
 * Every add / multiply is individually rounded (no FMA contraction: build with
 * --fmad=false / -ffp-contract=off) in exactly the order written here, which is the order
 * the CPU oracle defines (oracle/cmsis_port.c), so that host, device and oracle agree to
 * the bit.
 */
#ifndef T41RX_FFT_H
#define T41RX_FFT_H

#include <vector_types.h>   /* float2 (header-only, also fine for a plain host compiler) */

#ifdef __CUDACC__
#define T41RX_HD __host__ __device__ __forceinline__
#else
#define T41RX_HD inline
#endif

namespace t41rx {

constexpr float kC81 = 0.70710678118f;

/* octal digit reversal of a 9-bit index */
T41RX_HD unsigned OctRev3(unsigned p) { return ((p & 7u) << 6) | (p & 0x38u) | ((p >> 6) & 7u); }

/* 8-point DFT, natural-order output.  r[], i[] are overwritten. */
T41RX_HD void Dft8(float *r, float *i) {
  const float ar0 = r[0] + r[4], ai0 = i[0] + i[4];
  const float br0 = r[0] - r[4], bi0 = i[0] - i[4];
  const float ar1 = r[1] + r[5], ai1 = i[1] + i[5];
  const float br1 = r[1] - r[5], bi1 = i[1] - i[5];
  const float ar2 = r[2] + r[6], ai2 = i[2] + i[6];
  const float br2 = r[2] - r[6], bi2 = i[2] - i[6];
  const float ar3 = r[3] + r[7], ai3 = i[3] + i[7];
  const float br3 = r[3] - r[7], bi3 = i[3] - i[7];

  const float cr0 = ar0 + ar2, ci0 = ai0 + ai2;
  const float dr0 = ar0 - ar2, di0 = ai0 - ai2;
  const float cr1 = ar1 + ar3, ci1 = ai1 + ai3;
  const float dr1 = ar1 - ar3, di1 = ai1 - ai3;
  r[0] = cr0 + cr1; i[0] = ci0 + ci1;
  r[4] = cr0 - cr1; i[4] = ci0 - ci1;
  r[2] = dr0 + di1; i[2] = di0 - dr1;
  r[6] = dr0 - di1; i[6] = di0 + dr1;

  const float p = (br1 - br3) * kC81;
  const float q = (br1 + br3) * kC81;
  const float u = (bi1 - bi3) * kC81;
  const float v = (bi1 + bi3) * kC81;
  const float er0 = br0 + bi2, ei0 = bi0 - br2;
  const float fr0 = br0 - bi2, fi0 = bi0 + br2;
  const float er1 = p + v, ei1 = u - q;
  const float fr1 = v - p, fi1 = q + u;
  r[1] = er0 + er1; i[1] = ei0 + ei1;
  r[5] = er0 - er1; i[5] = ei0 - ei1;
  r[3] = fr0 + fr1; i[3] = fi0 - fi1;
  r[7] = fr0 - fr1; i[7] = fi0 + fi1;
}

/*
 * One radix-8 butterfly of pass `pass` (0,1,2) of the 512-point transform.
 * `b` in [0,64) enumerates the butterflies of the pass; buf is 512 interleaved complex.
 * tw: 512 (cos, sin) pairs of 2*pi*k/512.
 */
/* Where element i of a 512-point buffer lives in the phase kernels' shared memory.  With the elements in natural order
 * every pass but the first collides on the banks: pass 1's lanes are four groups 512 bytes apart (4 wavefronts where 2
 * would do), pass 2's lanes each walk a contiguous 64 bytes (16 where 2 would do; ncu: the butterflies' loads and stores
 * were 2/3 of the bit-exact front kernel's bank conflicts).  The low four bits of the index (the 16 eight-byte columns of
 * a 128-byte row) are XORed with bits of the row number so that in every pass the 16 lanes of a half-warp fall into 16
 * different columns:
 *   pass 0  i = b + 64 m        a half-warp is 16 consecutive elements of one row: any XOR keeps them apart;
 *   pass 1  i = 64 q + 8 m + j  lanes (q, j), q = 0..1 | 2..3: bit 3 ^= q & 1 parts the two q of a half-warp;
 *   pass 2  i = 8 b + m         lanes b: column = 8 (b & 1) + m; low three bits ^= (b >> 1) & 7 parts the eight pairs.
 * Both are bits of i >> 4 (bit 6 of i is q & 1 in pass 1 and bit 2 of b >> 1 in pass 2). */
T41RX_HD int FftPhys(int i) { return i ^ (((i >> 4) & 7) | (((i >> 6) & 1) << 3)); }

/* kSwz: the buffer is in FftPhys order (shared memory of the phase kernels); else natural order (host) */
template <bool kSwz = false>
T41RX_HD void Radix8Butterfly(float2 *buf, const float2 *tw, int pass, int b) {
  int n2, j, i0, stride;
  if (pass == 0) { n2 = 64; j = b; i0 = b; stride = 1; }
  else if (pass == 1) { n2 = 8; j = b & 7; i0 = (b >> 3) * 64 + j; stride = 8; }
  else { n2 = 1; j = 0; i0 = b * 8; stride = 64; }
  float r[8], im[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float2 x = buf[kSwz ? FftPhys(i0 + m * n2) : i0 + m * n2];
    r[m] = x.x;
    im[m] = x.y;
  }
  Dft8(r, im);
  buf[kSwz ? FftPhys(i0) : i0] = float2{r[0], im[0]};
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    float re = r[k], ie = im[k];
    if (j != 0) {
      const float2 w = tw[j * k * stride];
      const float rc = r[k] * w.x, is = im[k] * w.y;
      const float ic = im[k] * w.x, rs = r[k] * w.y;
      re = rc + is;
      ie = ic - rs;
    }
    buf[kSwz ? FftPhys(i0 + k * n2) : i0 + k * n2] = float2{re, ie};
  }
}

}  // namespace t41rx
#endif
