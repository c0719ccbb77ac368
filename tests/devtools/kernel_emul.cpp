/*
 * tests/devtools/kernel_emul.cpp — DEVELOPMENT / TEST AID, never part of libt41rx.so.
 *
 * Compiles the device phase functions of t41_sdr_b200/csrc/rx_phases.cuh for the host
 * (T41RX_HOST_EMUL) and steps them exactly as the CUDA kernel does: for every phase, every
 * thread id of a CTA in turn, then the next phase (the barrier).  That lets the CPU-only test
 * tier check the kernel's index arithmetic, shared-memory overlays, state hand-over and
 * bit-exactness against the oracle before GPU time is spent.  It is orders of magnitude
 * slower than the oracle and is not a processing path of the product.
 */
#define T41RX_HOST_EMUL 1
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/t41rx.h"
#include "../../t41_sdr_b200/csrc/rx_host.h"
#include "../../t41_sdr_b200/csrc/rx_phases.cuh"

using namespace t41rx;

struct Emul {
  HostModel host;
  std::vector<StreamState> state;
  std::vector<float> smem;
  /* audio-spectrum by-product: where emul_process puts it (emul_bind_audio_spectrum) and the scratch */
  int32_t *bind_ypixel = nullptr;
  float *bind_max_ave = nullptr;
  uint8_t *bind_spec_frames = nullptr, *bind_audio_frames = nullptr;
  std::vector<float2> aspec;
  std::vector<NrState> nr;
};

extern "C" {

Emul *emul_create(int n_streams) {
  Emul *e = new Emul();
  e->host.Init(n_streams);
  e->state.resize(n_streams);
  for (int s = 0; s < n_streams; ++s) HostStateInit(&e->state[s]);
  e->smem.assign(kSmemFloats + 64, 0.0f);
  e->nr.resize(n_streams);
  memset(e->nr.data(), 0, sizeof(NrState) * n_streams);
  return e;
}

void emul_destroy(Emul *e) { delete e; }

int emul_set_params_each(Emul *e, int first, int count, const t41rx_params *p) {
  for (int i = 0; i < count; ++i) {
    StatePatch patch;
    int new_fset = -1;
    if (e->host.Apply(first + i, p[i], &patch, &new_fset)) return T41RX_EINVAL;
    if (patch.set_rf_gain) e->state[first + i].rf_gain = patch.rf_gain;
    if (patch.reset_zoom_ptr) e->state[first + i].zoom_ptr = 0;
  }
  return 0;
}

/* emulation-only flag: produce the spectrum rows with the rows-only schedule (T41RX_ROWS_SCHEDULE) */
static const uint32_t kEmulSplitRows = 0x100u;
/* emulation-only flag: the chain as the product's front | serial | back kernels (T41RX_FRONT_SCHEDULE over every CTA,
   SerialReceiver for every receiver, T41RX_BACK_SCHEDULE over every CTA) instead of the fused schedule */
static const uint32_t kEmulSplitExact = 0x200u;

int emul_process(Emul *e, const float *iq, float *audio, int n_blocks, int row_every, int16_t *spec_rows,
                 uint16_t *wf_rows, int8_t *psk_bits, uint8_t *psk_chars, uint32_t flags) {
  HostModel &h = e->host;
  LaunchArgs a;
  memset(&a, 0, sizeof(a));
  a.iq = iq;
  a.audio = audio;
  a.spec_rows = row_every > 0 ? spec_rows : nullptr;
  a.wf_rows = row_every > 0 ? wf_rows : nullptr;
  a.psk_bits = psk_bits;
  a.psk_chars = psk_chars;
  a.cfg = h.cfg.data();
  a.st = e->state.data();
  a.fsets = h.fsets.data();
  a.nco_tab = h.nco_tab.data();
  a.twiddle = reinterpret_cast<const float2 *>(h.twiddle.data());
  a.hann = h.hann.data();
  a.sin_table = h.sin_table.data();
  a.zoom_iir = h.zoom_iir.data();
  a.eq_coeffs = h.eq_coeffs.data();
  a.cw_coeffs = h.cw_coeffs.data();
  a.sam_consts = h.sam_consts.data();
  a.nr = e->nr.data();
  a.nr_tab = h.nr_tab.data();
  a.gradient = h.gradient.data();
  a.varicode = h.varicode.data();
  a.n_streams = h.n_streams;
  a.n_blocks = n_blocks;
  a.row_every = row_every;
  a.n_rows = row_every > 0 ? (n_blocks + row_every - 1) / row_every : 0;
  a.t0 = 0;
  a.t_stride = n_blocks;
  a.flags = flags;
  if (a.n_rows > 0 && e->bind_spec_frames && a.spec_rows) a.spec_frames = e->bind_spec_frames;
  if (a.n_rows > 0 && (e->bind_ypixel || e->bind_max_ave || e->bind_audio_frames)) {
    a.audio_frames = e->bind_audio_frames;
    e->aspec.assign((size_t)h.n_streams * a.n_rows * kFft, float2{NAN, NAN});
    a.aspec = e->aspec.data();
    a.audio_ypixel = e->bind_ypixel;
    a.audio_max_ave = e->bind_max_ave;
  }

  /* 16-byte aligned shared-memory stand-in */
  float *smem = e->smem.data();
  while ((reinterpret_cast<uintptr_t>(smem) & 15u) != 0) ++smem;

  const int grid = (h.n_streams + kG - 1) / kG;
  const bool split_exact = (flags & kEmulSplitExact) != 0;
  std::vector<float4> ser_in;
  std::vector<float> ser_out;
  if (split_exact) {
    ser_in.assign((size_t)h.n_streams * n_blocks * kDec, float4{NAN, NAN, NAN, NAN});
    ser_out.assign((size_t)h.n_streams * n_blocks * kDec, NAN);
    a.ser_in = ser_in.data();
    a.ser_out = ser_out.data();
#define EMUL_PHASE(stmt) \
  do {                   \
    for (int tid = 0; tid < kNT; ++tid) { stmt; } \
  } while (0)
    auto cta_of = [&](int cta) {
      Cta c;
      c.a = a;
      c.smem = smem;
      c.s0 = cta * kG;
      c.ng = (h.n_streams - c.s0 < kG) ? (h.n_streams - c.s0) : kG;
      c.t = 0;
      c.row = 0;
      c.row_idx = 0;
      c.rows_only = 0;
      c.casc_warp = 0;
      c.dc_carried = 0;
      for (int i = 0; i < kSmemFloats; ++i) smem[i] = NAN;
      for (int tid = 0; tid < kNT; ++tid) PhCtaInit(c, tid);
      return c;
    };
    for (int cta = 0; cta < grid; ++cta) {          /* t41rx_exact_front_kernel */
      Cta c = cta_of(cta);
      EMUL_PHASE(PhStateIn(c, tid));
      for (int t = 0; t < n_blocks; ++t) {
        c.t = t;
        c.row = (row_every > 0) && (t % row_every == 0);
        c.row_idx = c.row ? t / row_every : 0;
        T41RX_FRONT_SCHEDULE(EMUL_PHASE)
      }
      EMUL_PHASE(PhFrontStateOut(c, tid));
    }
    for (int r = 0; r < h.n_streams; ++r) SerialReceiver(a, r, h.sin_table.data());   /* t41rx_exact_serial_kernel */
    for (int cta = 0; cta < grid; ++cta) {          /* t41rx_exact_back_kernel */
      Cta c = cta_of(cta);
      EMUL_PHASE(PhBackStateIn(c, tid));
      for (int t = 0; t < n_blocks; ++t) {
        c.t = t;
        T41RX_BACK_SCHEDULE(EMUL_PHASE)
      }
      EMUL_PHASE(PhBackStateOut(c, tid));
    }
#undef EMUL_PHASE
  }
  for (int cta = 0; cta < grid; ++cta) {
    Cta c;
    c.a = a;
    c.smem = smem;
    c.s0 = cta * kG;
    c.ng = (h.n_streams - c.s0 < kG) ? (h.n_streams - c.s0) : kG;
    c.t = 0;
    c.row = 0;
    c.row_idx = 0;
    c.rows_only = 0;
    c.casc_warp = 0;
    c.dc_carried = 0;
    /* poison: shared memory is uninitialised at CTA start on the GPU */
    for (int i = 0; i < kSmemFloats; ++i) smem[i] = NAN;
    for (int tid = 0; tid < kNT; ++tid) PhCtaInit(c, tid);
#define EMUL_PHASE(stmt) \
  do {                   \
    for (int tid = 0; tid < kNT; ++tid) { stmt; } \
  } while (0)
    const bool split_rows = (flags & kEmulSplitRows) != 0 && row_every > 0;
    if (split_rows) {
      /* what the product does around the throughput kernel: the rows-only schedule first, on the state
         of launch start, then the chain itself without rows */
      c.rows_only = 1;
      for (int t = 0; t < n_blocks; t += row_every) {
        c.t = t;
        c.row = 1;
        c.row_idx = t / row_every;
        T41RX_ROWS_SCHEDULE(EMUL_PHASE)
      }
      c.rows_only = 0;
      for (int i = 0; i < kSmemFloats; ++i) smem[i] = NAN;
      for (int tid = 0; tid < kNT; ++tid) PhCtaInit(c, tid);
    }
    if (!split_exact) {
      EMUL_PHASE(PhStateIn(c, tid));
      for (int t = 0; t < n_blocks; ++t) {
        c.t = t;
        c.row = !split_rows && (row_every > 0) && (t % row_every == 0);
        c.row_idx = c.row ? t / row_every : 0;
        T41RX_BLOCK_SCHEDULE(EMUL_PHASE)
      }
      EMUL_PHASE(PhStateOut(c, tid));
    }
    if (a.aspec || a.spec_frames) {
      /* t41rx_row_byproducts_kernel */
      c.row = 1;
      c.rows_only = 1;
      for (int i = 0; i < kSmemFloats; ++i) smem[i] = NAN;
      for (int tid = 0; tid < kNT; ++tid) PhCtaInit(c, tid);
      for (int r = 0; r < a.n_rows; ++r) {
        c.t = r * row_every;
        c.row_idx = r;
        if (a.aspec) {
          T41RX_AUDIO_SPEC_SCHEDULE(EMUL_PHASE)
        }
        if (a.spec_frames) {
          T41RX_SPEC_FRAME_SCHEDULE(EMUL_PHASE)
        }
      }
    }
#undef EMUL_PHASE
  }
  return 0;
}

void emul_bind_audio_spectrum(Emul *e, int32_t *audio_ypixel, float *audio_max_sq_ave) {
  e->bind_ypixel = audio_ypixel;
  e->bind_max_ave = audio_max_sq_ave;
}

void emul_bind_control_frames(Emul *e, uint8_t *spec_frames, uint8_t *audio_frames) {
  e->bind_spec_frames = spec_frames;
  e->bind_audio_frames = audio_frames;
}

int emul_get_debug(Emul *e, int stream, t41rx_debug *d) {
  const StreamState &st = e->state[stream];
  memset(d, 0, sizeof(*d));
  d->agc_state = st.agc_state;
  d->agc_decay_type = st.agc_decay_type;
  d->agc_hang_counter = st.agc_hang_counter;
  d->agc_action = st.agc_action;
  d->rf_gain = st.rf_gain;
  d->codec_timer = (int32_t)st.codec_timer;
  d->zoom_sample_ptr = st.zoom_ptr;
  d->first_block = st.first_block;
  d->agc_volts = st.agc_volts;
  d->agc_ring_max = st.agc_ring_max;
  d->agc_save_volts = st.agc_save_volts;
  d->agc_fast_backaverage = st.agc_fast_back;
  d->agc_hang_backaverage = st.agc_hang_back;
  d->sam_phzerror = st.sam_phzerror;
  d->sam_omega2 = st.sam_omega2;
  d->sam_fil_out = st.sam_fil_out;
  d->dc_state[0] = st.dc_d1;
  d->dc_state[1] = st.dc_d2;
  d->am_wold = st.am_wold;
  if (st.nco_closed) {
    const double r = sqrt(st.osc_q * st.osc_q + st.osc_i * st.osc_i);
    d->osc_vect_q = r * cos(st.nco_phase);
    d->osc_vect_i = r * sin(st.nco_phase);
  } else {
    d->osc_vect_q = st.osc_q;
    d->osc_vect_i = st.osc_i;
  }
  return 0;
}

int emul_nco_closed(Emul *e, int stream) { return e->state[stream].nco_closed; }

}  // extern "C"
