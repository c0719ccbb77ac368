/* rx_host.cpp — see rx_host.h.  Compile with -ffp-contract=off. */
#include "rx_host.h"

#include <math.h>
#include <string.h>

#include "rx_tables_data.h"

namespace t41rx {

const float *HostTwiddle512();

void DefaultParams(t41rx_params *p) {
  memset(p, 0, sizeof(*p));
  p->mode = T41RX_DEMOD_USB;      /* bands[] 20M..10M, T41_SDR.ino:163-167 */
  p->f_lo_cut = 200;
  p->f_hi_cut = 3000;
  p->nco_freq = 0;                /* T41_SDR.ino:793 */
  p->agc_mode = 1;                /* gwv.cpp:15 */
  p->agc_thresh = 20;
  p->audio_volume = 30;           /* gwv.cpp:16 */
  p->rf_gain_all_bands = 1;       /* gwv.cpp:17 */
  p->rf_gain = 1;
  p->spectrum_zoom = 1;           /* gwv.cpp:25 */
  p->current_scale = 1;           /* gwv.cpp:24 */
  p->pixel_offset = 20;
  p->current_nf = 0;
  p->spectrum_noise_floor = 247;  /* gwv.cpp:18 */
  p->nfm_filter_bw = 12000;       /* Filter.cpp:16 */
  p->psk31_enable = 0;
  p->iq_amp_correction = 1.0f;    /* gwv.cpp:70 */
  p->iq_phase_correction = 0.0f;  /* gwv.cpp:71 */
  p->receive_eq_flag = 0;         /* OFF */
  for (int i = 0; i < 14; ++i) p->equalizer_rec[i] = 100;   /* EEPROM.cpp:59,698 */
  p->nr_option = 0;               /* nrOptionSelect */
  p->anr_notch_on = 0;            /* ANR_notchOn */
  p->cw_receive = 0;              /* T41State = SSB_RECEIVE */
  p->cw_filter_index = 5;         /* CWFilterIndex: off (gwv.cpp) */
  p->nb_on = 0;                   /* NB_on (Process.cpp:39) */
}

void HostStateInit(StreamState *st) {
  memset(st, 0, sizeof(*st));
  st->osc_q = 1.0;           /* Freq_Shift.cpp:13 */
  st->osc_i = 0.0;
  st->first_block = 1;       /* Process.cpp:47 */
  st->rf_gain = 1;           /* bands[].RFgain, T41_SDR.ino:149-167 */
  st->agc_pos = kAgcRing - 1;
  st->anr_lidx = 120.0f;     /* Noise.cpp:47 */
  st->anr_ngamma = 0.001f;   /* Noise.cpp:52 */
}

static void BuildNcoTable(const StreamCfg &c, double *tab /*192*/) {
  const long double d = atan2l((long double)c.osc_sin, (long double)c.osc_cos);
  for (int k = 0; k < 64; ++k) {
    tab[2 * k] = (double)cosl(d * k);
    tab[2 * k + 1] = (double)sinl(d * k);
  }
  for (int m = 0; m < 32; ++m) {
    tab[128 + 2 * m] = (double)cosl(d * (64 * m));
    tab[128 + 2 * m + 1] = (double)sinl(d * (64 * m));
  }
}

void HostModel::Init(int n) {
  n_streams = n;
  const float *tw = HostTwiddle512();
  twiddle.assign(tw, tw + 1024);
  hann.resize(512);
  for (int i = 0; i < 512; ++i) hann[i] = (0.5 - 0.5 * cos(6.28 * i / 512));   /* FFT.cpp:110,221 */
  sin_table.resize(513);   /* arm_sin_f32's table: sin(2*pi*k/512) rounded to float */
  for (int k = 0; k <= 512; ++k) sin_table[k] = (float)sin(2.0 * 3.14159265358979323846 * (double)k / 512.0);
  sin_table[0] = 0.0f;
  sin_table[256] = 0.0f;
  sin_table[512] = 0.0f;
  zoom_iir.assign(&t41rx_zoom_iir[0][0], &t41rx_zoom_iir[0][0] + 80);
  eq_coeffs.assign(&t41rx_eq_coeffs[0][0], &t41rx_eq_coeffs[0][0] + 280);
  cw_coeffs.assign(&t41rx_cw_coeffs[0][0], &t41rx_cw_coeffs[0][0] + 150);
  {
    /* SAM PLL constants, Demod.cpp:13-18 with omegaN = 200, pll_fmax = 4000 (gwv.cpp:64-65);
       exp() of a float argument is the single-precision overload under ISO C++ */
    const float tpi = 6.283185307179586476925286766559f;
    const float omegaN = 200.0;
    const float pll_fmax = +4000.0;
    const int zeta_help = 65;
    const float zeta = (float)zeta_help / 100.0;
    sam_consts.resize(4);
    sam_consts[0] = tpi * -pll_fmax * 1 / 24000;
    sam_consts[1] = tpi * pll_fmax * 1 / 24000;
    sam_consts[2] = 1.0 - exp(-2.0 * omegaN * zeta * 1 / 24000);
    sam_consts[3] = -sam_consts[2] +
                    2.0 * (1 - expf(-omegaN * zeta * 1 / 24000) * cosf(omegaN * 1 / 24000 * sqrtf(1.0 - zeta * zeta)));
  }
  {
    /* tables of the spectral noise-reduction stages (rx_nr.cuh): the reference's own expressions, evaluated once with
       the host's libm like the reference build evaluates them on every call */
    nr_tab.assign(520, 0.0f);
    const float tinc = 0.00533333, tax = 0.0239, tap = 0.05062, asnr = 20, pspri = 0.5;      /* Noise.cpp:396-403 */
    const float xih1 = powf(10, (float)asnr / 10.0);
    nr_tab[0] = expf(-tinc / tax);                                   /* ax */
    nr_tab[1] = expf(-tinc / tap);                                   /* ap */
    nr_tab[2] = 1.0 / (1.0 + xih1) - 1.0;                            /* xih1r */
    nr_tab[3] = (1.0 / pspri - 1.0) * (1.0 + xih1);                  /* pfac */
    nr_tab[4] = powf(10, -(float)20 / 20.0);                         /* snr_prio_min */
    for (int idx = 0; idx < 256; ++idx)                              /* Noise.cpp:198-201 (PI is Arduino's double there) */
      nr_tab[8 + idx] = 0.5 * (float)(1.0 - (cosf(3.1415926535897932384626433832795 * 2.0 * (float)idx / (float)((256) - 1))));
    for (int idx = 0; idx < 256; ++idx) nr_tab[264 + idx] = t41rx_sqrt_hann[idx];
  }
  gradient.assign(t41rx_gradient, t41rx_gradient + 117);
  varicode.resize(128);
  for (int i = 0; i < 128; ++i)
    varicode[i] = (uint32_t)t41rx_varicode[i].code | ((uint32_t)t41rx_varicode[i].bits << 16) |
                  ((uint32_t)t41rx_varicode[i].ascii << 24);

  params.resize(n);
  last_set_rf_gain.assign(n, 1);
  sticky.resize(n);
  agc.resize(n);
  attack_buffsize.assign(n, 0);
  cfg.resize(n);
  memset(cfg.data(), 0, sizeof(StreamCfg) * n);
  nco_tab.assign((size_t)192 * n, 0.0);

  /* start-up state of every receiver: AGCPrep + AGCLoadValues for the default AGC mode,
     CalcFilters for the default band (T41_SDR.ino:473-667,753-795) */
  t41rx_params def;
  DefaultParams(&def);
  AgcStickyDefaults(&sticky[0]);
  DesignAgc(&sticky[0], def.agc_mode, def.agc_thresh, &agc[0], &attack_buffsize[0]);
  fsets.emplace_back();
  DesignFilterSet(def, &fsets[0]);
  fset_ids[std::make_tuple(0, 0, def.f_lo_cut, def.f_hi_cut)] = 0;
  fset_refs.assign(1, n);
  DesignStreamCfg(def, agc[0], 0, &cfg[0]);
  BuildNcoTable(cfg[0], &nco_tab[0]);
  for (int s = 0; s < n; ++s) {
    params[s] = def;
    sticky[s] = sticky[0];
    agc[s] = agc[0];
    attack_buffsize[s] = attack_buffsize[0];
    cfg[s] = cfg[0];
    memcpy(&nco_tab[(size_t)192 * s], &nco_tab[0], sizeof(double) * 192);
  }
}

int HostModel::Apply(int s, const t41rx_params &p, StatePatch *patch, int *new_fset) {
  *new_fset = -1;
  if (!ValidateParams(p)) return T41RX_EINVAL;
  const t41rx_params old = params[s];
  params[s] = p;
  if (p.agc_mode != old.agc_mode || p.agc_thresh != old.agc_thresh)
    DesignAgc(&sticky[s], p.agc_mode, p.agc_thresh, &agc[s], &attack_buffsize[s]);
  const bool nfm = p.mode == T41RX_DEMOD_NFM;
  const auto key = std::make_tuple(nfm ? 1 : 0, nfm ? p.nfm_filter_bw : 0, p.f_lo_cut, p.f_hi_cut);
  int fid;
  const int old_fid = cfg[s].filter_id;
  auto it = fset_ids.find(key);
  if (it != fset_ids.end()) {
    fid = it->second;
  } else {
    /* a slot no receiver references any more is re-used (a tuning sweep on a long-lived context would otherwise grow
       the table without bound); this receiver's own old slot counts once it is the only one on it */
    fid = -1;
    for (int i = 0; i < (int)fsets.size() && fid < 0; ++i)
      if (fset_refs[i] == 0 || (i == old_fid && fset_refs[i] == 1)) fid = i;
    if (fid >= 0) {
      for (auto e = fset_ids.begin(); e != fset_ids.end(); ++e)
        if (e->second == fid) { fset_ids.erase(e); break; }
    } else {
      fid = (int)fsets.size();
      fsets.emplace_back();
      fset_refs.push_back(0);
    }
    DesignFilterSet(p, &fsets[fid]);
    fset_ids[key] = fid;
    *new_fset = fid;
  }
  fset_refs[old_fid] -= 1;
  fset_refs[fid] += 1;
  DesignStreamCfg(p, agc[s], fid, &cfg[s]);
  if (p.nco_freq != old.nco_freq) {
    cfg[s].nco_epoch += 1;
    BuildNcoTable(cfg[s], &nco_tab[(size_t)192 * s]);
  }
  if (p.rf_gain != last_set_rf_gain[s]) {
    last_set_rf_gain[s] = p.rf_gain;
    patch->set_rf_gain = true;
    patch->rf_gain = p.rf_gain;
  }
  if (p.spectrum_zoom != old.spectrum_zoom) patch->reset_zoom_ptr = true;   /* FFT.cpp:54 */
  if (p.rf_gain_all_bands != old.rf_gain_all_bands) patch->clear_fast_native = true;
  return 0;
}

}  // namespace t41rx
