/*
 * rx_design.h — host-side control path of libt41rx: turns receiver parameters into the
 * coefficient tables and scalar constants the fused CUDA kernel consumes.
 *
 * Mirrors what the firmware recomputes on a parameter change, never per block
 * (SURVEY.md §3.3): CalcFilters / InitFilterMask / SetDecIntFilters (Filter.cpp:235-438),
 * CalcFIRCoeffs / CalcCplxFIRCoeffs / SetIIRCoeffs (FIR.cpp:908-1116), AGCLoadValues
 * (DSP_Fn.cpp:368-468), ZoomFFTPrep (FFT.cpp:35-55) and the per-block scalar set-up at the
 * top of FreqShift2 (Freq_Shift.cpp:121-124) and ProcessIQData (Process.cpp:117,482-490,929).
 * Must be compiled with -ffp-contract=off: the float/double promotion pattern of each
 * expression is part of the contract with the oracle.
 */
#ifndef T41RX_DESIGN_H
#define T41RX_DESIGN_H

#include <stdint.h>

#include "../../include/t41rx.h"
#include "rx_types.h"

namespace t41rx {

/* 512-point radix-8 FFT on the host, same butterfly schedule as the device FFT; used once
 * per filter change to turn the 257 complex taps into the frequency-domain mask. */
void HostFft512(float *interleaved);

void DesignKaiserLowpass(float *taps, int n_taps, float cutoff_hz, float stop_db, float fs_hz);
void DesignComplexBandpass(float *taps_re, float *taps_im, int n_taps, float lo_hz, float hi_hz, float fs_hz);
void DesignAmLowpass(float *coeffs5);

/* Sticky AGC tuning values that survive AGC mode changes (DSP_Fn.cpp:378-402, B12). */
struct AgcSticky {
  float hangtime;
  float tau_decay;
  float hang_thresh;
};
void AgcStickyDefaults(AgcSticky *s);
void DesignAgc(AgcSticky *sticky, int agc_mode, int agc_thresh, AgcConsts *out, int *attack_buffsize);

/* Filter tables of one (mode, cuts, nfm bandwidth) combination. */
void DesignFilterSet(const t41rx_params &p, FilterSet *fs);
/* Scalars and small tables of one receiver. */
void DesignStreamCfg(const t41rx_params &p, const AgcConsts &agc, int filter_id, StreamCfg *cfg);

int ValidateParams(const t41rx_params &p);

}  // namespace t41rx
#endif
