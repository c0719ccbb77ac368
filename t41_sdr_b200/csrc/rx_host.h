/*
 * rx_host.h — CUDA-free host model of a t41rx context: the per-receiver parameter cache and
 * every table the kernel reads, kept in host vectors.  rx_api.cu mirrors these vectors into
 * HBM; tests/devtools drives the same model when it single-steps the kernel phases on the
 * CPU.  The semantics of a parameter change follow the firmware's control path
 * (SURVEY.md §3.3): filters are re-designed when mode or cut-offs change (CalcFilters),
 * AGC constants when the AGC mode or threshold change (AGCLoadValues, sticky hang_thresh),
 * the zoom filters and ring pointer when the zoom changes (ZoomFFTPrep), and the oscillator
 * rotation when NCOFreq changes (FreqShift2 samples it per block).
 */
#ifndef T41RX_HOST_H
#define T41RX_HOST_H

#include <map>
#include <tuple>
#include <vector>

#include "rx_design.h"
#include "rx_types.h"

namespace t41rx {

struct StatePatch {
  bool set_rf_gain = false;
  int32_t rf_gain = 0;
  bool reset_zoom_ptr = false;
  bool clear_fast_native = false;   /* rfGainAllBands changed: the throughput kernel's DC-block form is stale */
};

struct HostModel {
  int n_streams = 0;
  std::vector<t41rx_params> params;
  std::vector<int> last_set_rf_gain;
  std::vector<AgcSticky> sticky;
  std::vector<AgcConsts> agc;
  std::vector<int> attack_buffsize;
  std::vector<StreamCfg> cfg;
  std::vector<double> nco_tab;          /* 192 per receiver: W[64], C[32] as (cos, sin) */
  std::vector<FilterSet> fsets;
  std::map<std::tuple<int, int, int, int>, int> fset_ids;
  std::vector<int> fset_refs;           /* receivers on each filter set: unreferenced slots are re-used */

  /* constant tables */
  std::vector<float> twiddle;           /* 512 (cos, sin) */
  std::vector<double> hann;             /* 512 */
  std::vector<float> sin_table;         /* 513 */
  std::vector<float> zoom_iir;          /* 80 */
  std::vector<float> eq_coeffs;         /* 14 x 20 */
  std::vector<float> cw_coeffs;         /* 5 x 30 */
  std::vector<float> sam_consts;        /* 4 */
  std::vector<float> nr_tab;            /* 520: constants, Kim's Hann window, sqrtHann (rx_nr.cuh) */
  std::vector<uint16_t> gradient;       /* 117 */
  std::vector<uint32_t> varicode;       /* 128 */

  void Init(int n);
  /* returns 0 or T41RX_EINVAL; *new_fset is the id of a newly designed filter set or -1 */
  int Apply(int s, const t41rx_params &p, StatePatch *patch, int *new_fset);
};

void HostStateInit(StreamState *st);
void DefaultParams(t41rx_params *p);

}  // namespace t41rx
#endif
