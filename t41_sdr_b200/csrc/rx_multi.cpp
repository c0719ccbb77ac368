/*
 * rx_multi.cpp — a bank of receivers over several CUDA devices behind the C-ABI (include/t41rx.h, t41rx_multi_*).
 *
 * receivers are independent, so GPU g owns the contiguous range [g S / N, (g + 1) S / N) in a
 * single-device context of its own (rx_api.cu) and nothing crosses GPUs on the hot path.  Every call fans out to one
 * host thread per device (each drives its context's streams) and joins them; the only inter-GPU traffic is the
 * OPTIONAL gather of spectrum / waterfall rows to the first device, over NCCL (NVLink / NVSwitch) when libnccl can be
 * loaded, else over cudaMemcpyPeer.  There is no CPU processing path here either.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/t41rx.h"

namespace {

constexpr size_t kBlockSamples = T41RX_BLOCK_SAMPLES;
constexpr size_t kRes = T41RX_SPECTRUM_RES;

void SetLastError(const std::string &s);

/* the handful of NCCL entry points the gather needs, bound at run time (the library must load without NCCL) */
struct Nccl {
  void *h = nullptr;
  int (*CommInitAll)(void **, int, const int *) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  bool ok = false;
  void Load() {
    if (h) return;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) return;
    CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
    Send = (decltype(Send))dlsym(h, "ncclSend");
    Recv = (decltype(Recv))dlsym(h, "ncclRecv");
    ok = CommInitAll && CommDestroy && GroupStart && GroupEnd && Send && Recv;
  }
};
constexpr int kNcclInt8 = 0;   /* ncclInt8 / ncclChar */

}  // namespace

struct t41rx_multi {
  int n_streams = 0;
  std::vector<int> devices;
  std::vector<int> first, count;          /* shard g: receivers [first[g], first[g] + count[g]) */
  std::vector<t41rx_ctx *> ctx;
  /* row gather */
  Nccl nccl;
  std::vector<void *> comms;              /* one communicator per device, or empty: peer copies */
  std::vector<cudaStream_t> gstreams;
  bool distinct_devices = true;
};

namespace {

thread_local std::string g_multi_error;
void SetLastError(const std::string &s) { g_multi_error = s; }

int Fail(int code, const std::string &what) {
  SetLastError(what);
  return code;
}

/* run fn(g) on one host thread per device; first failure wins, its text becomes this thread's last error */
template <class F>
int FanOut(t41rx_multi *m, F fn) {
  const int n = (int)m->ctx.size();
  std::vector<int> rc(n, 0);
  std::vector<std::string> err(n);
  auto body = [&](int g) {
    rc[g] = fn(g);
    if (rc[g]) err[g] = t41rx_last_error();      /* the worker thread's own last-error slot */
  };
  if (n == 1) {
    body(0);
  } else {
    std::vector<std::thread> th;
    th.reserve(n);
    for (int g = 0; g < n; ++g) th.emplace_back(body, g);
    for (auto &t : th) t.join();
  }
  for (int g = 0; g < n; ++g)
    if (rc[g]) {
      char buf[64];
      snprintf(buf, sizeof(buf), "device %d (shard %d): ", m->devices[g], g);
      return Fail(rc[g], std::string(buf) + err[g]);
    }
  return T41RX_OK;
}

}  // namespace

extern "C" {

const char *t41rx_multi_last_error(void) { return g_multi_error.c_str(); }

void t41rx_destroy_multi(t41rx_multi *m) {
  if (!m) return;
  for (size_t g = 0; g < m->comms.size(); ++g)
    if (m->comms[g]) m->nccl.CommDestroy(m->comms[g]);
  for (size_t g = 0; g < m->gstreams.size(); ++g)
    if (m->gstreams[g]) {
      cudaSetDevice(m->devices[g]);
      cudaStreamDestroy(m->gstreams[g]);
    }
  for (t41rx_ctx *c : m->ctx) t41rx_destroy(c);
  delete m;
}

int t41rx_create_multi(t41rx_multi **out, int n_streams, const int *device_ids, int n_dev) {
  if (!out || n_streams <= 0 || !device_ids || n_dev <= 0 || n_dev > n_streams)
    return Fail(T41RX_EINVAL, "t41rx_create_multi: bad arguments");
  *out = nullptr;
  t41rx_multi *m = new t41rx_multi();
  m->n_streams = n_streams;
  m->devices.assign(device_ids, device_ids + n_dev);
  for (int g = 0; g < n_dev; ++g)
    for (int k = 0; k < g; ++k)
      if (device_ids[g] == device_ids[k]) m->distinct_devices = false;
  m->first.resize(n_dev);
  m->count.resize(n_dev);
  m->ctx.assign(n_dev, nullptr);
  for (int g = 0; g < n_dev; ++g) {
    m->first[g] = (int)((long long)n_streams * g / n_dev);
    m->count[g] = (int)((long long)n_streams * (g + 1) / n_dev) - m->first[g];
  }
  const int rc = FanOut(m, [&](int g) { return t41rx_create(&m->ctx[g], m->count[g], m->devices[g]); });
  if (rc) {
    const std::string keep = g_multi_error;
    t41rx_destroy_multi(m);
    return Fail(rc, keep);
  }
  *out = m;
  return T41RX_OK;
}

int t41rx_multi_num_devices(const t41rx_multi *m) { return m ? (int)m->ctx.size() : 0; }
int t41rx_multi_num_streams(const t41rx_multi *m) { return m ? m->n_streams : 0; }

int t41rx_multi_shard(const t41rx_multi *m, int shard, int *device, int *first, int *count, t41rx_ctx **ctx) {
  if (!m || shard < 0 || shard >= (int)m->ctx.size()) return Fail(T41RX_EINVAL, "t41rx_multi_shard: bad arguments");
  if (device) *device = m->devices[shard];
  if (first) *first = m->first[shard];
  if (count) *count = m->count[shard];
  if (ctx) *ctx = m->ctx[shard];
  return T41RX_OK;
}

static int SetParamsMulti(t41rx_multi *m, int first, int count, const t41rx_params *p, bool each) {
  if (!m || !p || first < 0 || count <= 0 || first + count > m->n_streams)
    return Fail(T41RX_EINVAL, "t41rx_multi_set_params: bad range");
  return FanOut(m, [&](int g) {
    const int lo = std::max(first, m->first[g]), hi = std::min(first + count, m->first[g] + m->count[g]);
    if (lo >= hi) return (int)T41RX_OK;
    return each ? t41rx_set_params_each(m->ctx[g], lo - m->first[g], hi - lo, p + (lo - first))
                : t41rx_set_params(m->ctx[g], lo - m->first[g], hi - lo, p);
  });
}
int t41rx_multi_set_params(t41rx_multi *m, int first, int count, const t41rx_params *p) {
  return SetParamsMulti(m, first, count, p, false);
}
int t41rx_multi_set_params_each(t41rx_multi *m, int first, int count, const t41rx_params *p) {
  return SetParamsMulti(m, first, count, p, true);
}

int t41rx_multi_get_debug(t41rx_multi *m, int stream, t41rx_debug *d) {
  if (!m || stream < 0 || stream >= m->n_streams) return Fail(T41RX_EINVAL, "t41rx_multi_get_debug: bad arguments");
  for (size_t g = 0; g < m->ctx.size(); ++g)
    if (stream < m->first[g] + m->count[g]) {
      const int rc = t41rx_get_debug(m->ctx[g], stream - m->first[g], d);
      if (rc) SetLastError(t41rx_last_error());
      return rc;
    }
  return T41RX_EINVAL;
}

/* host buffers of the WHOLE bank ([n_streams][...] layouts of t41rx.h): every device works on its slice */
static int ProcessMulti(t41rx_multi *m, const void *iq, void *audio, bool q15, int n_blocks, int row_every,
                        int16_t *spec_rows, uint16_t *wf_rows, int8_t *psk_bits, uint8_t *psk_chars, uint32_t flags) {
  if (!m || !iq || !audio || n_blocks <= 0 || row_every < 0) return Fail(T41RX_EINVAL, "t41rx_multi_process: bad arguments");
  const size_t T = (size_t)n_blocks;
  const size_t n_rows = row_every > 0 ? (T + row_every - 1) / row_every : 0;
  return FanOut(m, [&](int g) {
    const size_t s0 = (size_t)m->first[g];
    int16_t *sp = spec_rows ? spec_rows + s0 * n_rows * kRes : nullptr;
    uint16_t *wf = wf_rows ? wf_rows + s0 * n_rows * kRes : nullptr;
    int8_t *pb = psk_bits ? psk_bits + s0 * T : nullptr;
    uint8_t *pc = psk_chars ? psk_chars + s0 * T : nullptr;
    if (q15)
      return t41rx_process_q15(m->ctx[g], (const int16_t *)iq + s0 * T * 2 * kBlockSamples,
                               (int16_t *)audio + s0 * T * kBlockSamples, n_blocks, row_every, sp, wf, pb, pc, flags);
    return t41rx_process(m->ctx[g], (const float *)iq + s0 * T * 2 * kBlockSamples, (float *)audio + s0 * T * kBlockSamples,
                         n_blocks, row_every, sp, wf, pb, pc, flags);
  });
}
int t41rx_multi_process(t41rx_multi *m, const float *iq, float *audio, int n_blocks, int row_every, int16_t *spec_rows,
                        uint16_t *wf_rows, int8_t *psk_bits, uint8_t *psk_chars, uint32_t flags) {
  return ProcessMulti(m, iq, audio, false, n_blocks, row_every, spec_rows, wf_rows, psk_bits, psk_chars, flags);
}
int t41rx_multi_process_q15(t41rx_multi *m, const int16_t *iq_q15, int16_t *audio_q15, int n_blocks, int row_every,
                            int16_t *spec_rows, uint16_t *wf_rows, int8_t *psk_bits, uint8_t *psk_chars, uint32_t flags) {
  return ProcessMulti(m, iq_q15, audio_q15, true, n_blocks, row_every, spec_rows, wf_rows, psk_bits, psk_chars, flags);
}

/* device-resident: per-shard DEVICE pointers (each on its shard's device, [count[g]][...] layouts); asynchronous */
int t41rx_multi_process_device(t41rx_multi *m, const float *const *iq, float *const *audio, int n_blocks, int row_every,
                               int16_t *const *spec_rows, uint16_t *const *wf_rows, uint32_t flags) {
  if (!m || !iq || !audio || n_blocks <= 0 || row_every < 0)
    return Fail(T41RX_EINVAL, "t41rx_multi_process_device: bad arguments");
  return FanOut(m, [&](int g) {
    return t41rx_process_device(m->ctx[g], iq[g], audio[g], n_blocks, row_every, spec_rows ? spec_rows[g] : nullptr,
                                wf_rows ? wf_rows[g] : nullptr, nullptr, nullptr, flags, nullptr);
  });
}

int t41rx_multi_synchronize(t41rx_multi *m) {
  if (!m) return Fail(T41RX_EINVAL, "t41rx_multi_synchronize: null context");
  return FanOut(m, [&](int g) { return t41rx_synchronize(m->ctx[g]); });
}

/* Optional: gather the shards' row buffers (device pointers, bytes_per_receiver bytes per receiver, e.g.
   n_rows * 512 * 2 for spec_rows) into dst on the FIRST device, in receiver order.  NCCL send / recv over NVLink when
   libnccl loads and the devices are distinct, else cudaMemcpyPeerAsync.  Blocking.  *used_nccl (may be NULL) tells
   which path ran. */
int t41rx_multi_gather_rows(t41rx_multi *m, const void *const *rows, size_t bytes_per_receiver, void *dst, int *used_nccl) {
  if (!m || !rows || !dst || bytes_per_receiver == 0) return Fail(T41RX_EINVAL, "t41rx_multi_gather_rows: bad arguments");
  const int n = (int)m->ctx.size();
  int rc = t41rx_multi_synchronize(m);     /* the rows must be complete */
  if (rc) return rc;
  if (m->gstreams.empty()) {
    m->gstreams.assign(n, nullptr);
    for (int g = 0; g < n; ++g) {
      if (cudaSetDevice(m->devices[g]) != cudaSuccess ||
          cudaStreamCreateWithFlags(&m->gstreams[g], cudaStreamNonBlocking) != cudaSuccess)
        return Fail(T41RX_ECUDA, "t41rx_multi_gather_rows: stream creation failed");
    }
  }
  bool use_nccl = false;
  if (n > 1 && m->distinct_devices) {
    m->nccl.Load();
    if (m->nccl.ok && m->comms.empty()) {
      m->comms.assign(n, nullptr);
      if (m->nccl.CommInitAll(m->comms.data(), n, m->devices.data()) != 0) m->comms.clear();
    }
    use_nccl = m->nccl.ok && !m->comms.empty();
  }
  if (used_nccl) *used_nccl = use_nccl ? 1 : 0;
  char *d = (char *)dst;
  if (use_nccl) {
    if (m->nccl.GroupStart() != 0) return Fail(T41RX_ECUDA, "t41rx_multi_gather_rows: ncclGroupStart failed");
    for (int g = 0; g < n; ++g) {
      const size_t bytes = (size_t)m->count[g] * bytes_per_receiver;
      /* rank 0 receives every shard (its own included, as a send to itself inside the group) */
      if (m->nccl.Send(rows[g], bytes, kNcclInt8, 0, m->comms[g], m->gstreams[g]) != 0 ||
          m->nccl.Recv(d + (size_t)m->first[g] * bytes_per_receiver, bytes, kNcclInt8, g, m->comms[0], m->gstreams[0]) != 0)
        return Fail(T41RX_ECUDA, "t41rx_multi_gather_rows: ncclSend / ncclRecv failed");
    }
    if (m->nccl.GroupEnd() != 0) return Fail(T41RX_ECUDA, "t41rx_multi_gather_rows: ncclGroupEnd failed");
  } else {
    for (int g = 0; g < n; ++g) {
      const size_t bytes = (size_t)m->count[g] * bytes_per_receiver;
      if (cudaSetDevice(m->devices[0]) != cudaSuccess ||
          cudaMemcpyPeerAsync(d + (size_t)m->first[g] * bytes_per_receiver, m->devices[0], rows[g], m->devices[g], bytes,
                              m->gstreams[0]) != cudaSuccess)
        return Fail(T41RX_ECUDA, "t41rx_multi_gather_rows: peer copy failed");
    }
  }
  for (int g = 0; g < n; ++g) {
    if (cudaSetDevice(m->devices[g]) != cudaSuccess || cudaStreamSynchronize(m->gstreams[g]) != cudaSuccess)
      return Fail(T41RX_ECUDA, "t41rx_multi_gather_rows: synchronisation failed");
  }
  return T41RX_OK;
}

}  // extern "C"
