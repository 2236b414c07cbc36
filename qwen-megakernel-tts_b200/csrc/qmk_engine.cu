// qmk_engine.cu — host side of the C ABI declared in include/qmk_b200.h.
//
// Per-device engine context (exchange buffers, watchdog word, epoch counter), weight re-packing into
// per-CTA streams, launch of the persistent decode kernel, and the upstream-compatible symbol
// launch_ldg_decode_direct (upstream csrc/kernel.cu:1485-1513).  No torch types anywhere.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "qmk_b200.h"
#include "qmk_device.cuh"
#include "qmk_device2.cuh"

using namespace qmk;

static thread_local std::string g_last_error;

static int set_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define QMK_CUDA(expr)                                                                           \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return set_error(QMK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct qmk_head {
  uint8_t* packed = nullptr;
  int rows = 0;
  int segs_max = 0;
};

struct qmk_engine {
  int device = 0;
  int G = 0;
  int version = 1;      // 1 = row-split kernel (qmk_device.cuh), 2 = group kernel (qmk_device2.cuh)
  size_t xbuf_bytes = 0;
  std::vector<int> roles;   // group kernel: blockIdx -> role
  size_t xbuf_ll_off = 0;   // start of the epoch-tagged words (the part that is cleared when the epoch wraps)
  uint8_t* xbuf = nullptr;
  float* res_spill = nullptr;
  int* delays = nullptr;
  int delay_o_idle = 4500;
  int o_sentinel = 0;   // measured: the extra hop costs more than the avoided early polls
  int coop = 1;   // cooperative launch = co-residency of all CTAs is checked by the driver
  long long* trace_dev = nullptr;
  int trace_stride = 0;
  int* status_dev = nullptr;
  uint32_t epoch = 0;
  long long timeout_cycles = 4000000000LL;  // ~2 s at 1.9 GHz
  std::mutex mu;
  // Stream hand-over: all launches of an engine share one set of exchange words, cumulative accumulator totals and one epoch
  // counter, so they must execute in submission order.  The engine remembers the stream of its latest launch; a launch on a
  // different stream first waits (event) for everything submitted to the previous one.
  cudaStream_t last_stream = nullptr;
  bool has_last_stream = false;
  cudaEvent_t handover = nullptr;
  bool tuned = false;       // group kernel: the groups' exchange buffers have been placed by the in-situ autotuner
  std::vector<int> slots;   // current pool slots of the groups' q/k/v and m buffers
};

struct qmk_model {
  qmk_engine* e = nullptr;
  Layout lay{};
  uint8_t* packed_layers = nullptr;
    uint8_t* aux_layers = nullptr;  // [L][2][AUX_BYTES]: input_ln | q_norm | k_norm  and  post_ln
  uint8_t* norm_seg = nullptr;    // AUX_BYTES: final RMSNorm weight (aux block of the head phase)
  int residual_fp32 = 1;
  int64_t packed_bytes = 0;
  std::vector<qmk_head> heads;
  const void* group_embed[QMK_CP_GROUPS] = {nullptr};
  unsigned long long rope_axis[2] = {0ull, 0ull};   // M-RoPE axis of every rotary frequency (2 bits each); 0 = standard RoPE
};

extern "C" int qmk_abi_version(void) { return QMK_ABI_VERSION; }
#ifndef QMK_SRC_HASH
#define QMK_SRC_HASH "unknown"
#endif
// The marker makes the hash readable from the file without loading the library (build_tts.library_hash).
static const char g_src_hash_marker[] = "QMK_SRC_HASH:" QMK_SRC_HASH;
extern "C" const char* qmk_source_hash(void) { return g_src_hash_marker + 13; }
extern "C" const char* qmk_last_error(void) { return g_last_error.c_str(); }

static Layout make_layout(int G, int L) {
  Layout y;
  y.G = G;
  y.L = L;
  y.qkv_max = (QKV_ROWS + G - 1) / G;
  y.o_max = (H + G - 1) / G;
  y.gu_max = (INTER + G - 1) / G;
  y.off_o = y.qkv_max;
  y.off_gu = y.off_o + 2 * y.o_max;
  y.off_down = y.off_gu + 2 * y.gu_max;
  y.layer_segs = y.off_down + 3 * y.o_max;
  return y;
}

extern "C" int qmk_engine_create(int device, int num_ctas, qmk_engine** out) {
  if (!out) return set_error(QMK_ERR_ARG, "qmk_engine_create: out is null");
  *out = nullptr;
  int ndev = 0;
  QMK_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return set_error(QMK_ERR_ARG, "qmk_engine_create: bad device %d", device);
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  QMK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 9)
    return set_error(QMK_ERR_UNSUPPORTED, "device %d is sm_%d%d; this engine needs TMA bulk copies (built for sm_100a)",
                     device, prop.major, prop.minor);
  int coop = 0;
  QMK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  if (!coop) return set_error(QMK_ERR_UNSUPPORTED, "device %d lacks cooperative launch", device);
  if (const char* env = getenv("QMK_NUM_CTAS")) {
    if (num_ctas <= 0) num_ctas = atoi(env);
  }
  // Kernel generation: 2 = group kernel (qmk_device2.cuh, 128 CTAs; default), 1 = row-split kernel (qmk_device.cuh, one CTA per
  // SM; also the only one with the staged debugging mode).  QMK_ENGINE overrides; an explicit CTA count other than 128
  // selects the row-split kernel.
  int version = prop.multiProcessorCount >= qmk2::G2 ? 2 : 1;
  if (const char* env = getenv("QMK_ENGINE")) version = atoi(env);
  if (version != 1 && version != 2) return set_error(QMK_ERR_ARG, "QMK_ENGINE=%d: expected 1 or 2", version);
  if (version == 2 && num_ctas > 0 && num_ctas != qmk2::G2) version = 1;
  int G = num_ctas > 0 ? num_ctas : prop.multiProcessorCount;
  if (G > prop.multiProcessorCount) G = prop.multiProcessorCount;
  if (version == 2) {
    if (prop.multiProcessorCount < qmk2::G2)
      return set_error(QMK_ERR_UNSUPPORTED, "the group kernel needs %d SMs, device %d has %d", qmk2::G2, device, prop.multiProcessorCount);
    G = qmk2::G2;
  }
  // every CTA must own >=1 row in every phase and a phase may span at most MAX_ST ring stages
  if (version == 1 && (G < 147 || G > 1024)) return set_error(QMK_ERR_UNSUPPORTED, "unsupported CTA count %d (need 147..1024: a phase spans at most 3 ring stages)", G);
  QMK_CUDA(cudaFuncSetAttribute(qmk_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  QMK_CUDA(cudaFuncSetAttribute(qmk_decode_kernel_traced, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  QMK_CUDA(cudaFuncSetAttribute(qmk2::qmk2_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, qmk2::SMEM2_BYTES));
  QMK_CUDA(cudaFuncSetAttribute(qmk2::qmk2_decode_kernel_traced, cudaFuncAttributeMaxDynamicSharedMemorySize, qmk2::SMEM2_BYTES));
  int occ = 0;
  if (version == 1) QMK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qmk_decode_kernel, NTHREADS, SMEM_BYTES));
  else QMK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qmk2::qmk2_decode_kernel, NTHREADS, qmk2::SMEM2_BYTES));
  if (occ < 1) return set_error(QMK_ERR_UNSUPPORTED, "decode kernel does not fit on an SM");

  qmk_engine* e = new qmk_engine();
  e->device = device;
  e->G = G;
  e->version = version;
  e->xbuf_bytes = version == 2 ? qmk2::XBUF2_BYTES : XBUF_BYTES;
  e->xbuf_ll_off = version == 2 ? qmk2::XB_LL : 0;
  if (const char* env = getenv("QMK_TIMEOUT_CYCLES")) e->timeout_cycles = atoll(env);
  int delay0 = version == 2 ? 500 : 800;
  int delay_token = 2500;   // the next step's token arrives two exchanges after a CTA has published its logits
  if (const char* env = getenv("QMK_POLL_DELAY")) delay0 = atoi(env);
  if (const char* env = getenv("QMK_POLL_DELAY_TOKEN")) delay_token = atoi(env);
  // attention CTAs: wait for q/k/v (DL_ATTN) and, in the O phase, for the other attention CTAs' output (DL_O)
  int delay_attn = version == 2 ? 200 : 400, delay_oa = 300;
  int delay_down = version == 2 ? 200 : delay0;
  if (const char* env = getenv("QMK_POLL_DELAY_DOWN")) delay_down = atoi(env);
  if (const char* env = getenv("QMK_POLL_DELAY_ATTN")) delay_attn = atoi(env);
  if (const char* env = getenv("QMK_POLL_DELAY_OA")) delay_oa = atoi(env);
  if (const char* env = getenv("QMK_POLL_DELAY_O")) e->delay_o_idle = atoi(env);
  if (const char* env = getenv("QMK_O_SENTINEL")) e->o_sentinel = atoi(env);
  if (const char* env = getenv("QMK_COOP")) e->coop = atoi(env);
  std::vector<int> delays((size_t)G * 3 * DL_N, 0);
  for (int c = 0; c < G; ++c)
    for (int d = 0; d < DL_N; ++d)
      delays[(size_t)c * 3 * DL_N + d] = (d == DL_TOKEN) ? delay_token : (d == DL_ATTN) ? delay_attn : (d == DL_O) ? delay_oa : (d == DL_DOWN) ? delay_down : delay0;
  cudaError_t err = cudaMalloc(&e->xbuf, e->xbuf_bytes);
  if (err == cudaSuccess) err = cudaMemset(e->xbuf, 0, e->xbuf_bytes);
  if (err == cudaSuccess) err = cudaMalloc(&e->res_spill, H * sizeof(float));
  if (err == cudaSuccess) err = cudaMemset(e->res_spill, 0, H * sizeof(float));
  if (err == cudaSuccess) err = cudaMalloc(&e->delays, delays.size() * sizeof(int));
  if (err == cudaSuccess) err = cudaMemcpy(e->delays, delays.data(), delays.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMalloc(&e->status_dev, 4 * sizeof(int));
  if (err == cudaSuccess) err = cudaMemset(e->status_dev, 0, 4 * sizeof(int));
  if (err == cudaSuccess && version == 2) {
    // CTA roles: identity (default), or (QMK_GROUP_BY_SM=1) sorted by the SM id a probe launch of the same shape observed, which
    // puts a group's 16 CTAs on neighbouring SMs.  Measured on B200: the block scheduler already hands consecutive CTA pairs
    // to different GPCs, and groups spread over the chip are FASTER than GPC-local ones (343 vs 348 us per talker step).
    std::vector<int> role(qmk2::G2);
    for (int i = 0; i < qmk2::G2; ++i) role[i] = i;
    int by_sm = 0;
    if (const char* env = getenv("QMK_GROUP_BY_SM")) by_sm = atoi(env);
    if (by_sm) {
      int* d_smid = nullptr;
      std::vector<int> smid(qmk2::G2, 0);
      cudaError_t pe = cudaMalloc(&d_smid, qmk2::G2 * sizeof(int));
      if (pe == cudaSuccess) pe = cudaFuncSetAttribute(qmk2::qmk2_probe_smid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, qmk2::SMEM2_BYTES);
      if (pe == cudaSuccess) {
        void* pargs[] = {&d_smid};
        pe = cudaLaunchCooperativeKernel((const void*)qmk2::qmk2_probe_smid_kernel, dim3(qmk2::G2), dim3(NTHREADS), pargs, qmk2::SMEM2_BYTES, 0);
      }
      if (pe == cudaSuccess) pe = cudaMemcpy(smid.data(), d_smid, qmk2::G2 * sizeof(int), cudaMemcpyDeviceToHost);
      if (d_smid) cudaFree(d_smid);
      if (pe == cudaSuccess) {
        std::vector<int> order(qmk2::G2);
        for (int i = 0; i < qmk2::G2; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return smid[a] < smid[b]; });
        for (int r = 0; r < qmk2::G2; ++r) role[order[r]] = r;
      } else {
        cudaGetLastError();
      }
    }
    err = cudaMemcpy(e->xbuf + qmk2::XB_ROLE, role.data(), qmk2::G2 * sizeof(int), cudaMemcpyHostToDevice);
    e->roles = role;
    // Placement of the groups' exchange buffers: default slots 2g, 2g+1; QMK_CALIBRATE=1 (default) measures the pool.
    std::vector<int> slots(qmk2::NGRP * 2);
    for (int g = 0; g < qmk2::NGRP; ++g) { slots[2 * g] = 2 * g; slots[2 * g + 1] = 2 * g + 1; }
    int calibrate = 1;
    if (const char* env = getenv("QMK_CALIBRATE")) calibrate = atoi(env);
    if (err == cudaSuccess && calibrate) {
      const int rounds = 40;
      unsigned* d_bar = nullptr;
      long long* d_out = nullptr;
      std::vector<long long> t((size_t)qmk2::NGRP * qmk2::XP_CAND, 0);
      cudaError_t ce = cudaMalloc(&d_bar, 2 * sizeof(unsigned));
      if (ce == cudaSuccess) ce = cudaMemset(d_bar, 0, 2 * sizeof(unsigned));
      if (ce == cudaSuccess) ce = cudaMalloc(&d_out, t.size() * sizeof(long long));
      if (ce == cudaSuccess) ce = cudaMemset(d_out, 0, t.size() * sizeof(long long));
      if (ce == cudaSuccess) ce = cudaFuncSetAttribute(qmk2::qmk2_calib_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, qmk2::SMEM2_BYTES);
      if (ce == cudaSuccess) {
        uint32_t* pool = reinterpret_cast<uint32_t*>(e->xbuf + qmk2::XB_POOL);
        const int* d_role = reinterpret_cast<const int*>(e->xbuf + qmk2::XB_ROLE);
        int r = rounds;
        void* cargs[] = {&pool, &r, &d_role, &d_bar, &d_out};
        ce = cudaLaunchCooperativeKernel((const void*)qmk2::qmk2_calib_kernel, dim3(qmk2::G2), dim3(NTHREADS), cargs, qmk2::SMEM2_BYTES, 0);
      }
      unsigned h_bar[2] = {0, 0};
      if (ce == cudaSuccess) ce = cudaMemcpy(t.data(), d_out, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      if (ce == cudaSuccess) ce = cudaMemcpy(h_bar, d_bar, sizeof(h_bar), cudaMemcpyDeviceToHost);
      if (getenv("QMK_CALIBRATE_VERBOSE")) fprintf(stderr, "[qmk] calibration: %s, barrier count %u, abandoned polls %u\n", cudaGetErrorString(ce), h_bar[0], h_bar[1]);
      if (ce == cudaSuccess && h_bar[1] != 0) ce = cudaErrorUnknown;   // unreliable measurement: keep the default slots
      if (d_bar) cudaFree(d_bar);
      if (d_out) cudaFree(d_out);
      if (ce == cudaSuccess) {
        // every group takes its two fastest slots that no other group has taken (groups in order of their best time)
        std::vector<char> used(qmk2::XP_CAND, 0);
        for (int g = 0; g < qmk2::NGRP; ++g) {
          for (int which = 0; which < 2; ++which) {
            int best = -1;
            for (int c = 0; c < qmk2::XP_CAND; ++c)
              if (!used[c] && t[(size_t)g * qmk2::XP_CAND + c] > 0 && (best < 0 || t[(size_t)g * qmk2::XP_CAND + c] < t[(size_t)g * qmk2::XP_CAND + best])) best = c;
            if (best >= 0) { slots[2 * g + which] = best; used[best] = 1; }
          }
        }
        bool ok = true;   // fall back to the default if a slot is duplicated (a measurement was missing)
        std::vector<char> seen(qmk2::XP_CAND, 0);
        for (int v : slots) { if (seen[v]) ok = false; seen[v] = 1; }
        if (!ok) for (int g = 0; g < qmk2::NGRP; ++g) { slots[2 * g] = 2 * g; slots[2 * g + 1] = 2 * g + 1; }
        if (getenv("QMK_CALIBRATE_VERBOSE")) {
          for (int g = 0; g < qmk2::NGRP; ++g) {
            long long mn = 1LL << 62, mx = 0;
            for (int c = 0; c < qmk2::XP_CAND; ++c) { long long v = t[(size_t)g * qmk2::XP_CAND + c]; if (v > 0) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; } }
            fprintf(stderr, "[qmk] group %d: exchange cycles per round min %lld max %lld, chosen slots %d (%lld) %d (%lld)\n", g, mn / rounds,
                    mx / rounds, slots[2 * g], t[(size_t)g * qmk2::XP_CAND + slots[2 * g]] / rounds, slots[2 * g + 1],
                    t[(size_t)g * qmk2::XP_CAND + slots[2 * g + 1]] / rounds);
          }
        }
      } else {
        cudaGetLastError();
      }
      // the calibration left epoch-tagged words behind
      if (err == cudaSuccess) err = cudaMemset(e->xbuf + qmk2::XB_LL, 0, qmk2::XB_ROLE - qmk2::XB_LL);
    }
    if (err == cudaSuccess) err = cudaMemcpy(e->xbuf + qmk2::XB_SLOTS, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice);
    e->slots = slots;
  }
  if (err != cudaSuccess) {
    if (e->xbuf) cudaFree(e->xbuf);
    if (e->res_spill) cudaFree(e->res_spill);
    if (e->delays) cudaFree(e->delays);
    if (e->status_dev) cudaFree(e->status_dev);
    delete e;
    return set_error(QMK_ERR_CUDA, "engine allocation failed: %s", cudaGetErrorString(err));
  }
  *out = e;
  return QMK_OK;
}

extern "C" void qmk_engine_destroy(qmk_engine* e) {
  if (!e) return;
  DeviceGuard guard(e->device);
  cudaDeviceSynchronize();
  cudaFree(e->xbuf);
  cudaFree(e->res_spill);
  cudaFree(e->delays);
  cudaFree(e->status_dev);
  if (e->trace_dev) cudaFree(e->trace_dev);
  if (e->handover) cudaEventDestroy(e->handover);
  delete e;
}

extern "C" int qmk_engine_trace_enable(qmk_engine* e, int stride) {
  if (!e || stride < 0) return set_error(QMK_ERR_ARG, "qmk_engine_trace_enable: bad argument");
  DeviceGuard guard(e->device);
  cudaDeviceSynchronize();
  if (e->trace_dev) cudaFree(e->trace_dev);
  e->trace_dev = nullptr;
  e->trace_stride = 0;
  if (stride == 0) return QMK_OK;
  QMK_CUDA(cudaMalloc(&e->trace_dev, sizeof(long long) * (size_t)e->G * stride));
  QMK_CUDA(cudaMemset(e->trace_dev, 0, sizeof(long long) * (size_t)e->G * stride));
  e->trace_stride = stride;
  return QMK_OK;
}

extern "C" int qmk_engine_trace_read(qmk_engine* e, void* stream, long long* host_out, int64_t max_elems) {
  if (!e || !host_out) return set_error(QMK_ERR_ARG, "qmk_engine_trace_read: null argument");
  if (!e->trace_dev) return set_error(QMK_ERR_ARG, "qmk_engine_trace_read: tracing is not enabled");
  DeviceGuard guard(e->device);
  QMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  int64_t n = (int64_t)e->G * e->trace_stride;
  if (n > max_elems) n = max_elems;
  QMK_CUDA(cudaMemcpy(host_out, e->trace_dev, sizeof(long long) * n, cudaMemcpyDeviceToHost));
  return e->trace_stride;
}

extern "C" int qmk_engine_num_ctas(const qmk_engine* e) { return e ? e->G : 0; }

extern "C" int qmk_engine_poll_stats(qmk_engine* e, void* stream, int32_t* host_out, int64_t max_elems) {
  if (!e || !host_out) return set_error(QMK_ERR_ARG, "qmk_engine_poll_stats: null argument");
  DeviceGuard guard(e->device);
  QMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  int64_t n = (int64_t)e->G * 3 * DL_N;
  if (n > max_elems) n = max_elems;
  QMK_CUDA(cudaMemcpy(host_out, e->delays, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  return 3 * DL_N;
}

extern "C" int qmk_engine_sync_status(qmk_engine* e, void* stream, int32_t* detail) {
  if (!e) return set_error(QMK_ERR_ARG, "qmk_engine_sync_status: engine is null");
  DeviceGuard guard(e->device);
  QMK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  int st[4] = {0, 0, 0, 0};
  QMK_CUDA(cudaMemcpy(st, e->status_dev, sizeof(st), cudaMemcpyDeviceToHost));
  if (detail) memcpy(detail, st, sizeof(st));
  if (st[0] != 0) {
    cudaMemset(e->status_dev, 0, sizeof(st));
    cudaMemset(e->xbuf, 0, e->version == 2 ? qmk2::XB_ROLE : e->xbuf_bytes);   // an aborted launch leaves the cumulative totals of the group kernel undefined
    e->epoch = 0;
    return set_error(QMK_ERR_KERNEL, "device watchdog fired: code %d (1=exchange wait 2=ring full wait 3=ring empty wait) cta %d phase %d aux %d",
                     st[0], st[1], st[2], st[3]);
  }
  return QMK_OK;
}

static int autotune_group_slots(qmk_model* m, cudaStream_t st);

extern "C" int qmk_model_create(qmk_engine* e, const LDGLayerWeights* layers, int num_layers,
                                const void* final_norm_weight, int residual_fp32, void* stream, qmk_model** out) {
  if (!e || !layers || !final_norm_weight || !out) return set_error(QMK_ERR_ARG, "qmk_model_create: null argument");
  if (num_layers < 1 || num_layers > 1024) return set_error(QMK_ERR_ARG, "qmk_model_create: bad num_layers %d", num_layers);
  static_assert(sizeof(LDGLayerWeights) == sizeof(LayerPtrs), "layer pointer struct mismatch");
  *out = nullptr;
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  qmk_model* m = new qmk_model();
  m->e = e;
  m->lay = make_layout(e->G, num_layers);
  m->residual_fp32 = residual_fp32 ? 1 : 0;
  const size_t bytes = e->version == 2 ? (size_t)qmk2::G2 * num_layers * qmk2::LAYER_BYTES2
                                       : (size_t)e->G * num_layers * m->lay.layer_segs * SEG_BYTES;
  cudaError_t err = cudaMalloc(&m->packed_layers, bytes);
  if (err == cudaSuccess) err = cudaMalloc(&m->aux_layers, (size_t)num_layers * 2 * AUX_BYTES);
  if (err == cudaSuccess) err = cudaMalloc(&m->norm_seg, AUX_BYTES);
  if (err == cudaSuccess) err = cudaMemsetAsync(m->norm_seg, 0, AUX_BYTES, st);
  if (err == cudaSuccess) {
    if (e->version == 2)
      qmk2::qmk2_pack_layers_kernel<<<dim3(qmk2::G2, num_layers), 128, 0, st>>>(reinterpret_cast<const LayerPtrs*>(layers), num_layers,
                                                                               m->packed_layers, m->aux_layers);
    else
      qmk_pack_layers_kernel<<<dim3(e->G, num_layers), 128, 0, st>>>(reinterpret_cast<const LayerPtrs*>(layers), m->lay,
                                                                    m->packed_layers, m->aux_layers);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMemcpyAsync(m->norm_seg, final_norm_weight, SEG_BYTES, cudaMemcpyDeviceToDevice, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) {
    qmk_model_destroy(m);
    return set_error(QMK_ERR_CUDA, "qmk_model_create: %s", cudaGetErrorString(err));
  }
  m->packed_bytes = (int64_t)bytes;
  *out = m;
  // First real model on a group-kernel engine: place the groups' exchange buffers by measuring them inside the running decode
  // kernel (the idle-system calibration of qmk_engine_create does not always predict the loaded behaviour: round 2 traced two
  // groups 0.9 k cycles per layer slower than the rest on slots the calibration had ranked best).
  if (e->version == 2 && !e->tuned && num_layers >= 5) {
    int want = 1;
    if (const char* env = getenv("QMK_AUTOTUNE")) want = atoi(env);
    if (want) {
      const int rc = autotune_group_slots(m, st);
      if (rc != QMK_OK && getenv("QMK_CALIBRATE_VERBOSE")) fprintf(stderr, "[qmk] autotune skipped: %s\n", qmk_last_error());
    }
    e->tuned = true;
  }
  return QMK_OK;
}

extern "C" int qmk_model_add_head(qmk_model* m, const void* lm_head_weight, int rows, void* stream) {
  if (!m || !lm_head_weight) return set_error(QMK_ERR_ARG, "qmk_model_add_head: null argument");
  if (rows < m->e->G || rows > MAX_HEAD_ROWS) return set_error(QMK_ERR_ARG, "qmk_model_add_head: rows %d out of range", rows);
  DeviceGuard guard(m->e->device);
  cudaStream_t st = (cudaStream_t)stream;
  qmk_head h;
  h.rows = rows;
  size_t bytes;
  if (m->e->version == 2) {
    if (rows % qmk2::G2 != 0 || rows / qmk2::G2 > 16 * qmk2::HEAD_TILES_MAX)
      return set_error(QMK_ERR_ARG, "qmk_model_add_head: the group kernel needs rows %% 128 == 0 and rows <= 4096 (got %d)", rows);
    h.segs_max = (rows / qmk2::G2 + 15) / 16;   // 16-row tiles per CTA
    bytes = (size_t)qmk2::G2 * h.segs_max * qmk2::SLOT2;
    QMK_CUDA(cudaMalloc(&h.packed, bytes));
    qmk2::qmk2_pack_head_kernel<<<qmk2::G2, 128, 0, st>>>(reinterpret_cast<const uint4*>(lm_head_weight), rows, h.packed);
  } else {
  h.segs_max = (rows + m->e->G - 1) / m->e->G;
  if (h.segs_max > MAX_ITEMS) return set_error(QMK_ERR_ARG, "qmk_model_add_head: %d rows per CTA exceed %d", h.segs_max, MAX_ITEMS);
  bytes = (size_t)m->e->G * h.segs_max * SEG_BYTES;
  QMK_CUDA(cudaMalloc(&h.packed, bytes));
  qmk_pack_head_kernel<<<m->e->G, 128, 0, st>>>(reinterpret_cast<const uint4*>(lm_head_weight), rows, m->e->G,
                                                h.segs_max, h.packed);
  }
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) {
    cudaFree(h.packed);
    return set_error(QMK_ERR_CUDA, "qmk_model_add_head: %s", cudaGetErrorString(err));
  }
  m->packed_bytes += (int64_t)bytes;
  m->heads.push_back(h);
  return (int)m->heads.size() - 1;
}

extern "C" int qmk_model_set_group_embedding(qmk_model* m, int group, const void* embedding_weight) {
  if (!m || group < 0 || group >= QMK_CP_GROUPS) return set_error(QMK_ERR_ARG, "qmk_model_set_group_embedding: bad argument");
  m->group_embed[group] = embedding_weight;
  return QMK_OK;
}

extern "C" void qmk_model_destroy(qmk_model* m) {
  if (!m) return;
  DeviceGuard guard(m->e->device);
  cudaDeviceSynchronize();
  if (m->packed_layers) cudaFree(m->packed_layers);
  if (m->aux_layers) cudaFree(m->aux_layers);
  if (m->norm_seg) cudaFree(m->norm_seg);
  for (auto& h : m->heads) cudaFree(h.packed);
  delete m;
}

extern "C" int64_t qmk_model_packed_bytes(const qmk_model* m) { return m ? m->packed_bytes : 0; }

// Called with e->mu held, before anything is enqueued for a launch on `st`.
static int order_after_previous_stream(qmk_engine* e, cudaStream_t st) {
  if (e->has_last_stream && e->last_stream != st) {
    if (!e->handover) QMK_CUDA(cudaEventCreateWithFlags(&e->handover, cudaEventDisableTiming));
    cudaError_t err = cudaEventRecord(e->handover, e->last_stream);
    if (err == cudaSuccess) err = cudaStreamWaitEvent(st, e->handover, 0);
    if (err != cudaSuccess) {   // the previous stream no longer exists: its work is ordered by a device-wide wait
      cudaGetLastError();
      QMK_CUDA(cudaDeviceSynchronize());
    }
  }
  e->last_stream = st;
  e->has_last_stream = true;
  return QMK_OK;
}

static int launch_slice(qmk_engine* e, Params& p, int begin, int end, cudaStream_t st) {
  p.phase_begin = begin;
  p.phase_end = end;
  void* args[] = {&p};
  const void* fn = e->trace_dev ? (const void*)qmk_decode_kernel_traced : (const void*)qmk_decode_kernel;
  size_t smem = SMEM_BYTES;
  if (e->version == 2) {
    fn = e->trace_dev ? (const void*)qmk2::qmk2_decode_kernel_traced : (const void*)qmk2::qmk2_decode_kernel;
    smem = qmk2::SMEM2_BYTES;
  }
  if (e->coop) QMK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(e->G), dim3(NTHREADS), args, smem, st));
  else QMK_CUDA(cudaLaunchKernel(fn, dim3(e->G), dim3(NTHREADS), args, smem, st));
  return QMK_OK;
}

// Fields shared by every launch of `m` on its engine.
static void fill_model(ModelDesc& d, const qmk_model* m, const void* cos_table, const void* sin_table, void* k_cache,
                       void* v_cache, int max_seq_len) {
  d.packed_layers = m->packed_layers;
  d.aux_layers = m->aux_layers;
  d.cos_t = reinterpret_cast<const __nv_bfloat16*>(cos_table);
  d.sin_t = reinterpret_cast<const __nv_bfloat16*>(sin_table);
  d.k_cache = reinterpret_cast<__nv_bfloat16*>(k_cache);
  d.v_cache = reinterpret_cast<__nv_bfloat16*>(v_cache);
  d.max_seq = max_seq_len;
  d.L = m->lay.L;
  d.residual_fp32 = m->residual_fp32;
  d.rope_axis[0] = m->rope_axis[0];
  d.rope_axis[1] = m->rope_axis[1];
}
static void init_step(StepDesc& sd, int position) {
  sd.position = position;
  sd.rope_pos[0] = sd.rope_pos[1] = sd.rope_pos[2] = position;
  sd.model = 0;
  sd.in_mode_next = -1;
  sd.pos_per_frame = 0;
  sd.epoch_off = 0;
  sd.group = -1;
}
// Epoch offsets of a finished step program (every step uses L + 2 epochs of its model).
static uint32_t finish_program(Params& p) {
  uint32_t off = 0;
  for (int s = 0; s < p.n_steps; ++s) {
    p.steps[s].epoch_off = (int)off;
    off += (uint32_t)p.models[p.steps[s].model].L + 2u;
  }
  p.frames.epochs_per_frame = (int)off;
  return off;
}
static void fill_common(Params& p, qmk_model* m, const void* cos_table, const void* sin_table, void* k_cache,
                        void* v_cache, int max_seq_len, float attn_scale) {
  qmk_engine* e = m->e;
  memset(&p, 0, sizeof(p));
  fill_model(p.models[0], m, cos_table, sin_table, k_cache, v_cache, max_seq_len);
  p.models[1] = p.models[0];
  p.sum_rows0 = p.sum_rows = 1;
  for (int s = 0; s < MAX_STEPS; ++s) init_step(p.steps[s], 0);
  p.lay = m->lay;
  p.packed_layers = m->packed_layers;
  p.aux_layers = m->aux_layers;
  p.cos_t = reinterpret_cast<const __nv_bfloat16*>(cos_table);
  p.sin_t = reinterpret_cast<const __nv_bfloat16*>(sin_table);
  p.k_cache = reinterpret_cast<__nv_bfloat16*>(k_cache);
  p.v_cache = reinterpret_cast<__nv_bfloat16*>(v_cache);
  p.max_seq = max_seq_len;
  p.attn_scale = attn_scale;
  p.residual_fp32 = m->residual_fp32;
  p.xbuf = e->xbuf;
  p.res_spill = e->res_spill;
  p.delays = e->delays;
  p.delay_o_idle = e->delay_o_idle;
  p.o_sentinel = e->o_sentinel;
  p.status = e->status_dev;
  p.timeout_cycles = e->timeout_cycles;
  p.trace = e->trace_dev;
  p.trace_stride = e->trace_stride;
  p.sample_temperature = 1.0f;
}
static void set_head(StepDesc& sd, qmk_model* m, int head_index) {
  sd.head.aux = m->norm_seg;
  if (head_index >= 0) {
    sd.head.packed = m->heads[head_index].packed;
    sd.head.rows = m->heads[head_index].rows;
    sd.head.segs_max = m->heads[head_index].segs_max;
  } else {
    sd.head.packed = nullptr;
    sd.head.rows = 0;
    sd.head.segs_max = 0;
  }
}
// 16-bit epochs: when the counter would wrap, clear the exchange words (stream-ordered) and restart at 0.
static int reserve_epochs(qmk_engine* e, uint32_t need, cudaStream_t st, uint32_t* base) {
  if (e->epoch + need >= 0xfff0u) {
    const size_t ll_end = e->version == 2 ? qmk2::XB_ROLE : e->xbuf_bytes;
    QMK_CUDA(cudaMemsetAsync(e->xbuf + e->xbuf_ll_off, 0, ll_end - e->xbuf_ll_off, st));
    e->epoch = 0;
  }
  *base = e->epoch;
  e->epoch += need;
  return QMK_OK;
}

static int decode_step_impl(qmk_model* m, int head_index, int input_token_id, const void* embed_weight,
                            const int64_t* codes, const void* const* group_tables, int talker_vocab, int group_vocab,
                            const void* extra_bf16, const void* cos_table, const void* sin_table, void* k_cache, void* v_cache,
                            void* hidden_buffer, float* normalized_out, int32_t* out_token, int position,
                            const int32_t* rope_pos, int max_seq_len, float attn_scale, int mode, void* stream) {
  if (!m) return set_error(QMK_ERR_ARG, "qmk_decode_step: model is null");
  if (!cos_table || !sin_table || !k_cache || !v_cache || !hidden_buffer)
    return set_error(QMK_ERR_ARG, "qmk_decode_step: null table / cache / hidden_buffer pointer");
  if (max_seq_len < 1 || position < 0 || position >= max_seq_len)
    return set_error(QMK_ERR_ARG, "qmk_decode_step: position %d outside [0, max_seq_len=%d)", position, max_seq_len);
  if (input_token_id >= 0 && !embed_weight) return set_error(QMK_ERR_ARG, "qmk_decode_step: token given but embed_weight is null");
  if (head_index >= (int)m->heads.size()) return set_error(QMK_ERR_ARG, "qmk_decode_step: head index %d not registered", head_index);
  if (head_index >= 0 && !out_token) return set_error(QMK_ERR_ARG, "qmk_decode_step: out_token is null");
  qmk_engine* e = m->e;
  if (rope_pos) {
    if (e->version != 2) return set_error(QMK_ERR_UNSUPPORTED, "M-RoPE positions need the group kernel");
    for (int a = 0; a < 3; ++a)
      if (rope_pos[a] < 0) return set_error(QMK_ERR_ARG, "qmk_decode_step_mrope: negative rope position");
  }
  if (codes && (talker_vocab < 1 || group_vocab < 1)) return set_error(QMK_ERR_ARG, "qmk_decode_step_codes: table row counts must be positive");
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(e->mu);
  int rc = order_after_previous_stream(e, st);
  if (rc != QMK_OK) return rc;

  Params p;
  fill_common(p, m, cos_table, sin_table, k_cache, v_cache, max_seq_len, attn_scale);
  p.n_steps = 1;
  StepDesc& sd = p.steps[0];
  init_step(sd, position);
  if (rope_pos) { sd.rope_pos[0] = rope_pos[0]; sd.rope_pos[1] = rope_pos[1]; sd.rope_pos[2] = rope_pos[2]; }
  sd.in_table = reinterpret_cast<const __nv_bfloat16*>(embed_weight);
  sd.in_vec = hidden_buffer;
  sd.in_mode = input_token_id >= 0 ? IN_TABLE_TOKEN : IN_VEC_BF16;
  sd.token = input_token_id >= 0 ? input_token_id : 0;
  if (codes) {
    sd.in_mode = IN_CODES_SUM;
    sd.codes = reinterpret_cast<const long long*>(codes);
    sd.in_vec = extra_bf16;
    for (int g = 0; g < 15; ++g) p.sum_tables[g] = reinterpret_cast<const __nv_bfloat16*>(group_tables[g]);
    p.sum_rows0 = talker_vocab;
    p.sum_rows = group_vocab;
  }
  set_head(sd, m, head_index);
  sd.out_token = out_token;
  sd.out_norm = normalized_out;
  sd.hidden_out = reinterpret_cast<__nv_bfloat16*>(hidden_buffer);

  rc = reserve_epochs(e, finish_program(p), st, &p.epoch_base);
  if (rc != QMK_OK) return rc;

  const int n_idx = m->lay.L * PH_PER_LAYER + 2;
  if (mode == 0 || e->version == 2) return launch_slice(e, p, 0, n_idx, st);   // the group kernel has no staged mode
  for (int idx = 0; idx < n_idx; ++idx) {
    rc = launch_slice(e, p, idx, idx + 1, st);
    if (rc != QMK_OK) return rc;
  }
  return QMK_OK;
}

extern "C" int qmk_decode_step(qmk_model* m, int head_index, int input_token_id, const void* embed_weight,
                               const void* cos_table, const void* sin_table, void* k_cache, void* v_cache,
                               void* hidden_buffer, float* normalized_out, int32_t* out_token, int position,
                               int max_seq_len, float attn_scale, int mode, void* stream) {
  return decode_step_impl(m, head_index, input_token_id, embed_weight, nullptr, nullptr, 0, 0, nullptr, cos_table, sin_table,
                          k_cache, v_cache, hidden_buffer, normalized_out, out_token, position, nullptr, max_seq_len,
                          attn_scale, mode, stream);
}

// ---------------------------------------------------------------------------------------------------
// In-situ placement of the group buffers.  24 trials: in trial t group g uses pool slot (2 t + 6 g) % 48 for its q/k/v buffer and
// the next one for its m buffer (all distinct); a trial runs a few real decode steps of `m` on scratch buffers and reads, per
// CTA, the cycles between a CTA's publish and the completion of its gather for the two group-local exchanges (kernel statistics
// row 2).  Every group then takes the q slot and the m slot with the smallest mean wait.  ~40 ms, once per engine.
// ---------------------------------------------------------------------------------------------------
static int autotune_group_slots(qmk_model* m, cudaStream_t st) {
  qmk_engine* e = m->e;
  const int L = m->lay.L, S = 8, G = qmk2::G2, steps = 4, trials = qmk2::XP_CAND / 2;
  const size_t kv_elems = (size_t)L * NKVH * S * HD;
  __nv_bfloat16 *kc = nullptr, *vc = nullptr, *tab = nullptr, *hid = nullptr;
  cudaError_t err = cudaMalloc(&kc, kv_elems * 2);
  if (err == cudaSuccess) err = cudaMalloc(&vc, kv_elems * 2);
  if (err == cudaSuccess) err = cudaMalloc(&tab, (size_t)S * HD * 2);
  if (err == cudaSuccess) err = cudaMalloc(&hid, H * 2);
  if (err == cudaSuccess) err = cudaMemsetAsync(kc, 0, kv_elems * 2, st);
  if (err == cudaSuccess) err = cudaMemsetAsync(vc, 0, kv_elems * 2, st);
  if (err == cudaSuccess) err = cudaMemsetAsync(tab, 0, (size_t)S * HD * 2, st);
  std::vector<uint16_t> ones(H, 0x3F80);   // bf16 1.0: a finite, non-zero input
  if (err == cudaSuccess) err = cudaMemcpyAsync(hid, ones.data(), H * 2, cudaMemcpyHostToDevice, st);
  auto cleanup = [&]() { cudaFree(kc); cudaFree(vc); cudaFree(tab); cudaFree(hid); };
  if (err != cudaSuccess) { cleanup(); return set_error(QMK_ERR_CUDA, "autotune: %s", cudaGetErrorString(err)); }
  const std::vector<int> saved = e->slots;
  std::vector<int> stats((size_t)G * 3 * DL_N), zero((size_t)G * 3 * DL_N);
  std::vector<double> cost_q((size_t)qmk2::NGRP * qmk2::XP_CAND, -1.0), cost_m = cost_q;
  std::vector<int> delays0((size_t)G * 3 * DL_N);
  err = cudaMemcpy(delays0.data(), e->delays, delays0.size() * sizeof(int), cudaMemcpyDeviceToHost);
  int rc = QMK_OK;
  for (int t = 0; t < trials && rc == QMK_OK && err == cudaSuccess; ++t) {
    std::vector<int> slots(qmk2::NGRP * 2);
    for (int g = 0; g < qmk2::NGRP; ++g) { slots[2 * g] = (2 * t + 6 * g) % qmk2::XP_CAND; slots[2 * g + 1] = (2 * t + 1 + 6 * g) % qmk2::XP_CAND; }
    err = cudaMemcpyAsync(e->xbuf + qmk2::XB_SLOTS, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, st);
    // statistics rows 1, 2 start from zero (row 0 = the poll delays, kept)
    for (int c = 0; c < G; ++c) for (int k = 0; k < 3 * DL_N; ++k) zero[(size_t)c * 3 * DL_N + k] = k < DL_N ? delays0[(size_t)c * 3 * DL_N + k] : 0;
    if (err == cudaSuccess) err = cudaMemcpyAsync(e->delays, zero.data(), zero.size() * sizeof(int), cudaMemcpyHostToDevice, st);
    for (int s = 0; s < steps && rc == QMK_OK && err == cudaSuccess; ++s)
      rc = decode_step_impl(m, -1, -1, nullptr, nullptr, nullptr, 0, 0, nullptr, tab, tab, kc, vc, hid, nullptr, nullptr, 1 + s, nullptr, S,
                            0.08838834764831845f, 0, st);
    if (rc != QMK_OK || err != cudaSuccess) break;
    err = cudaStreamSynchronize(st);
    if (err == cudaSuccess) err = cudaMemcpy(stats.data(), e->delays, stats.size() * sizeof(int), cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) break;
    for (int g = 0; g < qmk2::NGRP; ++g) {
      double wq = 0, wm = 0;
      for (int b = 0; b < G; ++b) {           // blockIdx b runs role e->roles[b]
        if (e->roles[b] / qmk2::GSZ != g) continue;
        wq += stats[(size_t)b * 3 * DL_N + 2 * DL_N + DL_ATTN];
        wm += stats[(size_t)b * 3 * DL_N + 2 * DL_N + DL_DOWN];
      }
      cost_q[(size_t)g * qmk2::XP_CAND + slots[2 * g]] = wq;
      cost_m[(size_t)g * qmk2::XP_CAND + slots[2 * g + 1]] = wm;
    }
  }
  // the scratch run wrote the hidden buffer only; restore the statistics rows
  cudaMemcpy(e->delays, delays0.data(), delays0.size() * sizeof(int), cudaMemcpyHostToDevice);
  std::vector<int> best = saved;
  if (rc == QMK_OK && err == cudaSuccess) {
    std::vector<char> used(qmk2::XP_CAND, 0);
    for (int which = 0; which < 2; ++which)
      for (int g = 0; g < qmk2::NGRP; ++g) {
        const std::vector<double>& cost = which ? cost_m : cost_q;
        int bi = -1;
        for (int c = 0; c < qmk2::XP_CAND; ++c)
          if (!used[c] && cost[(size_t)g * qmk2::XP_CAND + c] > 0 && (bi < 0 || cost[(size_t)g * qmk2::XP_CAND + c] < cost[(size_t)g * qmk2::XP_CAND + bi])) bi = c;
        if (bi < 0) { best = saved; which = 2; break; }
        best[2 * g + which] = bi;
        used[bi] = 1;
      }
    if (getenv("QMK_CALIBRATE_VERBOSE")) {
      for (int g = 0; g < qmk2::NGRP; ++g) {
        double mnq = 1e30, mxq = 0, mnm = 1e30, mxm = 0;
        for (int c = 0; c < qmk2::XP_CAND; ++c) {
          const double a = cost_q[(size_t)g * qmk2::XP_CAND + c], b = cost_m[(size_t)g * qmk2::XP_CAND + c];
          if (a > 0) { mnq = std::min(mnq, a); mxq = std::max(mxq, a); }
          if (b > 0) { mnm = std::min(mnm, b); mxm = std::max(mxm, b); }
        }
        const double per = 16.0 / (double)(steps * L * qmk2::GSZ);   // statistics are cycles / 16 summed over steps, layers and the group's CTAs
        fprintf(stderr, "[qmk] autotune group %d: q/k/v wait %.0f..%.0f cycles over the slots -> slot %d; m wait %.0f..%.0f -> slot %d\n", g,
                mnq * per, mxq * per, best[2 * g], mnm * per, mxm * per, best[2 * g + 1]);
      }
    }
  }
  cudaMemcpy(e->xbuf + qmk2::XB_SLOTS, best.data(), best.size() * sizeof(int), cudaMemcpyHostToDevice);
  e->slots = best;
  cleanup();
  if (err != cudaSuccess) return set_error(QMK_ERR_CUDA, "autotune: %s", cudaGetErrorString(err));
  return rc;
}

extern "C" int qmk_decode_step_mrope(qmk_model* m, int head_index, int input_token_id, const void* embed_weight,
                                     const void* cos_table, const void* sin_table, void* k_cache, void* v_cache,
                                     void* hidden_buffer, float* normalized_out, int32_t* out_token, int position,
                                     const int32_t* rope_pos, int max_seq_len, float attn_scale, void* stream) {
  if (!rope_pos) return set_error(QMK_ERR_ARG, "qmk_decode_step_mrope: rope_pos is null");
  return decode_step_impl(m, head_index, input_token_id, embed_weight, nullptr, nullptr, 0, 0, nullptr, cos_table, sin_table,
                          k_cache, v_cache, hidden_buffer, normalized_out, out_token, position, rope_pos, max_seq_len,
                          attn_scale, 0, stream);
}

// Axis map of multimodal RoPE.  Chunked (Qwen2-VL `apply_multimodal_rotary_pos_emb`: the 64 rotary frequencies are cut into
// consecutive sections of section[0], section[1], section[2] entries that take the temporal / height / width position) or
// interleaved (Qwen3-VL `apply_interleaved_mrope`: frequency i < 3 section[a] with i % 3 == a takes axis a = 1, 2; the rest axis 0).
extern "C" int qmk_model_set_mrope(qmk_model* m, const int32_t* section, int interleaved) {
  if (!m) return set_error(QMK_ERR_ARG, "qmk_model_set_mrope: model is null");
  m->rope_axis[0] = m->rope_axis[1] = 0ull;
  if (!section) return QMK_OK;   // back to standard RoPE
  if (m->e->version != 2) return set_error(QMK_ERR_UNSUPPORTED, "M-RoPE needs the group kernel");
  if (section[0] < 0 || section[1] < 0 || section[2] < 0 || section[0] + section[1] + section[2] != HD / 2)
    return set_error(QMK_ERR_ARG, "qmk_model_set_mrope: sections %d + %d + %d must add up to %d", section[0], section[1], section[2], HD / 2);
  for (int i = 0; i < HD / 2; ++i) {
    unsigned long long axis;
    if (interleaved) axis = (i % 3 == 1 && i < 3 * section[1]) ? 1ull : ((i % 3 == 2 && i < 3 * section[2]) ? 2ull : 0ull);
    else axis = i < section[0] ? 0ull : (i < section[0] + section[1] ? 1ull : 2ull);
    m->rope_axis[i >> 5] |= axis << (2 * (i & 31));
  }
  return QMK_OK;
}

extern "C" int qmk_decode_step_codes(qmk_model* m, int head_index, const int64_t* codes, const void* talker_embed_weight,
                                     int talker_vocab, const void* const* group_embedding_tables, int group_vocab,
                                     const void* extra_embed_bf16, const void* cos_table, const void* sin_table, void* k_cache,
                                     void* v_cache, void* hidden_buffer, float* normalized_out, int32_t* out_token, int position,
                                     int max_seq_len, float attn_scale, void* stream) {
  if (!codes || !talker_embed_weight || !group_embedding_tables || !extra_embed_bf16)
    return set_error(QMK_ERR_ARG, "qmk_decode_step_codes: null argument");
  for (int g = 0; g < 15; ++g)
    if (!group_embedding_tables[g]) return set_error(QMK_ERR_ARG, "qmk_decode_step_codes: group table %d is null", g);
  return decode_step_impl(m, head_index, -1, talker_embed_weight, codes, group_embedding_tables, talker_vocab, group_vocab,
                          extra_embed_bf16, cos_table, sin_table, k_cache, v_cache, hidden_buffer, normalized_out, out_token,
                          position, nullptr, max_seq_len, attn_scale, 0, stream);
}

// Step program of one code-predictor frame on model slot `mi` (16 steps, 15 heads; upstream model_tts.py:742-773).
static void fill_cp_steps(Params& p, int mi, qmk_model* m, const float* talker_hidden, int first_codebook_token,
                          const int32_t* first_token_dev, int talker_vocab, const void* talker_embed_weight, bool sample,
                          int64_t* out_codes, float* logits_out, float* hidden_out) {
  for (int s = 0; s < QMK_CP_GROUPS + 1; ++s) {
    StepDesc& sd = p.steps[s];
    init_step(sd, s);
    sd.model = mi;
    sd.group = s - 1;
    if (s == 0) {           // the talker's hidden state (upstream model_tts.py:745)
      sd.in_mode = IN_VEC_F32;
      sd.in_vec = talker_hidden;
      set_head(sd, m, -1);
    } else {
      if (s == 1) {         // embedding of the talker's token (model_tts.py:746-748)
        sd.in_mode = IN_TABLE_TOKEN;
        sd.in_table = reinterpret_cast<const __nv_bfloat16*>(talker_embed_weight);
        sd.token = first_token_dev ? talker_vocab - 1 : first_codebook_token;   // device token: clamp bound
        sd.token_ptr = first_token_dev;
      } else {              // embedding of the previous group's code (model_tts.py:768-770)
        sd.in_mode = IN_TABLE_PREV;
        sd.in_table = reinterpret_cast<const __nv_bfloat16*>(m->group_embed[s - 2]);
        sd.token = m->heads[s - 2].rows - 1;   // clamp bound: the previous head's vocabulary = rows of its embedding table
      }
      set_head(sd, m, s - 1);
      sd.select = sample ? 1 : 0;
      sd.out_code = out_codes ? reinterpret_cast<long long*>(out_codes) + s : nullptr;
      sd.logits_out = logits_out ? logits_out + (size_t)(s - 1) * m->heads[s - 1].rows : nullptr;
      sd.out_norm = hidden_out ? hidden_out + (size_t)(s - 1) * H : nullptr;
    }
  }
}
static int check_cp_model(const qmk_model* m, const char* who) {
  if ((int)m->heads.size() < QMK_CP_GROUPS) return set_error(QMK_ERR_ARG, "%s: %d group heads registered, need 15", who, (int)m->heads.size());
  for (int g = 0; g < QMK_CP_GROUPS - 1; ++g)
    if (!m->group_embed[g]) return set_error(QMK_ERR_ARG, "%s: group embedding %d not set", who, g);
  return QMK_OK;
}

static int cp_predict_impl(qmk_model* m, const float* talker_hidden, int first_codebook_token,
                           const int32_t* first_token_dev, int talker_vocab, const void* talker_embed_weight,
                           const void* cos_table, const void* sin_table, void* k_cache, void* v_cache, int max_seq_len,
                           int do_sample, float temperature, int top_k, uint64_t seed, uint64_t frame_counter,
                           const int32_t* forced_tokens, int64_t* out_codes, float* logits_out, float* hidden_out,
                           void* stream) {
  if (!m || !talker_hidden || !talker_embed_weight || !cos_table || !sin_table || !k_cache || !v_cache || !out_codes)
    return set_error(QMK_ERR_ARG, "qmk_cp_predict: null argument");
  if (max_seq_len < QMK_CP_GROUPS + 1) return set_error(QMK_ERR_ARG, "qmk_cp_predict: max_seq_len %d < 16", max_seq_len);
  if (!first_token_dev && first_codebook_token < 0) return set_error(QMK_ERR_ARG, "qmk_cp_predict: negative first token");
  if (first_token_dev && talker_vocab < 1) return set_error(QMK_ERR_ARG, "qmk_cp_predict_dev: talker_vocab must be positive");
  int rc = check_cp_model(m, "qmk_cp_predict");
  if (rc != QMK_OK) return rc;
  const bool sample = do_sample && temperature > 0.f;
  qmk_engine* e = m->e;
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(e->mu);
  rc = order_after_previous_stream(e, st);
  if (rc != QMK_OK) return rc;

  Params p;
  fill_common(p, m, cos_table, sin_table, k_cache, v_cache, max_seq_len, 0.08838834764831845f /* 1/sqrt(128) */);
  p.n_steps = QMK_CP_GROUPS + 1;
  p.sample_temperature = sample ? temperature : 1.0f;
  p.sample_top_k = top_k;
  p.sample_seed = seed;
  p.sample_counter = frame_counter;
  p.forced_tokens = forced_tokens;
  p.code0_out = reinterpret_cast<long long*>(out_codes);
  p.code0 = first_codebook_token;
  p.code0_ptr = first_token_dev;
  fill_cp_steps(p, 0, m, talker_hidden, first_codebook_token, first_token_dev, talker_vocab, talker_embed_weight, sample, out_codes,
                logits_out, hidden_out);
  rc = reserve_epochs(e, finish_program(p), st, &p.epoch_base);
  if (rc != QMK_OK) return rc;
  return launch_slice(e, p, 0, m->lay.L * PH_PER_LAYER + 2, st);
}

extern "C" int qmk_cp_predict(qmk_model* m, const float* talker_hidden, int first_codebook_token,
                              const void* talker_embed_weight, const void* cos_table, const void* sin_table,
                              void* k_cache, void* v_cache, int max_seq_len, int do_sample, float temperature,
                              int top_k, uint64_t seed, uint64_t frame_counter, const int32_t* forced_tokens,
                              int64_t* out_codes, float* logits_out, float* hidden_out, void* stream) {
  return cp_predict_impl(m, talker_hidden, first_codebook_token, nullptr, 0, talker_embed_weight, cos_table, sin_table,
                         k_cache, v_cache, max_seq_len, do_sample, temperature, top_k, seed, frame_counter, forced_tokens,
                         out_codes, logits_out, hidden_out, stream);
}

extern "C" int qmk_cp_predict_dev(qmk_model* m, const float* talker_hidden, const int32_t* first_token_dev, int talker_vocab,
                                  const void* talker_embed_weight, const void* cos_table, const void* sin_table,
                                  void* k_cache, void* v_cache, int max_seq_len, int do_sample, float temperature,
                                  int top_k, uint64_t seed, uint64_t frame_counter, int64_t* out_codes, void* stream) {
  if (!first_token_dev) return set_error(QMK_ERR_ARG, "qmk_cp_predict_dev: first_token_dev is null");
  return cp_predict_impl(m, talker_hidden, 0, first_token_dev, talker_vocab, talker_embed_weight, cos_table, sin_table,
                         k_cache, v_cache, max_seq_len, do_sample, temperature, top_k, seed, frame_counter, nullptr,
                         out_codes, nullptr, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------------
// Device-autonomous frame loop: N codec frames (code-predictor frame + embedding sum + talker step each) with no host
// round trip, EOS detected on the device.  Upstream analogue: launch_ldg_generate_nosync (kernel.cu:1555-1613) with
// ldg_update_step (:1437-1448) between steps; here the loop lives inside the persistent kernel.
// ---------------------------------------------------------------------------------------------------
extern "C" int qmk_generate_args_size(void) { return (int)sizeof(qmk_generate_args); }

extern "C" int qmk_generate_nosync(const qmk_generate_args* a, void* stream) {
  if (!a) return set_error(QMK_ERR_ARG, "qmk_generate_nosync: args is null");
  qmk_model* tk = a->talker;
  qmk_model* cp = a->cp;
  if (!tk || !cp) return set_error(QMK_ERR_ARG, "qmk_generate_nosync: model is null");
  if (tk->e != cp->e) return set_error(QMK_ERR_ARG, "qmk_generate_nosync: both models must live on the same engine");
  qmk_engine* e = tk->e;
  if (e->version != 2) return set_error(QMK_ERR_UNSUPPORTED, "qmk_generate_nosync needs the group kernel");
  if (!a->talker_embed_weight || !a->talker_cos || !a->talker_sin || !a->talker_k_cache || !a->talker_v_cache || !a->hidden_buffer ||
      !a->talker_hidden || !a->talker_token || !a->cp_cos || !a->cp_sin || !a->cp_k_cache || !a->cp_v_cache ||
      !a->group_embedding_tables || !a->pad_embed || !a->codes_out || !a->gen_state)
    return set_error(QMK_ERR_ARG, "qmk_generate_nosync: null pointer argument");
  if (a->talker_head < 0 || a->talker_head >= (int)tk->heads.size()) return set_error(QMK_ERR_ARG, "qmk_generate_nosync: talker head %d not registered", a->talker_head);
  int rc = check_cp_model(cp, "qmk_generate_nosync");
  if (rc != QMK_OK) return rc;
  for (int g = 0; g < 15; ++g)
    if (!a->group_embedding_tables[g]) return set_error(QMK_ERR_ARG, "qmk_generate_nosync: group table %d is null", g);
  if (a->n_frames < 1) return set_error(QMK_ERR_ARG, "qmk_generate_nosync: n_frames %d", a->n_frames);
  if (a->position < 0 || a->position + a->n_frames > a->talker_max_seq)
    return set_error(QMK_ERR_ARG, "qmk_generate_nosync: positions %d .. %d exceed max_seq_len %d", a->position, a->position + a->n_frames - 1, a->talker_max_seq);
  if (a->cp_max_seq < QMK_CP_GROUPS + 1) return set_error(QMK_ERR_ARG, "qmk_generate_nosync: cp_max_seq %d < 16", a->cp_max_seq);
  if (a->talker_vocab < 1 || a->cp_vocab < 1) return set_error(QMK_ERR_ARG, "qmk_generate_nosync: vocab sizes must be positive");
  if (a->n_trailing < 0 || a->trailing_offset < 0 || (a->n_trailing > 0 && !a->trailing_text)) return set_error(QMK_ERR_ARG, "qmk_generate_nosync: bad trailing text");
  const bool sample = a->do_sample && a->temperature > 0.f;
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(e->mu);
  rc = order_after_previous_stream(e, st);
  if (rc != QMK_OK) return rc;
  if (a->reset_state) QMK_CUDA(cudaMemsetAsync(a->gen_state, 0, 4 * sizeof(int32_t), st));

  const uint32_t epochs_per_frame = (uint32_t)(QMK_CP_GROUPS + 1) * ((uint32_t)cp->lay.L + 2u) + (uint32_t)tk->lay.L + 2u;
  int max_chunk = (int)(0x7000u / epochs_per_frame);   // 16-bit epochs: a launch must not wrap (the buffer is cleared between launches)
  if (max_chunk < 1) return set_error(QMK_ERR_UNSUPPORTED, "qmk_generate_nosync: a frame needs %u epochs", epochs_per_frame);
  if (const char* env = getenv("QMK_FRAMES_PER_LAUNCH")) { int v = atoi(env); if (v >= 1 && v < max_chunk) max_chunk = v; }
  for (int done = 0; done < a->n_frames;) {
    const int chunk = std::min(max_chunk, a->n_frames - done);
    Params p;
    fill_common(p, tk, a->talker_cos, a->talker_sin, a->talker_k_cache, a->talker_v_cache, a->talker_max_seq, 0.08838834764831845f);
    fill_model(p.models[1], cp, a->cp_cos, a->cp_sin, a->cp_k_cache, a->cp_v_cache, a->cp_max_seq);
    p.n_steps = QMK_CP_GROUPS + 2;
    p.sample_temperature = sample ? a->temperature : 1.0f;
    p.sample_top_k = a->top_k;
    p.sample_seed = a->seed;
    p.sample_counter = a->frame_counter + (uint64_t)done;
    p.code0_ptr = a->talker_token;
    fill_cp_steps(p, 1, cp, a->talker_hidden, 0, a->talker_token, a->talker_vocab, a->talker_embed_weight, sample, nullptr, nullptr, nullptr);
    p.steps[0].in_mode_next = IN_PREV_NORM;    // later frames: the talker's hidden state never leaves the chip
    p.steps[1].in_mode_next = IN_TABLE_PREV;   // ... nor does its token
    StepDesc& sd = p.steps[QMK_CP_GROUPS + 1];
    init_step(sd, a->position + done);
    if (a->rope_pos) for (int x = 0; x < 3; ++x) sd.rope_pos[x] = a->rope_pos[x] + done;
    sd.model = 0;
    sd.pos_per_frame = 1;
    sd.in_mode = IN_CODES_SUM;                 // codes == null: the codes this CTA selected itself
    sd.in_table = reinterpret_cast<const __nv_bfloat16*>(a->talker_embed_weight);
    sd.in_vec = a->pad_embed;
    for (int g = 0; g < 15; ++g) p.sum_tables[g] = reinterpret_cast<const __nv_bfloat16*>(a->group_embedding_tables[g]);
    p.sum_rows0 = a->talker_vocab;
    p.sum_rows = a->cp_vocab;
    set_head(sd, tk, a->talker_head);
    sd.out_token = a->talker_token;
    sd.out_norm = a->talker_hidden;
    sd.hidden_out = reinterpret_cast<__nv_bfloat16*>(a->hidden_buffer);
    p.frames.n_frames = chunk;
    p.frames.eos_token = a->eos_token;
    p.frames.trailing = reinterpret_cast<const __nv_bfloat16*>(a->trailing_text);
    p.frames.n_trailing = a->n_trailing;
    p.frames.trailing_offset = a->trailing_offset;
    p.frames.pad_embed = reinterpret_cast<const __nv_bfloat16*>(a->pad_embed);
    p.frames.codes_out = reinterpret_cast<long long*>(a->codes_out) + (size_t)done * 16;
    p.frames.tokens_out = a->tokens_out ? a->tokens_out + done : nullptr;
    p.frames.gen_state = a->gen_state;
    if (finish_program(p) != epochs_per_frame) return set_error(QMK_ERR_ARG, "qmk_generate_nosync: internal epoch accounting");
    rc = reserve_epochs(e, epochs_per_frame * (uint32_t)chunk, st, &p.epoch_base);
    if (rc != QMK_OK) return rc;
    rc = launch_slice(e, p, 0, tk->lay.L * PH_PER_LAYER + 2, st);
    if (rc != QMK_OK) return rc;
    done += chunk;
  }
  return QMK_OK;
}

// ---------------------------------------------------------------------------------------------------
// upstream-compatible entry (kernel.cu:1485-1513).  Process-wide cache: device -> engine,
// (device, blob pointer, num_layers) -> re-packed model.
// ---------------------------------------------------------------------------------------------------
namespace {
struct LegacyCfg {
  int residual_fp32 = -1;   // -1 = infer from num_layers (see launch_ldg_decode_direct)
  int lm_head_rows = -1;    // -1 = 3072 unless the head table is all zero (upstream's dummy), 0 = skip
};
struct LegacyModel {
  qmk_model* m = nullptr;
  const void* final_norm = nullptr;
  int residual_fp32 = 1;
  std::vector<uint64_t> blob;            // host copy of the 11 * L pointers the model was packed from
  std::map<const void*, int> heads;      // LM-head table -> head index (-1: all-zero dummy table, no head)
};
std::mutex g_legacy_mu;
std::map<int, qmk_engine*> g_engines;
std::map<std::tuple<int, const void*, int>, LegacyModel> g_models;
std::map<const void*, LegacyCfg> g_cfg;
int g_legacy_status = QMK_OK;

// 1 if the bf16 table is entirely zero (upstream's CodePredictorKernel passes torch.zeros as LM head, model_tts.py:657-659)
__global__ void qmk_any_nonzero_kernel(const uint4* p, size_t n16, int* flag) {
  int nz = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    nz |= (v.x | v.y | v.z | v.w) != 0u;
  }
  if (nz) *flag = 1;
}
int table_is_zero(const void* table, size_t bytes, cudaStream_t st, bool* out) {
  int* d_flag = nullptr;
  int h_flag = 0;
  QMK_CUDA(cudaMalloc(&d_flag, sizeof(int)));
  cudaError_t err = cudaMemsetAsync(d_flag, 0, sizeof(int), st);
  if (err == cudaSuccess) {
    qmk_any_nonzero_kernel<<<148, 256, 0, st>>>(reinterpret_cast<const uint4*>(table), bytes / 16, d_flag);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  cudaFree(d_flag);
  if (err != cudaSuccess) return set_error(QMK_ERR_CUDA, "table_is_zero: %s", cudaGetErrorString(err));
  *out = h_flag == 0;
  return QMK_OK;
}
}  // namespace

extern "C" int qmk_legacy_configure(const LDGLayerWeights* layer_weights, int residual_fp32, int lm_head_rows) {
  if (!layer_weights) return set_error(QMK_ERR_ARG, "qmk_legacy_configure: null blob");
  if (lm_head_rows > 0 && (lm_head_rows < 64 || lm_head_rows > MAX_HEAD_ROWS))
    return set_error(QMK_ERR_ARG, "qmk_legacy_configure: lm_head_rows %d", lm_head_rows);
  std::lock_guard<std::mutex> lock(g_legacy_mu);
  LegacyCfg c;
  c.residual_fp32 = residual_fp32 < 0 ? -1 : (residual_fp32 ? 1 : 0);
  c.lm_head_rows = lm_head_rows < 0 ? -1 : lm_head_rows;
  g_cfg[layer_weights] = c;
  return QMK_OK;
}

extern "C" int qmk_legacy_status(void) { return g_legacy_status; }

// Drop the re-packed copy of one blob (all devices): call after the weights behind it changed, or when the blob's
// address may be re-used by another model.  Null drops every cached model but keeps the engines.
extern "C" void qmk_legacy_invalidate(const LDGLayerWeights* layer_weights) {
  std::lock_guard<std::mutex> lock(g_legacy_mu);
  for (auto it = g_models.begin(); it != g_models.end();) {
    if (!layer_weights || std::get<1>(it->first) == (const void*)layer_weights) {
      qmk_model_destroy(it->second.m);
      it = g_models.erase(it);
    } else {
      ++it;
    }
  }
  if (layer_weights) g_cfg.erase(layer_weights); else g_cfg.clear();
}

// Synchronise `stream` and return + clear the device-side status of the legacy engine of the current device
// (QMK_OK or QMK_ERR_KERNEL): what qmk_engine_sync_status is for engines the caller owns.
extern "C" int qmk_legacy_sync_status(void* stream) {
  std::lock_guard<std::mutex> lock(g_legacy_mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_error(QMK_ERR_CUDA, "qmk_legacy_sync_status: no CUDA device");
  auto it = g_engines.find(dev);
  if (it == g_engines.end()) return QMK_OK;
  const int rc = qmk_engine_sync_status(it->second, stream, nullptr);
  g_legacy_status = rc;
  return rc;
}

extern "C" void qmk_legacy_release(void) {
  std::lock_guard<std::mutex> lock(g_legacy_mu);
  for (auto& kv : g_models) qmk_model_destroy(kv.second.m);
  g_models.clear();
  for (auto& kv : g_engines) qmk_engine_destroy(kv.second);
  g_engines.clear();
  g_cfg.clear();
}

extern "C" void launch_ldg_decode_direct(int input_token_id, int* output_token_id, const void* embed_weight,
                                         const LDGLayerWeights* layer_weights, const void* final_norm_weight,
                                         const void* lm_head_weight, const void* cos_table, const void* sin_table,
                                         void* k_cache, void* v_cache, void* hidden_buffer, void* /*g_activations*/,
                                         void* /*g_residual*/, void* /*g_q*/, void* /*g_k*/, void* /*g_v*/,
                                         void* /*g_attn_out*/, void* /*g_mlp_intermediate*/, void* g_normalized,
                                         void* /*block_max_vals*/, void* /*block_max_idxs*/, int num_layers,
                                         int position, int max_seq_len, float attn_scale, void* stream) {
  std::lock_guard<std::mutex> lock(g_legacy_mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    g_legacy_status = set_error(QMK_ERR_CUDA, "launch_ldg_decode_direct: no CUDA device");
    return;
  }
  if (!layer_weights || num_layers < 1 || num_layers > 1024) {
    g_legacy_status = set_error(QMK_ERR_ARG, "launch_ldg_decode_direct: bad layer blob / num_layers %d", num_layers);
    return;
  }
  qmk_engine*& e = g_engines[dev];
  if (!e) {
    int rc = qmk_engine_create(dev, 0, &e);
    if (rc != QMK_OK) {
      g_engines.erase(dev);
      g_legacy_status = rc;
      fprintf(stderr, "[qmk] %s\n", qmk_last_error());
      return;
    }
  }
  LegacyCfg cfg;
  auto ci = g_cfg.find(layer_weights);
  if (ci != g_cfg.end()) cfg = ci->second;
  // Which upstream PyTorch path a blob is judged against: the 5-layer stack is the code predictor (bf16 residual stream,
  // model_tts.py:567-619), anything else the talker (fp32 residual, validate_kernel.py:123-188); qmk_legacy_configure overrides.
  const int residual_fp32 = cfg.residual_fp32 >= 0 ? cfg.residual_fp32 : (num_layers == 5 ? 0 : 1);
  LegacyModel& lm = g_models[std::make_tuple(dev, (const void*)layer_weights, num_layers)];
  if (!lm.m || lm.final_norm != final_norm_weight || lm.residual_fp32 != residual_fp32) {
    if (lm.m) qmk_model_destroy(lm.m);
    lm = LegacyModel();
    int rc = qmk_model_create(e, layer_weights, num_layers, final_norm_weight, residual_fp32, stream, &lm.m);
    if (rc == QMK_OK) {   // remember which pointers were packed (the create call has synchronised the stream already)
      lm.blob.resize((size_t)num_layers * 11);
      if (cudaMemcpy(lm.blob.data(), layer_weights, lm.blob.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_error(QMK_ERR_CUDA, "launch_ldg_decode_direct: blob copy failed");
    }
    if (rc != QMK_OK) {
      if (lm.m) qmk_model_destroy(lm.m);
      g_models.erase(std::make_tuple(dev, (const void*)layer_weights, num_layers));
      g_legacy_status = rc;
      fprintf(stderr, "[qmk] %s\n", qmk_last_error());
      return;
    }
    lm.final_norm = final_norm_weight;
    lm.residual_fp32 = residual_fp32;
  }
  int head = -1;
  if (cfg.lm_head_rows != 0 && lm_head_weight) {
    auto hi = lm.heads.find(lm_head_weight);
    if (hi == lm.heads.end()) {
      const int rows = cfg.lm_head_rows > 0 ? cfg.lm_head_rows : 3072;   // upstream compile-time LDG_VOCAB_SIZE (build_tts.py:47)
      bool zero = false;
      int rc = cfg.lm_head_rows > 0 ? QMK_OK : table_is_zero(lm_head_weight, (size_t)rows * H * 2, (cudaStream_t)stream, &zero);
      if (rc == QMK_OK && zero) rc = -1;   // dummy table: no head; the argmax of all-zero logits is token 0
      else if (rc == QMK_OK) rc = qmk_model_add_head(lm.m, lm_head_weight, rows, stream);
      if (rc < 0 && !zero) {
        g_legacy_status = rc;
        fprintf(stderr, "[qmk] %s\n", qmk_last_error());
        return;
      }
      hi = lm.heads.emplace(lm_head_weight, zero ? -1 : rc).first;
    }
    head = hi->second;
  }
  if (head < 0 && output_token_id) cudaMemsetAsync(output_token_id, 0, sizeof(int), (cudaStream_t)stream);
  g_legacy_status = qmk_decode_step(lm.m, head, input_token_id, embed_weight, cos_table, sin_table, k_cache, v_cache,
                                    hidden_buffer, reinterpret_cast<float*>(g_normalized), output_token_id, position,
                                    max_seq_len, attn_scale, 0, stream);
  if (g_legacy_status != QMK_OK) fprintf(stderr, "[qmk] %s\n", qmk_last_error());
}

// Compare the blob a cached model was packed from with the caller's current blob (HOST copy of the 11 * L pointers):
// returns 1 if they differ (the model is dropped and re-packed on the next call), 0 if equal or unknown.  The Python op
// calls this with the blob tensor's identity / version so that a recycled allocator address cannot alias a stale model.
extern "C" int qmk_legacy_check_blob(const LDGLayerWeights* layer_weights, int num_layers, const uint64_t* host_blob) {
  if (!layer_weights || !host_blob) return 0;
  std::lock_guard<std::mutex> lock(g_legacy_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  auto it = g_models.find(std::make_tuple(dev, (const void*)layer_weights, num_layers));
  if (it == g_models.end() || it->second.blob.empty()) return 0;
  if (memcmp(it->second.blob.data(), host_blob, it->second.blob.size() * 8) == 0) return 0;
  qmk_model_destroy(it->second.m);
  g_models.erase(it);
  return 1;
}
