// qmk_device.cuh — device side of the B200 decode engine (sm_100a).
//
// One persistent cooperative kernel, one CTA per SM.  Inside a CTA:
//   * warp 12 (one elected lane) is the PRODUCER: it streams this CTA's slice of the re-packed weights
//     with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) into an 8-slot x 24 KB
//     shared-memory ring.  Weight addresses do not depend on activations, so the producer runs ahead
//     through every data dependency of the layer and HBM stays busy while consumers synchronise.
//   * warps 0-11 are CONSUMERS: one 1024-element row segment (2 KB) per warp per stage, activations in
//     registers, 128-bit conflict-free shared loads, fp32 FMA, warp-shuffle reduction.
// CTAs exchange activations through "LL" words in global memory: a 64-bit store carries a 32-bit payload
// and a 32-bit epoch, so data and flag arrive in one single-copy-atomic access and a consumer simply
// re-reads a word until its epoch matches.  There is no grid barrier, no fence on the critical path and
// no separate flag: one store->L2->load hop per dependency.
//
// Numerics follow the upstream *PyTorch* path, not upstream kernel.cu (see oracle/tts_oracle.py for the
// rounding points and the upstream file:line of each): bf16 rounding after every projection / norm /
// RoPE product / SwiGLU factor, fp32 accumulation, fp32 (talker) or bf16 (code predictor) residual.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmk {

typedef unsigned long long u64;

constexpr int H = 1024, INTER = 3072, QSZ = 2048, KVSZ = 1024, HD = 128, NQH = 16, NKVH = 8;
constexpr int QKV_ROWS = QSZ + 2 * KVSZ;  // 4096
constexpr int SEG_ELEMS = 1024, SEG_BYTES = 2048;
constexpr int NCW = 12;              // consumer warps
constexpr int NCT = NCW * 32;        // 384 consumer threads
constexpr int NTHREADS = NCT + 32;   // + producer warp
constexpr int STAGE_SEGS = 12;
constexpr int STAGE_BYTES = STAGE_SEGS * SEG_BYTES;  // 24576
constexpr int NSTAGES = 8;
constexpr int ATT_PER_WARP = 5;
constexpr int ATT_ROUND = NCW * ATT_PER_WARP;  // 60 cached positions per round per item
constexpr int S_MAX = 18;                      // max KV splits per kv head (8 * 18 = 144 CTAs)
constexpr int PART_STRIDE = 132;               // words per (q head, split) partial: m, l, acc[128], pad
constexpr int MAX_HEAD_ROWS = 3072;
constexpr float EPS = 1e-6f;

enum Phase { PH_QKV = 0, PH_ATTN = 1, PH_O = 2, PH_GU = 3, PH_DOWN = 4, PH_PER_LAYER = 5 };
enum Status { ST_OK = 0, ST_TIMEOUT_LL = 1, ST_TIMEOUT_FULL = 2, ST_TIMEOUT_EMPTY = 3, ST_BAD_CONFIG = 4 };

// Exchange-buffer word counts
constexpr int XW_RES = H, XW_QKV = QKV_ROWS, XW_A = QSZ / 2, XW_RES2 = H, XW_M = INTER,
              XW_LOGITS = MAX_HEAD_ROWS, XW_PART = NQH * S_MAX * PART_STRIDE;
constexpr int XW_TOTAL = XW_RES + XW_QKV + XW_A + XW_RES2 + XW_M + XW_LOGITS + XW_PART;

// Segment layout of one (cta, layer) block of the packed weight stream, in 2 KB segments:
//   [aux: input_layernorm][qkv rows] | [o items (row,kb) kb<2] | [aux: post_ln][gate_j, up_j pairs] |
//   [down items (row,kb) kb<3]
struct Layout {
  int G;        // CTAs (= SMs)
  int L;        // layers
  int qkv_max;  // ceil(4096 / G)
  int o_max;    // ceil(1024 / G)
  int gu_max;   // ceil(3072 / G)
  int off_o, off_gu, off_down, layer_segs;
};

struct HeadDesc {
  const uint8_t* packed;  // [G][segs_max][2048]: per CTA [aux: final norm][rows]; rows==0 -> single aux segment
  int rows;
  int segs_max;
};

struct Params {
  Layout lay;
  const uint8_t* packed_layers;    // [G][L][layer_segs][2048]
  const __nv_bfloat16* qk_norm;    // [L][2][128]
  HeadDesc head;
  const __nv_bfloat16* embed;      // [vocab][1024] (token >= 0)
  const __nv_bfloat16* in_vec;     // bf16[1024]     (token < 0: upstream sentinel path)
  const __nv_bfloat16* cos_t;      // [max_seq][128]
  const __nv_bfloat16* sin_t;
  __nv_bfloat16* k_cache;          // [L][8][max_seq][128]
  __nv_bfloat16* v_cache;
  int max_seq;
  u64* xbuf;                       // exchange words, XW_TOTAL
  int token;
  int position;
  int* out_token;
  float* out_norm;                 // f32[1024]
  __nv_bfloat16* hidden_out;       // bf16[1024]
  float attn_scale;
  int residual_fp32;
  uint32_t epoch_base;
  int* status;                     // int[4]: code, cta, phase index, aux
  int phase_begin, phase_end;      // half-open range in the linear phase index space
  long long timeout_cycles;
  int replicas;                    // R copies of every exchange buffer; CTA c polls copy c % R
  int probe;                       // 1: one warp probes a sample of words before the CTA-wide gather
  long long* trace;                // optional [G][trace_stride] clock64 stamps at phase starts
  int trace_stride;
};

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int row_begin(int cta, int n_rows, int G) {
  return (int)(((long long)cta * n_rows) / G);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t bf16_bits(float x) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mbarrier / TMA bulk copy (PTX ISA: mbarrier, cp.async.bulk)
__device__ __forceinline__ void mbar_init(u64* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u64* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, u64* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// LL words: {payload (low 32), epoch (high 32)} moved with one 64-bit relaxed gpu-scope access.
__device__ __forceinline__ void ll_st(u64* p, uint32_t payload, uint32_t epoch) {
  u64 v = ((u64)epoch << 32) | payload;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ll_ld(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_cg_u2(const void* p) {
  uint2 v;
  asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
// Publish one exchange word into every replica (stride XW_TOTAL words).
__device__ __forceinline__ void ll_pub(u64* p, uint32_t payload, uint32_t epoch, int replicas, int stride_words) {
  for (int r = 0; r < replicas; ++r) ll_st(p + (size_t)r * stride_words, payload, epoch);
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up
// ------------------------------------------------------------------------------------------------
constexpr int SM_RING = 0;
constexpr int SM_VEC = SM_RING + NSTAGES * STAGE_BYTES;   // float[3072]  (aliased: attention merge acc [12][2][128])
constexpr int SM_SMALL = SM_VEC + INTER * 4;              // float[1024]  attention scratch
constexpr int SM_PART = SM_SMALL + 1024 * 4;              // float[128]   per-item partial dot products
constexpr int SM_BAR = SM_PART + 128 * 4;                 // u64 full[8], empty[8]
constexpr int SM_MISC = SM_BAR + 2 * NSTAGES * 8;         // int abort; int pad[3]; float red[..]
constexpr int SMEM_BYTES = SM_MISC + 256;

// s_small sub-offsets (floats)
constexpr int SS_QRAW = 0, SS_KRAW = 256, SS_V = 384, SS_QN = 512, SS_KN = 768, SS_ML = 896;  // ML: [12][2][2]

struct Ctx {
  const Params& p;
  uint8_t* ring;
  float* s_vec;
  float* s_small;
  float* s_part;
  u64* full;
  u64* empty;
  volatile int* s_abort;
  float* s_red;
  const u64* xrd;  // exchange replica this CTA reads
  int tid, warp, lane, cta;
  uint32_t k;     // stage counter (same sequence in producer and consumers)
  long long t0;
  int cur_idx;
  __device__ Ctx(const Params& pp) : p(pp) {}
};

// Watchdog: every spin loop funnels through here.  Out-of-line and by-value so that Ctx stays in registers.
__device__ __noinline__ bool check_abort_slow(int* status, volatile int* s_abort, long long t0, long long timeout,
                                              int cta, int cur_idx, int code, int aux) {
  if (*s_abort) return true;
  if (*((volatile int*)status) != 0) {
    *s_abort = 1;
    return true;
  }
  if (clock64() - t0 > timeout) {
    if (atomicCAS(status, 0, code) == 0) {
      status[1] = cta;
      status[2] = cur_idx;
      status[3] = aux;
      __threadfence();
    }
    *s_abort = 1;
    return true;
  }
  return false;
}
__device__ __forceinline__ bool check_abort(Ctx& c, int code, int aux) {
  return check_abort_slow(c.p.status, c.s_abort, c.t0, c.p.timeout_cycles, c.cta, c.cur_idx, code, aux);
}

__device__ __forceinline__ void wait_full(Ctx& c, uint32_t k) {
  u64* bar = &c.full[k % NSTAGES];
  uint32_t parity = (k / NSTAGES) & 1u;
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 63u) == 0 && check_abort(c, ST_TIMEOUT_FULL, (int)k)) return;
  }
}
__device__ __forceinline__ void wait_empty(Ctx& c, uint32_t k) {
  u64* bar = &c.empty[k % NSTAGES];
  uint32_t parity = ((k / NSTAGES) & 1u) ^ 1u;
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 63u) == 0 && check_abort(c, ST_TIMEOUT_EMPTY, (int)k)) return;
  }
}
__device__ __forceinline__ u64 ll_wait_slow(Ctx& c, const u64* p, uint32_t epoch) {
  uint32_t spins = 0;
  for (;;) {
    u64 w = ll_ld(p);
    if ((uint32_t)(w >> 32) == epoch) return w;
    if ((++spins & 127u) == 0 && check_abort(c, ST_TIMEOUT_LL, (int)(p - c.p.xbuf))) return w;
  }
}
__device__ __forceinline__ uint32_t ll_wait(Ctx& c, const u64* p, uint32_t epoch) {
  u64 w = ll_ld(p);
  if ((uint32_t)(w >> 32) != epoch) w = ll_wait_slow(c, p, epoch);
  return (uint32_t)w;
}

// Debug trace: trace[cta][(idx - phase_begin) * 8 + sub] = clock64() (thread 0 of the CTA only).
__device__ __forceinline__ void trace_sub(const Ctx& c, int sub) {
  if (c.p.trace != nullptr && c.tid == 0) {
    const int slot = (c.cur_idx - c.p.phase_begin) * 8 + sub;
    if (slot < c.p.trace_stride) c.p.trace[(size_t)c.cta * c.p.trace_stride + slot] = clock64();
  }
}

// Gather N exchange words (all loads issued before the first check) and hand each payload to `store`.
// With p.probe, warp 0 first polls a 32-word sample while the other warps sleep at a barrier: thousands
// of threads spinning on the same few L2 lines delay the very stores they are waiting for.
template <int N, typename F>
__device__ __forceinline__ void gather_words(Ctx& c, const u64* buf, uint32_t epoch, F store, int n = N) {
  constexpr int PER = (N + NCT - 1) / NCT;
  if (c.p.probe) {
    if (c.warp == 0) {
      const int idx = (int)(((long long)c.lane * n) >> 5) + (n >> 6);
      (void)ll_wait(c, buf + (idx < n ? idx : n - 1), epoch);
    }
    consumer_bar();
  }
  u64 w[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    int idx = u * NCT + c.tid;
    if (idx < n) w[u] = ll_ld(buf + idx);
  }
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    int idx = u * NCT + c.tid;
    if (idx < n) {
      if ((uint32_t)(w[u] >> 32) != epoch) w[u] = ll_wait_slow(c, buf + idx, epoch);
      store(idx, (uint32_t)w[u]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// phase descriptors (identical arithmetic in producer and consumers)
// ------------------------------------------------------------------------------------------------
struct PhaseDesc {
  const uint8_t* src;
  int total_segs;  // aux + items
  int has_aux;
  int n_items;
};

__device__ __forceinline__ PhaseDesc layer_phase_desc(const Params& p, int l, int ph, int cta) {
  const Layout& y = p.lay;
  PhaseDesc d;
  const uint8_t* base = p.packed_layers + ((size_t)((size_t)cta * y.L + l) * y.layer_segs) * SEG_BYTES;
  if (ph == PH_QKV) {
    d.n_items = row_begin(cta + 1, QKV_ROWS, y.G) - row_begin(cta, QKV_ROWS, y.G);
    d.has_aux = 1;
    d.src = base;
  } else if (ph == PH_O) {
    d.n_items = 2 * (row_begin(cta + 1, H, y.G) - row_begin(cta, H, y.G));
    d.has_aux = 0;
    d.src = base + (size_t)y.off_o * SEG_BYTES;
  } else if (ph == PH_GU) {
    d.n_items = 2 * (row_begin(cta + 1, INTER, y.G) - row_begin(cta, INTER, y.G));
    d.has_aux = 1;
    d.src = base + (size_t)y.off_gu * SEG_BYTES;
  } else if (ph == PH_DOWN) {
    d.n_items = 3 * (row_begin(cta + 1, H, y.G) - row_begin(cta, H, y.G));
    d.has_aux = 0;
    d.src = base + (size_t)y.off_down * SEG_BYTES;
  } else {
    d.n_items = 0;
    d.has_aux = 0;
    d.src = base;
  }
  d.total_segs = d.n_items + d.has_aux;
  return d;
}
__device__ __forceinline__ PhaseDesc head_phase_desc(const Params& p, int cta) {
  PhaseDesc d;
  d.has_aux = 1;
  if (p.head.rows > 0) {
    d.n_items = row_begin(cta + 1, p.head.rows, p.lay.G) - row_begin(cta, p.head.rows, p.lay.G);
    d.src = p.head.packed + ((size_t)cta * p.head.segs_max) * SEG_BYTES;
  } else {
    d.n_items = 0;
    d.src = p.head.packed;  // single aux segment shared by all CTAs
  }
  d.total_segs = d.n_items + 1;
  return d;
}
__device__ __forceinline__ int n_stages_of(const PhaseDesc& d) { return (d.total_segs + STAGE_SEGS - 1) / STAGE_SEGS; }

// ------------------------------------------------------------------------------------------------
// producer
// ------------------------------------------------------------------------------------------------
__device__ void producer_loop(Ctx& c) {
  const Params& p = c.p;
  const int nlayer_idx = p.lay.L * PH_PER_LAYER;
  for (int idx = p.phase_begin; idx < p.phase_end; ++idx) {
    c.cur_idx = idx;
    PhaseDesc d;
    if (idx < nlayer_idx) {
      int ph = idx % PH_PER_LAYER;
      if (ph == PH_ATTN) continue;
      d = layer_phase_desc(p, idx / PH_PER_LAYER, ph, c.cta);
    } else if (idx == nlayer_idx) {
      d = head_phase_desc(p, c.cta);
    } else {
      continue;
    }
    const int nst = n_stages_of(d);
    for (int s = 0; s < nst; ++s, ++c.k) {
      const int slot = c.k % NSTAGES;
      wait_empty(c, c.k);
      int segs = d.total_segs - s * STAGE_SEGS;
      if (segs > STAGE_SEGS) segs = STAGE_SEGS;
      const uint32_t bytes = (uint32_t)segs * SEG_BYTES;
      mbar_arrive_expect_tx(&c.full[slot], bytes);
      tma_bulk_g2s(c.ring + (size_t)slot * STAGE_BYTES, d.src + (size_t)s * STAGE_BYTES, bytes, &c.full[slot]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// consumer building blocks
// ------------------------------------------------------------------------------------------------
// 32 activations per lane: element (j, e) <-> k = j*256 + lane*8 + e of the warp's 1024-wide segment.
__device__ __forceinline__ void load_xr(const float* vec1024, int lane, float (&xr)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4* v = reinterpret_cast<const float4*>(vec1024 + j * 256 + lane * 8);
    float4 a = v[0], b = v[1];
    xr[j * 8 + 0] = a.x; xr[j * 8 + 1] = a.y; xr[j * 8 + 2] = a.z; xr[j * 8 + 3] = a.w;
    xr[j * 8 + 4] = b.x; xr[j * 8 + 5] = b.y; xr[j * 8 + 6] = b.z; xr[j * 8 + 7] = b.w;
  }
}

// xr <- r( r(x) / sqrt(mean(r(x)^2) + eps) * w ), w = bf16[1024] laid out like a weight segment.
__device__ __forceinline__ void rmsnorm_regs(const float* vec1024, const uint4* w_seg, int lane, float (&xr)[32]) {
  load_xr(vec1024, lane, xr);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    xr[i] = bf16_round(xr[i]);
    ss = fmaf(xr[i], xr[i], ss);
  }
  ss = warp_sum(ss);
  const float rms = sqrtf(ss * (1.0f / H) + EPS);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 wv = w_seg[j * 32 + lane];
    uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      xr[j * 8 + 2 * q] = bf16_round((xr[j * 8 + 2 * q] / rms) * bf16_lo(ww[q]));
      xr[j * 8 + 2 * q + 1] = bf16_round((xr[j * 8 + 2 * q + 1] / rms) * bf16_hi(ww[q]));
    }
  }
}

__device__ __forceinline__ float seg_dot(const uint4* w_seg, int lane, const float (&xr)[32]) {
  float acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 wv = w_seg[j * 32 + lane];
    float a = bf16_lo(wv.x) * xr[j * 8 + 0];
    a = fmaf(bf16_hi(wv.x), xr[j * 8 + 1], a);
    a = fmaf(bf16_lo(wv.y), xr[j * 8 + 2], a);
    a = fmaf(bf16_hi(wv.y), xr[j * 8 + 3], a);
    a = fmaf(bf16_lo(wv.z), xr[j * 8 + 4], a);
    a = fmaf(bf16_hi(wv.z), xr[j * 8 + 5], a);
    a = fmaf(bf16_lo(wv.w), xr[j * 8 + 6], a);
    a = fmaf(bf16_hi(wv.w), xr[j * 8 + 7], a);
    acc[j] = a;
  }
  return warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
}

// Consume all stages of a phase: warp w owns segment 12*s + w of stage s; partial dot -> s_part[item].
__device__ __forceinline__ void run_stages(Ctx& c, const PhaseDesc& d, const float (&xr)[32]) {
  const int nst = n_stages_of(d);
  for (int s = 0; s < nst; ++s, ++c.k) {
    const int slot = c.k % NSTAGES;
    wait_full(c, c.k);
    const int seg = s * STAGE_SEGS + c.warp;
    if (seg < d.total_segs && seg >= d.has_aux) {
      const uint4* w = reinterpret_cast<const uint4*>(c.ring + (size_t)slot * STAGE_BYTES + c.warp * SEG_BYTES);
      float v = seg_dot(w, c.lane, xr);
      if (c.lane == 0) c.s_part[seg - d.has_aux] = v;
    }
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.empty[slot]);
  }
}

__device__ __forceinline__ const __nv_bfloat16* step_input(const Params& p) {
  return p.token >= 0 ? p.embed + (size_t)p.token * H : p.in_vec;
}

// ------------------------------------------------------------------------------------------------
// attention work items: (kv head g, split s).  Item (g, s) runs on CTA s*8 + g.
// ------------------------------------------------------------------------------------------------
struct AttnItem {
  int g, s, S, p0, p1;
  bool owner;  // holds the new position (= p.position) -> takes k,v from the exchange and appends to the cache
};
__device__ __forceinline__ bool attn_item(const Params& p, int cta, AttnItem& it) {
  const int n = p.position + 1;
  int smax = p.lay.G / NKVH;
  if (smax > S_MAX) smax = S_MAX;
  int S0 = (n + ATT_ROUND - 1) / ATT_ROUND;
  if (S0 > smax) S0 = smax;
  const int C = (n + S0 - 1) / S0;
  const int S = (n + C - 1) / C;
  if (cta >= NKVH * S) return false;
  it.g = cta % NKVH;
  it.s = cta / NKVH;
  it.S = S;
  it.p0 = it.s * C;
  it.p1 = it.p0 + C < n ? it.p0 + C : n;
  it.owner = (it.p1 == n);
  return true;
}

struct KvRegs {
  uint2 k[ATT_PER_WARP];
  uint2 v[ATT_PER_WARP];
};
__device__ __forceinline__ void attn_prefetch(const Ctx& c, int l, const AttnItem& it, int round, KvRegs& r) {
  const Params& p = c.p;
  const size_t base = ((size_t)(l * NKVH + it.g) * p.max_seq) * HD;
#pragma unroll
  for (int i = 0; i < ATT_PER_WARP; ++i) {
    const int pos = it.p0 + round * ATT_ROUND + c.warp + NCW * i;
    if (pos < it.p1 && pos != p.position) {
      const size_t off = base + (size_t)pos * HD + c.lane * 4;
      r.k[i] = ld_cg_u2(p.k_cache + off);
      r.v[i] = ld_cg_u2(p.v_cache + off);
    }
  }
}

__device__ void phase_attn(Ctx& c, int l, uint32_t epoch, const AttnItem& it, KvRegs& kv, bool prefetched) {
  const Params& p = c.p;
  const u64* x_qkv = c.xrd + XW_RES;                  // this CTA's replica
  u64* x_a = p.xbuf + XW_RES + XW_QKV;                // writers publish into every replica
  u64* x_part = p.xbuf + (XW_TOTAL - XW_PART);        // split partials: replica 0 only (few readers)
  float* s_small = c.s_small;
  const int R = p.replicas;
  if (!prefetched) attn_prefetch(c, l, it, 0, kv);

  // 1) raw q (2 heads), k, v of this kv group
  {
    const int t = c.tid;
    if (t < 256) {
      s_small[SS_QRAW + t] = __uint_as_float(ll_wait(c, x_qkv + (2 * it.g) * HD + t, epoch));
    } else if (it.owner) {
      const int d = t - 256;  // 0..127
      s_small[SS_KRAW + d] = __uint_as_float(ll_wait(c, x_qkv + QSZ + it.g * HD + d, epoch));
      s_small[SS_V + d] = __uint_as_float(ll_wait(c, x_qkv + QSZ + KVSZ + it.g * HD + d, epoch));
    }
  }
  consumer_bar();
  trace_sub(c, 1);

  // 2) per-head RMSNorm + rotate-half RoPE in bf16 steps (warp 0,1: q heads; warp 2: k)
  if (c.warp < 2 || (c.warp == 2 && it.owner)) {
    const float* raw = (c.warp < 2) ? s_small + SS_QRAW + c.warp * HD : s_small + SS_KRAW;
    float* dst = (c.warp < 2) ? s_small + SS_QN + c.warp * HD : s_small + SS_KN;
    const __nv_bfloat16* wn = p.qk_norm + ((size_t)l * 2 + (c.warp < 2 ? 0 : 1)) * HD;
    float t[4], ss = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      t[e] = raw[c.lane * 4 + e];
      ss = fmaf(t[e], t[e], ss);
    }
    ss = warp_sum(ss);
    const float rms = sqrtf(ss * (1.0f / HD) + EPS);
    const int dbase = (c.lane * 4) & 63;
    const __nv_bfloat16* cr = p.cos_t + (size_t)p.position * HD + dbase;
    const __nv_bfloat16* sr = p.sin_t + (size_t)p.position * HD + dbase;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float n = bf16_round((t[e] / rms) * __bfloat162float(wn[c.lane * 4 + e]));
      const float o = __shfl_xor_sync(0xffffffffu, n, 16);
      const float cs = __bfloat162float(cr[e]), sn = __bfloat162float(sr[e]);
      const float a = bf16_round(n * cs), b = bf16_round(o * sn);
      dst[c.lane * 4 + e] = bf16_round(c.lane < 16 ? a - b : a + b);
    }
  }
  consumer_bar();
  trace_sub(c, 2);

  // 3) scores / online softmax / PV over this item's positions; lane owns dims 4*lane..4*lane+3
  float q0[4], q1[4], acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    q0[e] = s_small[SS_QN + c.lane * 4 + e];
    q1[e] = s_small[SS_QN + HD + c.lane * 4 + e];
  }
  const int len = it.p1 - it.p0;
  const int nrounds = (len + ATT_ROUND - 1) / ATT_ROUND;
  for (int r = 0; r < nrounds; ++r) {
    if (r > 0) attn_prefetch(c, l, it, r, kv);
    float sc0[ATT_PER_WARP], sc1[ATT_PER_WARP];
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int i = 0; i < ATT_PER_WARP; ++i) {
      const int pos = it.p0 + r * ATT_ROUND + c.warp + NCW * i;
      const bool valid = pos < it.p1;
      float kf[4];
      if (valid && pos == p.position) {
#pragma unroll
        for (int e = 0; e < 4; ++e) kf[e] = s_small[SS_KN + c.lane * 4 + e];
      } else {
        kf[0] = bf16_lo(kv.k[i].x); kf[1] = bf16_hi(kv.k[i].x);
        kf[2] = bf16_lo(kv.k[i].y); kf[3] = bf16_hi(kv.k[i].y);
      }
      float d0 = 0.f, d1 = 0.f;
      if (valid) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          d0 = fmaf(q0[e], kf[e], d0);
          d1 = fmaf(q1[e], kf[e], d1);
        }
      }
      d0 = warp_sum(d0);
      d1 = warp_sum(d1);
      sc0[i] = valid ? d0 * p.attn_scale : -INFINITY;
      sc1[i] = valid ? d1 * p.attn_scale : -INFINITY;
      mx0 = fmaxf(mx0, sc0[i]);
      mx1 = fmaxf(mx1, sc1[i]);
    }
    if (mx0 != -INFINITY) {  // warp-uniform (scores are warp-reduced)
      const float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - mx0);
      const float c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - mx1);
      l0 *= c0; l1 *= c1;
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc0[e] *= c0; acc1[e] *= c1; }
#pragma unroll
      for (int i = 0; i < ATT_PER_WARP; ++i) {
        const int pos = it.p0 + r * ATT_ROUND + c.warp + NCW * i;
        if (pos < it.p1) {
          float vf[4];
          if (pos == p.position) {
#pragma unroll
            for (int e = 0; e < 4; ++e) vf[e] = s_small[SS_V + c.lane * 4 + e];
          } else {
            vf[0] = bf16_lo(kv.v[i].x); vf[1] = bf16_hi(kv.v[i].x);
            vf[2] = bf16_lo(kv.v[i].y); vf[3] = bf16_hi(kv.v[i].y);
          }
          const float e0 = __expf(sc0[i] - mx0), e1 = __expf(sc1[i] - mx1);
          l0 += e0; l1 += e1;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc0[e] = fmaf(e0, vf[e], acc0[e]);
            acc1[e] = fmaf(e1, vf[e], acc1[e]);
          }
        }
      }
      m0 = mx0; m1 = mx1;
    }
  }
  trace_sub(c, 3);
  // 4) cross-warp merge through shared memory (s_vec region is free during attention)
  float* s_acc = c.s_vec;  // [12][2][128]
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    s_acc[(c.warp * 2 + 0) * HD + c.lane * 4 + e] = acc0[e];
    s_acc[(c.warp * 2 + 1) * HD + c.lane * 4 + e] = acc1[e];
  }
  if (c.lane == 0) {
    s_small[SS_ML + (c.warp * 2 + 0) * 2 + 0] = m0; s_small[SS_ML + (c.warp * 2 + 0) * 2 + 1] = l0;
    s_small[SS_ML + (c.warp * 2 + 1) * 2 + 0] = m1; s_small[SS_ML + (c.warp * 2 + 1) * 2 + 1] = l1;
  }
  consumer_bar();
  trace_sub(c, 4);
  if (c.tid < 128) {
    const int h = c.tid >> 6, dp = c.tid & 63;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < NCW; ++w) M = fmaxf(M, s_small[SS_ML + (w * 2 + h) * 2]);
    float Lsum = 0.f, A0 = 0.f, A1 = 0.f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) {
      const float mw = s_small[SS_ML + (w * 2 + h) * 2];
      const float f = (mw == -INFINITY) ? 0.f : __expf(mw - M);
      Lsum = fmaf(s_small[SS_ML + (w * 2 + h) * 2 + 1], f, Lsum);
      A0 = fmaf(s_acc[(w * 2 + h) * HD + 2 * dp], f, A0);
      A1 = fmaf(s_acc[(w * 2 + h) * HD + 2 * dp + 1], f, A1);
    }
    const int hq = 2 * it.g + h;
    if (it.S == 1) {
      const uint32_t pk = bf16_bits(A0 / Lsum) | (bf16_bits(A1 / Lsum) << 16);
      ll_pub(x_a + hq * 64 + dp, pk, epoch, R, XW_TOTAL);
    } else {
      u64* part = x_part + ((size_t)hq * S_MAX + it.s) * PART_STRIDE;
      ll_st(part + 2 + 2 * dp, __float_as_uint(A0), epoch);
      ll_st(part + 3 + 2 * dp, __float_as_uint(A1), epoch);
      if (dp == 0) {
        ll_st(part + 0, __float_as_uint(M), epoch);
        ll_st(part + 1, __float_as_uint(Lsum), epoch);
      }
      if (it.s == 0) {  // this CTA merges the splits of its two q heads, fixed order s = 0..S-1
        float Mx = -INFINITY;
        for (int s = 0; s < it.S; ++s) {
          const u64* ps = x_part + ((size_t)hq * S_MAX + s) * PART_STRIDE;
          Mx = fmaxf(Mx, __uint_as_float(ll_wait(c, ps, epoch)));
        }
        float Lt = 0.f, B0 = 0.f, B1 = 0.f;
        for (int s = 0; s < it.S; ++s) {
          const u64* ps = x_part + ((size_t)hq * S_MAX + s) * PART_STRIDE;
          const float f = __expf(__uint_as_float(ll_wait(c, ps, epoch)) - Mx);
          Lt = fmaf(__uint_as_float(ll_wait(c, ps + 1, epoch)), f, Lt);
          B0 = fmaf(__uint_as_float(ll_wait(c, ps + 2 + 2 * dp, epoch)), f, B0);
          B1 = fmaf(__uint_as_float(ll_wait(c, ps + 3 + 2 * dp, epoch)), f, B1);
        }
        const uint32_t pk = bf16_bits(B0 / Lt) | (bf16_bits(B1 / Lt) << 16);
        ll_pub(x_a + hq * 64 + dp, pk, epoch, R, XW_TOTAL);
      }
    }
  }
  trace_sub(c, 5);
  // 5) append the new K/V row (off the critical path: after `a` has been published)
  if (it.owner && c.tid < 128) {
    const size_t off = ((size_t)(l * NKVH + it.g) * p.max_seq + p.position) * HD + c.tid;
    p.k_cache[off] = __float2bfloat16_rn(s_small[SS_KN + c.tid]);
    p.v_cache[off] = __float2bfloat16_rn(s_small[SS_V + c.tid]);
    __threadfence();
  }
  consumer_bar();  // s_small / s_acc are reused by the next phase
  trace_sub(c, 6);
}

// ------------------------------------------------------------------------------------------------
// consumer main loop
// ------------------------------------------------------------------------------------------------
__device__ void consumer_loop(Ctx& c) {
  const Params& p = c.p;
  const Layout& y = p.lay;
  const int nlayer_idx = y.L * PH_PER_LAYER;
  // writer views (replica 0; ll_pub fans out) and reader views (replica cta % R)
  u64* w_res = p.xbuf;
  u64* w_qkv = w_res + XW_RES;
  u64* w_res2 = w_qkv + XW_QKV + XW_A;
  u64* w_m = w_res2 + XW_RES2;
  u64* w_logits = w_m + XW_M;
  const u64* x_res = c.xrd;
  const u64* x_a = x_res + XW_RES + XW_QKV;
  const u64* x_res2 = x_a + XW_A;
  const u64* x_m = x_res2 + XW_RES2;
  const u64* x_logits = x_m + XW_M;
  const int R = p.replicas;
  float xr[32];
  KvRegs kv;
  AttnItem item;
  const bool has_item = attn_item(p, c.cta, item);
  int prefetched_layer = -1;
  const int o_row0 = row_begin(c.cta, H, y.G);
  const int o_rows = row_begin(c.cta + 1, H, y.G) - o_row0;

  for (int idx = p.phase_begin; idx < p.phase_end; ++idx) {
    c.cur_idx = idx;
    trace_sub(c, 0);
    if (idx < nlayer_idx) {
      const int l = idx / PH_PER_LAYER, ph = idx % PH_PER_LAYER;
      const uint32_t epoch = p.epoch_base + 1u + (uint32_t)l;
      if (ph == PH_QKV) {
        const PhaseDesc d = layer_phase_desc(p, l, PH_QKV, c.cta);
        if (has_item) {  // KV rows of older positions do not depend on this layer: fetch them now
          attn_prefetch(c, l, item, 0, kv);
          prefetched_layer = l;
        }
        if (l == 0) {
          const __nv_bfloat16* x = step_input(p);
          for (int i = c.tid; i < H; i += NCT) c.s_vec[i] = __bfloat162float(x[i]);
        } else {
          gather_words<XW_RES>(c, x_res, epoch - 1u, [&](int i, uint32_t v) { c.s_vec[i] = __uint_as_float(v); });
        }
        consumer_bar();
        trace_sub(c, 1);
        wait_full(c, c.k);
        trace_sub(c, 2);
        rmsnorm_regs(c.s_vec, reinterpret_cast<const uint4*>(c.ring + (size_t)(c.k % NSTAGES) * STAGE_BYTES), c.lane, xr);
        trace_sub(c, 3);
        run_stages(c, d, xr);
        trace_sub(c, 4);
        consumer_bar();
        trace_sub(c, 5);
        if (c.tid < d.n_items) {
          const int row = row_begin(c.cta, QKV_ROWS, y.G) + c.tid;
          ll_pub(w_qkv + row, __float_as_uint(bf16_round(c.s_part[c.tid])), epoch, R, XW_TOTAL);
        }
      } else if (ph == PH_ATTN) {
        if (has_item) phase_attn(c, l, epoch, item, kv, prefetched_layer == l);
      } else if (ph == PH_O) {
        const PhaseDesc d = layer_phase_desc(p, l, PH_O, c.cta);
        float res_old = 0.f;
        if (c.tid < o_rows) {
          res_old = (l == 0) ? __bfloat162float(step_input(p)[o_row0 + c.tid])
                             : __uint_as_float(ll_wait(c, x_res + o_row0 + c.tid, epoch - 1u));
        }
        gather_words<XW_A>(c, x_a, epoch, [&](int i, uint32_t v) {
          c.s_vec[2 * i] = bf16_lo(v);
          c.s_vec[2 * i + 1] = bf16_hi(v);
        });
        consumer_bar();
        trace_sub(c, 1);
        load_xr(c.s_vec + (c.warp % 2) * SEG_ELEMS, c.lane, xr);
        run_stages(c, d, xr);
        trace_sub(c, 4);
        consumer_bar();
        trace_sub(c, 5);
        if (c.tid < o_rows) {
          const float o = bf16_round(c.s_part[2 * c.tid] + c.s_part[2 * c.tid + 1]);
          const float res = p.residual_fp32 ? res_old + o : bf16_round(res_old + o);
          ll_pub(w_res2 + o_row0 + c.tid, __float_as_uint(res), epoch, R, XW_TOTAL);
        }
      } else if (ph == PH_GU) {
        const PhaseDesc d = layer_phase_desc(p, l, PH_GU, c.cta);
        gather_words<XW_RES2>(c, x_res2, epoch, [&](int i, uint32_t v) { c.s_vec[i] = __uint_as_float(v); });
        consumer_bar();
        trace_sub(c, 1);
        wait_full(c, c.k);
        trace_sub(c, 2);
        rmsnorm_regs(c.s_vec, reinterpret_cast<const uint4*>(c.ring + (size_t)(c.k % NSTAGES) * STAGE_BYTES), c.lane, xr);
        trace_sub(c, 3);
        run_stages(c, d, xr);
        trace_sub(c, 4);
        consumer_bar();
        trace_sub(c, 5);
        if (c.tid < d.n_items / 2) {
          const float g = bf16_round(c.s_part[2 * c.tid]);
          const float u = bf16_round(c.s_part[2 * c.tid + 1]);
          const float sg = bf16_round(g / (1.0f + expf(-g)));
          const int pair = row_begin(c.cta, INTER, y.G) + c.tid;
          ll_pub(w_m + pair, __float_as_uint(bf16_round(sg * u)), epoch, R, XW_TOTAL);
        }
      } else {  // PH_DOWN
        const PhaseDesc d = layer_phase_desc(p, l, PH_DOWN, c.cta);
        float res_old = 0.f;
        if (c.tid < o_rows) res_old = __uint_as_float(ll_wait(c, x_res2 + o_row0 + c.tid, epoch));
        gather_words<XW_M>(c, x_m, epoch, [&](int i, uint32_t v) { c.s_vec[i] = __uint_as_float(v); });
        consumer_bar();
        trace_sub(c, 1);
        load_xr(c.s_vec + (c.warp % 3) * SEG_ELEMS, c.lane, xr);
        run_stages(c, d, xr);
        trace_sub(c, 4);
        consumer_bar();
        trace_sub(c, 5);
        if (c.tid < o_rows) {
          const float dn = bf16_round((c.s_part[3 * c.tid] + c.s_part[3 * c.tid + 1]) + c.s_part[3 * c.tid + 2]);
          const float res = p.residual_fp32 ? res_old + dn : bf16_round(res_old + dn);
          ll_pub(w_res + o_row0 + c.tid, __float_as_uint(res), epoch, R, XW_TOTAL);
        }
      }
    } else if (idx == nlayer_idx) {
      // final RMSNorm (+ LM head rows of this CTA)
      const uint32_t epoch_last = p.epoch_base + (uint32_t)y.L;
      const PhaseDesc d = head_phase_desc(p, c.cta);
      gather_words<XW_RES>(c, x_res, epoch_last, [&](int i, uint32_t v) { c.s_vec[i] = __uint_as_float(v); });
      consumer_bar();
      wait_full(c, c.k);
      if (c.cta == 0 && c.warp == 1 && p.hidden_out != nullptr) {
        for (int i = c.lane; i < H; i += 32) p.hidden_out[i] = __float2bfloat16_rn(c.s_vec[i]);
      }
      rmsnorm_regs(c.s_vec, reinterpret_cast<const uint4*>(c.ring + (size_t)(c.k % NSTAGES) * STAGE_BYTES), c.lane, xr);
      if (c.cta == 0 && c.warp == 0 && p.out_norm != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 8; ++e) p.out_norm[j * 256 + c.lane * 8 + e] = xr[j * 8 + e];
      }
      run_stages(c, d, xr);
      consumer_bar();
      if (c.tid < d.n_items) {
        const int row = row_begin(c.cta, p.head.rows, y.G) + c.tid;
        ll_pub(w_logits + row, __float_as_uint(bf16_round(c.s_part[c.tid])), epoch_last + 1u, R, XW_TOTAL);
      }
    } else {
      // argmax over the bf16 logits, lowest index wins ties (CTA 0)
      if (c.cta != 0 || p.head.rows <= 0) continue;
      const uint32_t epoch_head = p.epoch_base + (uint32_t)y.L + 1u;
      float best = -INFINITY;
      int best_i = 0x7fffffff;
      gather_words<XW_LOGITS>(c, x_logits, epoch_head, [&](int i, uint32_t u) {
        const float v = __uint_as_float(u);   // indices arrive in ascending order per thread
        if (v > best) { best = v; best_i = i; }
      }, p.head.rows);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
      }
      consumer_bar();
      if (c.lane == 0) {
        c.s_red[c.warp * 2] = best;
        c.s_red[c.warp * 2 + 1] = __int_as_float(best_i);
      }
      consumer_bar();
      if (c.tid == 0) {
        for (int w = 1; w < NCW; ++w) {
          const float ov = c.s_red[w * 2];
          const int oi = __float_as_int(c.s_red[w * 2 + 1]);
          if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        // a failed launch must not look like a token: encode the watchdog code as a negative id
        const int st = *((volatile int*)p.status);
        *p.out_token = (st != 0 || *c.s_abort) ? -1000 - st : best_i;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) qmk_decode_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctx c(p);
  c.ring = smem + SM_RING;
  c.s_vec = reinterpret_cast<float*>(smem + SM_VEC);
  c.s_small = reinterpret_cast<float*>(smem + SM_SMALL);
  c.s_part = reinterpret_cast<float*>(smem + SM_PART);
  c.full = reinterpret_cast<u64*>(smem + SM_BAR);
  c.empty = c.full + NSTAGES;
  c.s_abort = reinterpret_cast<volatile int*>(smem + SM_MISC);
  c.s_red = reinterpret_cast<float*>(smem + SM_MISC + 16);
  c.tid = threadIdx.x;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.cta = blockIdx.x;
  c.k = 0;
  c.t0 = clock64();
  c.cur_idx = p.phase_begin;
  c.xrd = p.xbuf + (size_t)(blockIdx.x % p.replicas) * XW_TOTAL;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGES; ++i) {
      mbar_init(&c.full[i], 1);
      mbar_init(&c.empty[i], NCW);
    }
    *c.s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (c.warp == NCW) {
    if (c.lane == 0) producer_loop(c);
  } else {
    consumer_loop(c);
  }
  __syncthreads();
  c.cur_idx = p.phase_end;
  trace_sub(c, 0);
  if (*c.s_abort && threadIdx.x == 0) {
    // failure path only: give in-flight bulk copies time to land before the CTA's shared memory is released
    const long long t = clock64();
    while (clock64() - t < 2000000) {
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight re-packing (one-time): upstream [out,in] row-major bf16 -> per-CTA segment streams
// ------------------------------------------------------------------------------------------------
struct LayerPtrs {  // = upstream LDGLayerWeights (kernel.cu:78-90)
  const uint4* w[11];
};
enum { W_IN = 0, W_Q, W_K, W_V, W_QN, W_KN, W_O, W_POST, W_GATE, W_UP, W_DOWN };

// grid = (G, L), block = 128 threads: thread t copies uint4 t of each 2 KB segment.
__global__ void qmk_pack_layers_kernel(const LayerPtrs* layers, Layout y, uint8_t* packed, __nv_bfloat16* qk_norm) {
  const int cta = blockIdx.x, l = blockIdx.y, t = threadIdx.x;
  const LayerPtrs lp = layers[l];
  uint4* dst = reinterpret_cast<uint4*>(packed + ((size_t)((size_t)cta * y.L + l) * y.layer_segs) * SEG_BYTES);
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const int q0 = row_begin(cta, QKV_ROWS, y.G), nq = row_begin(cta + 1, QKV_ROWS, y.G) - q0;
  const int o0 = row_begin(cta, H, y.G), no = row_begin(cta + 1, H, y.G) - o0;
  const int g0 = row_begin(cta, INTER, y.G), ng = row_begin(cta + 1, INTER, y.G) - g0;
  for (int seg = 0; seg < y.layer_segs; ++seg) {
    uint4 v = zero;
    if (seg < y.off_o) {
      const int s = seg;
      if (s == 0) {
        v = lp.w[W_IN][t];
      } else if (s - 1 < nq) {
        const int row = q0 + s - 1;
        if (row < QSZ) v = lp.w[W_Q][(size_t)row * 128 + t];
        else if (row < QSZ + KVSZ) v = lp.w[W_K][(size_t)(row - QSZ) * 128 + t];
        else v = lp.w[W_V][(size_t)(row - QSZ - KVSZ) * 128 + t];
      }
    } else if (seg < y.off_gu) {
      const int s = seg - y.off_o;
      if (s < 2 * no) v = lp.w[W_O][((size_t)(o0 + s / 2) * 2 + (s % 2)) * 128 + t];
    } else if (seg < y.off_down) {
      const int s = seg - y.off_gu;
      if (s == 0) {
        v = lp.w[W_POST][t];
      } else if (s - 1 < 2 * ng) {
        const int it = s - 1, row = g0 + it / 2;
        v = (it % 2 == 0) ? lp.w[W_GATE][(size_t)row * 128 + t] : lp.w[W_UP][(size_t)row * 128 + t];
      }
    } else {
      const int s = seg - y.off_down;
      if (s < 3 * no) v = lp.w[W_DOWN][((size_t)(o0 + s / 3) * 3 + (s % 3)) * 128 + t];
    }
    dst[(size_t)seg * 128 + t] = v;
  }
  if (cta == 0) {
    const __nv_bfloat16* qn = reinterpret_cast<const __nv_bfloat16*>(lp.w[W_QN]);
    const __nv_bfloat16* kn = reinterpret_cast<const __nv_bfloat16*>(lp.w[W_KN]);
    qk_norm[((size_t)l * 2 + 0) * HD + t] = qn[t];
    qk_norm[((size_t)l * 2 + 1) * HD + t] = kn[t];
  }
}

// grid = G, block = 128: per CTA [aux: final norm][rows of this CTA]
__global__ void qmk_pack_head_kernel(const uint4* head_w, const uint4* final_norm, int rows, int G, int segs_max,
                                     uint8_t* packed) {
  const int cta = blockIdx.x, t = threadIdx.x;
  uint4* dst = reinterpret_cast<uint4*>(packed + ((size_t)cta * segs_max) * SEG_BYTES);
  const int r0 = row_begin(cta, rows, G), n = row_begin(cta + 1, rows, G) - r0;
  for (int seg = 0; seg < segs_max; ++seg) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (seg == 0) v = final_norm[t];
    else if (seg - 1 < n) v = head_w[(size_t)(r0 + seg - 1) * 128 + t];
    dst[(size_t)seg * 128 + t] = v;
  }
}

}  // namespace qmk
