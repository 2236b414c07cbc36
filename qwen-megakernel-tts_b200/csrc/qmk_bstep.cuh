// qmk_bstep.cuh — persistent batched decode step (B = 1 .. 64 lanes) on tcgen05 / TMEM, sm_100a.
//
// Round 1 ran a batched step as a chain of ~230 dependent launches (one GEMM + one epilogue kernel per projection):
// 1.0-1.1 ms per step whatever B, bound by launch / drain latency (0.15-0.17 of the HBM roofline).  This kernel runs the WHOLE
// step -- every layer, the LM head and the argmax -- in ONE cooperative launch of one CTA per SM:
//   * a phase is a list of independent items (a 128-row weight tile x a K range for the projections; a (lane, kv head) pair
//     for attention; a lane for residual + RMSNorm ...); CTA c takes items c, c + G, ...; phases are separated by a grid barrier
//     (one atomic arrive + spin per CTA, ~1 us) instead of a kernel boundary (~5 us);
//   * the projections are tcgen05.mma (cta_group::1, kind::f16, M = 128, N = lanes, fp32 accumulators in TMEM); weights are
//     read in place in the upstream [out, in] layout through TMA tensor maps (SWIZZLE_128B), and stream through a 10-slot
//     shared-memory ring: a CTA's items are static, so the weight tiles of its NEXT projections are requested as soon as ring
//     slots are free -- HBM keeps streaming through the epilogue phases and the grid barriers;
//   * K is split across CTAs so that every projection has ~144 items (all SMs pull weights); the fp32 partials stay in L2 and
//     are summed by the phase that consumes them, which also carries the step's rounding points (bf16 after every projection,
//     per-head RMSNorm, RoPE, SwiGLU, residual) exactly like the B = 1 kernel.
// Two ways to use the lanes: DECODE (lane = independent utterance: own KV cache [B][L][8][S][128], own position) and PREFILL
// (lanes = consecutive positions of ONE utterance sharing a cache [L][8][S][128], causal: SURVEY.md section 8f row 4).
#pragma once

#include "qmk_bgemm.cuh"

namespace qmkb {

constexpr int H_ = 1024, INTER_ = 3072, QSZ_ = 2048, KVSZ_ = 1024, HD_ = 128, NKVH_ = 8;
constexpr int QKV_ROWS_ = 4096, GU_ROWS_ = 6144;
constexpr float EPS_ = 1e-6f;
constexpr int NT = 256;                 // consumer threads per CTA (8 warps)
constexpr int NT_ALL = NT + 32;         // + one producer warp that does nothing but request weight tiles
constexpr int NSLOT_A = 10;             // weight ring: 10 x 16 KB k-blocks (an item uses at most 6)
constexpr int NSLOT_B = 6;              // activation k-blocks of the current item: 6 x (N x 128 B <= 8 KB)
constexpr int PS_SMEM = NSLOT_A * A_TILE_BYTES + NSLOT_B * B_TILE_BYTES + 1024 /*alignment slack*/ + 512;
static_assert(PS_SMEM <= 232448, "shared memory budget");

enum PhaseKind { G_QKV = 0, G_O = 1, G_GU = 2, G_DOWN = 3, G_HEAD = 4 };

struct GemmShape { int M, KB, S; };   // rows, k-blocks of 64, K splits
// item i -> tile i / S, split i % S, k-blocks [split * KB / S, (split + 1) * KB / S).  QKV / gate-up / head: 128-144 items so that
// all SMs pull weights; O / down: 72 items of 3-6 k-blocks on the CTAs at the END of the grid (their weights are prefetched phases
// ahead, and every extra split is one more fp32 partial the consuming phase has to read per element).
__device__ __forceinline__ GemmShape gemm_shape(int kind, int head_rows) {
  switch (kind) {
    case G_QKV: return {QKV_ROWS_, 16, 4};
    case G_O: return {H_, 32, 9};
    case G_GU: return {GU_ROWS_, 16, 3};
    case G_DOWN: return {H_, 48, 9};
    default: return {head_rows, 16, 6};
  }
}
constexpr int S_QKV = 4, S_O = 9, S_GU = 3, S_DOWN = 9, S_HEAD = 6;

struct BStepParams {
  // Tensor maps travel as kernel parameters; one map per projection kind covers ALL layers, so eight descriptors serve the whole
  // step.  The weights are re-packed K-BLOCK-MAJOR at create time, [layer][k-block][row][64], so that a 128-row x 64-k tile is
  // 16 KB of CONTIGUOUS memory (in the upstream [out, in] layout it is 128 pieces of 128 B, 2-6 KB apart: measured ~19 GB/s per
  // SM through TMA, i.e. the projections ran at 2.8 TB/s aggregate).
  CUtensorMap map_w[5];             // 2-D views [L * KB * rows, 64] of qkv, o, gate/up, down and the head
  CUtensorMap map_x[3];             // activations: xn [N, 1024], abuf [N, 2048], mbuf [N, 3072]
  int L, B, N;                      // layers, lanes in use, UMMA N (B rounded up to 16)
  int head_rows, vocab;
  int residual_fp32;
  int prefill;                      // 1: lanes are consecutive positions of ONE utterance (shared cache, causal)
  int max_seq;
  float attn_scale;
  const __nv_bfloat16* const* ln_in;    // device arrays of L pointers
  const __nv_bfloat16* const* ln_post;
  const __nv_bfloat16* const* qn;
  const __nv_bfloat16* const* kn;
  const __nv_bfloat16* final_norm;
  const __nv_bfloat16* embed;
  const __nv_bfloat16* cos_t;
  const __nv_bfloat16* sin_t;
  const int* token_ids;             // int32[B] or null
  const __nv_bfloat16* embeds;      // bf16[B][1024] or null
  int* positions;                   // int32[B]; advanced by one at the end of the step (decode) -- prefill: positions[0] = first position
  __nv_bfloat16* k_cache;
  __nv_bfloat16* v_cache;
  float* hidden_out;                // f32[B][1024] or null
  int* tokens_out;                  // int32[B]
  float* res;                       // f32[B][1024] residual stream
  float* partial;                   // fp32 split-K partials
  float* qbuf;                      // f32[B][2048] normalised + rotated q (prefill: attention runs in its own phase)
  __nv_bfloat16* xn;                // bf16[N][1024]
  __nv_bfloat16* abuf;              // bf16[N][2048]
  __nv_bfloat16* mbuf;              // bf16[N][3072]
  unsigned* bar;                    // grid barrier counter (monotonic; the host passes the value it has before this launch)
  unsigned bar_base;
  int* status;                      // watchdog: set to 1 if a wait timed out
  long long timeout_cycles;
  long long* trace;                 // optional: CTA 0 records clock64() when it enters and leaves every grid barrier
};

__device__ __forceinline__ float b_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t b_bits(float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float b_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float b_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float w_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct BCtx {
  const BStepParams& p;
  uint8_t* sA;
  uint8_t* sB;
  uint64_t* fullA;      // [NSLOT_A]
  uint64_t* fullB;      // [NSLOT_B]
  uint64_t* done;
  float* s_red;         // [16]
  volatile int* s_abort;
  volatile uint32_t* s_consumed;   // weight k-blocks consumed so far (written by thread 0, polled by the producer warp)
  uint32_t tmem;
  int cta, G, tid, warp, lane;
  // weight stream (thread 0 only): next k-block to request / number consumed so far
  uint32_t a_issued, a_consumed;
  int pl, pph, pkb;     // producer cursor: layer, projection kind, k-block inside this CTA's item
  uint32_t b_uses;      // how many items have used the activation slots (parity)
  uint32_t done_uses;
  unsigned bar_target;
  int trace_n;
  long long t0;
  __device__ BCtx(const BStepParams& pp) : p(pp) {}
};

// barrier over the 8 consumer warps (the producer warp never joins)
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

__device__ __forceinline__ bool b_timed_out(BCtx& c) {
  if (*c.s_abort) return true;
  if (clock64() - c.t0 > c.p.timeout_cycles || *((volatile int*)c.p.status) != 0) {
    atomicExch(c.p.status, 1);
    *c.s_abort = 1;
    return true;
  }
  return false;
}
__device__ __forceinline__ void b_mbar_wait(BCtx& c, uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && (++spins & 1023u) == 0 && b_timed_out(c)) return;
  }
}

// ---- grid barrier: every CTA arrives once per phase; the counter only grows (the host tracks its value across launches) ----
__device__ __forceinline__ void grid_barrier(BCtx& c) {
  // writes of this phase (generic proxy) must be visible to TMA reads (async proxy) of other CTAs in the next phase
  asm volatile("fence.proxy.async;" ::: "memory");
  cta_sync();
  c.bar_target += (unsigned)c.G;
  if (c.p.trace != nullptr && c.cta == 0 && c.tid == 0) c.p.trace[c.trace_n++] = clock64();
  if (c.tid == 0) {
    __threadfence();
    atomicAdd(c.p.bar, 1u);
    unsigned spins = 0;
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c.p.bar) : "memory");
      if ((int)(v - c.bar_target) >= 0) break;
      if ((++spins & 255u) == 0 && b_timed_out(c)) break;
    }
    if (c.p.trace != nullptr && c.cta == 0) c.p.trace[c.trace_n++] = clock64();
  }
  cta_sync();
  asm volatile("fence.proxy.async;" ::: "memory");
}

// ---- weight stream ----------------------------------------------------------------------------------------------------------
// This CTA's k-block sequence: for every layer its QKV, O, gate/up and down item (if it has one), then its head item.
__device__ __forceinline__ bool item_range(const BStepParams& p, int kind, int cta, int G, int& tile, int& kb0, int& nkb, int& split) {
  const GemmShape g = gemm_shape(kind, p.head_rows);
  const int items = (g.M / BM) * g.S;
  if (kind == G_O || kind == G_DOWN) cta = G - 1 - cta;   // the 72-item projections sit at the end of the grid
  if (cta >= items) return false;
  tile = cta / g.S;
  split = cta % g.S;
  kb0 = (split * g.KB) / g.S;
  nkb = ((split + 1) * g.KB) / g.S - kb0;
  return true;
}
// Producer warp (lane 0): requests this CTA's weight k-blocks in consumption order, as far ahead as the ring allows, for the
// whole step.  It never joins the consumers' barriers: a TMA issue blocks the issuing thread while the copy engine's queue is
// full (measured: ~1.7 k cycles per 16 KB request under load), which would otherwise hold the whole CTA in front of a grid barrier.
__device__ __noinline__ void weights_producer(BCtx& c, volatile uint32_t* s_consumed) {
  const BStepParams& p = c.p;
  uint32_t spins = 0;
  while (c.pl <= p.L) {
    if (c.a_issued - *s_consumed >= (uint32_t)NSLOT_A) {
      __nanosleep(64);
      if ((++spins & 4095u) == 0 && b_timed_out(c)) return;
      continue;
    }
    const int kind = c.pl < p.L ? c.pph : G_HEAD;
    int tile, kb0, nkb, split;
    if (!item_range(p, kind, c.cta, c.G, tile, kb0, nkb, split) || c.pkb >= nkb) {
      // next item
      c.pkb = 0;
      if (c.pl < p.L && c.pph < G_DOWN) ++c.pph; else { c.pph = G_QKV; ++c.pl; }
      continue;
    }
    const GemmShape g = gemm_shape(kind, p.head_rows);
    const uint32_t slot = c.a_issued % NSLOT_A;
    mbar_expect_tx(&c.fullA[slot], (uint32_t)A_TILE_BYTES);
    tma_load_2d(c.sA + slot * A_TILE_BYTES, &p.map_w[kind], 0, ((c.pl < p.L ? c.pl : 0) * g.KB + kb0 + c.pkb) * g.M + tile * BM, &c.fullA[slot]);
    ++c.a_issued;
    ++c.pkb;
  }
}

// One projection phase: this CTA's item = 128 rows x k-blocks [kb0, kb0 + nkb) -> fp32 partial[split][n][row].
__device__ void gemm_phase(BCtx& c, int layer, int kind) {
  const BStepParams& p = c.p;
  int tile, kb0, nkb, split;
  if (!item_range(p, kind, c.cta, c.G, tile, kb0, nkb, split)) return;
  const GemmShape g = gemm_shape(kind, p.head_rows);
  const CUtensorMap* map_x = &p.map_x[kind == G_O ? 1 : (kind == G_DOWN ? 2 : 0)];
  const uint32_t b_par = c.b_uses & 1u;
  if (c.tid == 0) {
    for (int kb = 0; kb < NSLOT_B; ++kb) {
      if (kb < nkb) {
        mbar_expect_tx(&c.fullB[kb], (uint32_t)(p.N * BK * 2));
        tma_load_2d(c.sB + kb * B_TILE_BYTES, map_x, (kb0 + kb) * BK, 0, &c.fullB[kb]);
      } else {   // unused activation slots complete an empty phase, so that ONE parity (items so far) is valid for all of them
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&c.fullB[kb])) : "memory");
      }
    }
  } else if (c.tid == 32) {
    const uint32_t idesc = make_instr_desc(p.N);
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t k_glob = c.a_consumed + (uint32_t)kb;
      const uint32_t slot = k_glob % NSLOT_A;
      b_mbar_wait(c, &c.fullA[slot], (k_glob / NSLOT_A) & 1u);
      b_mbar_wait(c, &c.fullB[kb], b_par);

      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t adesc = make_smem_desc(smem_u32(c.sA + slot * A_TILE_BYTES));
      const uint64_t bdesc = make_smem_desc(smem_u32(c.sB + kb * B_TILE_BYTES));
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k)
        umma_f16(c.tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
    }
    umma_commit(c.done);
  }
  __syncwarp();
  // epilogue (warps 0-3): TMEM lane = row of the tile, column = lane of the batch
  if (c.warp < 4) {
    b_mbar_wait(c, c.done, c.done_uses & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tile * BM + c.warp * 32 + c.lane;
    float* out = p.partial + (size_t)split * p.N * g.M;
    for (int n0 = 0; n0 < p.N; n0 += 16) {
      uint32_t v[16];
      tmem_ld16(c.tmem + ((uint32_t)(c.warp * 32) << 16) + (uint32_t)n0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) out[(size_t)(n0 + j) * g.M + row] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  // every consumer thread tracks the ring (uniform values)
  c.a_consumed += (uint32_t)nkb;
  c.b_uses += 1u;
  c.done_uses += 1u;
  cta_sync();                 // the MMAs have completed (epilogue warps saw `done`): the item's ring slots are free
  if (c.tid == 0) *c.s_consumed = c.a_consumed;   // hand them back to the producer warp
}

// sum over the 256 threads; every thread gets the total
__device__ __forceinline__ float cta_sum(BCtx& c, float v) {
  v = w_sum(v);
  if (c.lane == 0) c.s_red[c.warp] = v;
  cta_sync();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += c.s_red[w];
  cta_sync();
  return t;
}
__device__ __forceinline__ uint2 rmsnorm4_b(BCtx& c, const float (&x)[4], const __nv_bfloat16* w) {
  float r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) r[e] = b_round(x[e]);
  const float tot = cta_sum(c, fmaf(r[0], r[0], r[1] * r[1]) + fmaf(r[2], r[2], r[3] * r[3]));
  const float inv = rsqrtf(tot * (1.0f / H_) + EPS_);
  const uint2 wv = *reinterpret_cast<const uint2*>(w + c.tid * 4);
  const __nv_bfloat162 n01 = __floats2bfloat162_rn((r[0] * inv) * b_lo(wv.x), (r[1] * inv) * b_hi(wv.x));
  const __nv_bfloat162 n23 = __floats2bfloat162_rn((r[2] * inv) * b_lo(wv.y), (r[3] * inv) * b_hi(wv.y));
  return make_uint2(*reinterpret_cast<const uint32_t*>(&n01), *reinterpret_cast<const uint32_t*>(&n23));
}

// ---- phase: step input -> fp32 residual + layer-0 input norm (item = lane) ---------------------------------------------------
__device__ void input_phase(BCtx& c) {
  const BStepParams& p = c.p;
  for (int b = c.cta; b < p.B; b += c.G) {
    int tok = p.token_ids ? p.token_ids[b] : -1;
    if (tok >= p.vocab) tok = p.vocab - 1;
    if (tok < 0 && p.embeds == nullptr) tok = 0;
    const __nv_bfloat16* src = tok >= 0 ? p.embed + (size_t)tok * H_ : p.embeds + (size_t)b * H_;
    const uint2 v = *reinterpret_cast<const uint2*>(src + c.tid * 4);
    const float x[4] = {b_lo(v.x), b_hi(v.x), b_lo(v.y), b_hi(v.y)};
    *reinterpret_cast<float4*>(p.res + (size_t)b * H_ + c.tid * 4) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<uint2*>(p.xn + (size_t)b * H_ + c.tid * 4) = rmsnorm4_b(c, x, p.ln_in[0]);
  }
}

// ---- phase: O / down epilogue: split-K sum -> bf16 -> residual -> next RMSNorm (item = lane) ------------------------------------
template <int SPLITS>
__device__ void resid_norm_phase(BCtx& c, const __nv_bfloat16* w_norm, float* hidden_out) {
  const BStepParams& p = c.p;
  for (int b = c.cta; b < p.B; b += c.G) {
    // all loads of the item are independent: issue them together (one L2 round trip instead of SPLITS + 1)
    float4 q[SPLITS];
#pragma unroll
    for (int s = 0; s < SPLITS; ++s) q[s] = *reinterpret_cast<const float4*>(p.partial + ((size_t)s * p.N + b) * H_ + c.tid * 4);
    const float4 r = *reinterpret_cast<const float4*>(p.res + (size_t)b * H_ + c.tid * 4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < SPLITS; ++s) { acc.x += q[s].x; acc.y += q[s].y; acc.z += q[s].z; acc.w += q[s].w; }
    float x[4] = {r.x + b_round(acc.x), r.y + b_round(acc.y), r.z + b_round(acc.z), r.w + b_round(acc.w)};
    if (!p.residual_fp32) {
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = b_round(x[e]);
    }
    *reinterpret_cast<float4*>(p.res + (size_t)b * H_ + c.tid * 4) = make_float4(x[0], x[1], x[2], x[3]);
    const uint2 n = rmsnorm4_b(c, x, w_norm);
    *reinterpret_cast<uint2*>(p.xn + (size_t)b * H_ + c.tid * 4) = n;
    if (hidden_out)
      *reinterpret_cast<float4*>(hidden_out + (size_t)b * H_ + c.tid * 4) = make_float4(b_lo(n.x), b_hi(n.x), b_lo(n.y), b_hi(n.y));
  }
}

// ---- phase: QKV epilogue (+ decode attention), item = (lane, kv head) -------------------------------------------------------
__device__ __forceinline__ int lane_position(const BStepParams& p, int b) {
  int pos = p.prefill ? p.positions[0] + b : p.positions[b];
  return pos < 0 ? 0 : (pos >= p.max_seq ? p.max_seq - 1 : pos);
}
__device__ __forceinline__ size_t kv_base(const BStepParams& p, int b, int layer, int g) {
  return (((size_t)(p.prefill ? 0 : b) * p.L + layer) * NKVH_ + g) * p.max_seq * HD_;
}
// warps 0-3: split-K sum -> bf16 -> per-head RMSNorm + rotate-half RoPE; q -> s_q (and qbuf in prefill mode), k / v -> cache row
__device__ __forceinline__ void qkv_finish(BCtx& c, int b, int g, int layer, int pos, float (*s_q)[HD_]) {
  const BStepParams& p = c.p;
  const int warp = c.warp, lane = c.lane;
  if (warp < 4) {
    const int row0 = warp < 2 ? (2 * g + warp) * HD_ : (warp == 2 ? QSZ_ + g * HD_ : QSZ_ + KVSZ_ + g * HD_);
    float4 q4[S_QKV];
#pragma unroll
    for (int s = 0; s < S_QKV; ++s) q4[s] = *reinterpret_cast<const float4*>(p.partial + ((size_t)s * p.N + b) * QKV_ROWS_ + row0 + lane * 4);
    // norm weights and the RoPE row do not depend on the partials: requested in the same round trip
    const int dbase = (lane * 4) & 63;
    const uint2 wn_raw = *reinterpret_cast<const uint2*>((warp < 2 ? p.qn[layer] : p.kn[layer]) + lane * 4);
    const uint2 cs_raw = *reinterpret_cast<const uint2*>(p.cos_t + (size_t)pos * HD_ + dbase);
    const uint2 sn_raw = *reinterpret_cast<const uint2*>(p.sin_t + (size_t)pos * HD_ + dbase);
    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < S_QKV; ++s) { t[0] += q4[s].x; t[1] += q4[s].y; t[2] += q4[s].z; t[3] += q4[s].w; }
#pragma unroll
    for (int e = 0; e < 4; ++e) t[e] = b_round(t[e]);
    const size_t base = kv_base(p, b, layer, g);
    if (warp == 3) {
      *reinterpret_cast<uint2*>(p.v_cache + base + (size_t)pos * HD_ + lane * 4) =
          make_uint2(b_bits(t[0]) | (b_bits(t[1]) << 16), b_bits(t[2]) | (b_bits(t[3]) << 16));
    } else {
      const float wn[4] = {b_lo(wn_raw.x), b_hi(wn_raw.x), b_lo(wn_raw.y), b_hi(wn_raw.y)};
      const float cs4[4] = {b_lo(cs_raw.x), b_hi(cs_raw.x), b_lo(cs_raw.y), b_hi(cs_raw.y)};
      const float sn4[4] = {b_lo(sn_raw.x), b_hi(sn_raw.x), b_lo(sn_raw.y), b_hi(sn_raw.y)};
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) ss = fmaf(t[e], t[e], ss);
      ss = w_sum(ss);
      const float rms = sqrtf(ss * (1.0f / HD_) + EPS_);
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float n = b_round((t[e] / rms) * wn[e]);
        const float other = __shfl_xor_sync(0xffffffffu, n, 16);
        const float cs = cs4[e], sn = sn4[e];
        const float x = b_round(n * cs), y = b_round(other * sn);
        o[e] = b_round(lane < 16 ? x - y : x + y);
      }
      if (warp < 2) {
        *reinterpret_cast<float4*>(&s_q[warp][lane * 4]) = make_float4(o[0], o[1], o[2], o[3]);
        if (p.prefill) *reinterpret_cast<float4*>(p.qbuf + (size_t)b * QSZ_ + (2 * g + warp) * HD_ + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        *reinterpret_cast<uint2*>(p.k_cache + base + (size_t)pos * HD_ + lane * 4) =
            make_uint2(b_bits(o[0]) | (b_bits(o[1]) << 16), b_bits(o[2]) | (b_bits(o[3]) << 16));
      }
    }
  }
}
// all 8 warps: positions 0 .. pos of (lane b, kv head g), 4 positions per warp and iteration (independent loads in flight),
// fp32 online softmax, fixed-order cross-warp merge -> abuf
__device__ __forceinline__ void attention_item(BCtx& c, int b, int g, int layer, int pos, float (*s_q)[HD_], float (*s_acc)[2][HD_],
                                               float (*s_m)[2], float (*s_l)[2]) {
  const BStepParams& p = c.p;
  const int warp = c.warp, lane = c.lane;
  const size_t base = kv_base(p, b, layer, g);
  const int n = pos + 1;
  const float4 qa = *reinterpret_cast<const float4*>(&s_q[0][lane * 4]), qb = *reinterpret_cast<const float4*>(&s_q[1][lane * 4]);
  const float q0[4] = {qa.x, qa.y, qa.z, qa.w}, q1[4] = {qb.x, qb.y, qb.z, qb.w};
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f, acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
  for (int pb = warp * 4; pb < n; pb += 32) {
    uint2 kk[4], vv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pp = pb + i < n ? pb + i : n - 1;
      kk[i] = *reinterpret_cast<const uint2*>(p.k_cache + base + (size_t)pp * HD_ + lane * 4);
      vv[i] = *reinterpret_cast<const uint2*>(p.v_cache + base + (size_t)pp * HD_ + lane * 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (pb + i >= n) break;
      const float kf[4] = {b_lo(kk[i].x), b_hi(kk[i].x), b_lo(kk[i].y), b_hi(kk[i].y)};
      const float vf[4] = {b_lo(vv[i].x), b_hi(vv[i].x), b_lo(vv[i].y), b_hi(vv[i].y)};
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) { d0 = fmaf(q0[e], kf[e], d0); d1 = fmaf(q1[e], kf[e], d1); }
      d0 = w_sum(d0) * p.attn_scale;
      d1 = w_sum(d1) * p.attn_scale;
      const float nm0 = fmaxf(m0, d0), nm1 = fmaxf(m1, d1);
      const float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - nm0), c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - nm1);
      const float e0 = __expf(d0 - nm0), e1 = __expf(d1 - nm1);
      l0 = l0 * c0 + e0; l1 = l1 * c1 + e1;
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc0[e] = fmaf(e0, vf[e], acc0[e] * c0); acc1[e] = fmaf(e1, vf[e], acc1[e] * c1); }
      m0 = nm0; m1 = nm1;
    }
  }
  if (lane == 0) { s_m[warp][0] = m0; s_m[warp][1] = m1; s_l[warp][0] = l0; s_l[warp][1] = l1; }
#pragma unroll
  for (int e = 0; e < 4; ++e) { s_acc[warp][0][lane * 4 + e] = acc0[e]; s_acc[warp][1][lane * 4 + e] = acc1[e]; }
  cta_sync();
  const int h = c.tid >> 7, d = c.tid & 127;
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < 8; ++w) M = fmaxf(M, s_m[w][h]);
  float A = 0.f, Ls = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const float f = (s_m[w][h] == -INFINITY) ? 0.f : __expf(s_m[w][h] - M);
    A = fmaf(s_acc[w][h][d], f, A);
    Ls = fmaf(s_l[w][h], f, Ls);
  }
  p.abuf[(size_t)b * QSZ_ + (2 * g + h) * HD_ + d] = __float2bfloat16_rn(A / Ls);
  cta_sync();   // the scratch is reused by the CTA's next item
}
// mode 0: finish + attention in one item (decode); 1: finish only; 2: attention only (prefill: the rows of the lanes in front
// of this one are written by other CTAs in phase 1)
__device__ void qkv_attention_phase(BCtx& c, int layer, int mode, float* scratch) {
  const BStepParams& p = c.p;
  float (*s_q)[HD_] = reinterpret_cast<float (*)[HD_]>(scratch);                       // [2][128]
  float (*s_acc)[2][HD_] = reinterpret_cast<float (*)[2][HD_]>(scratch + 2 * HD_);      // [8][2][128]
  float (*s_m)[2] = reinterpret_cast<float (*)[2]>(scratch + 2 * HD_ + 8 * 2 * HD_);    // [8][2]
  float (*s_l)[2] = s_m + 8;
  for (int item = c.cta; item < p.B * NKVH_; item += c.G) {
    const int b = item / NKVH_, g = item % NKVH_;
    const int pos = lane_position(p, b);
    if (mode != 2) qkv_finish(c, b, g, layer, pos, s_q);
    else if (c.tid < 2 * HD_) s_q[c.tid >> 7][c.tid & 127] = p.qbuf[(size_t)b * QSZ_ + 2 * g * HD_ + c.tid];
    cta_sync();   // q in shared memory; this CTA's own K / V row is visible to the whole CTA
    if (mode != 1) attention_item(c, b, g, layer, pos, s_q, s_acc, s_m, s_l);
  }
}

// ---- phase: gate/up epilogue m = r( r(silu(r(g))) * r(u) ); item = (lane, 1024-column chunk) ------------------------------------
__device__ void swiglu_phase(BCtx& c) {
  const BStepParams& p = c.p;
  for (int item = c.cta; item < p.B * 3; item += c.G) {
    const int b = item / 3, j = (item % 3) * 1024 + c.tid * 4;
    float4 pg[S_GU], pu[S_GU];
#pragma unroll
    for (int s = 0; s < S_GU; ++s) {
      const float* q = p.partial + ((size_t)s * p.N + b) * GU_ROWS_;
      pg[s] = *reinterpret_cast<const float4*>(q + j);
      pu[s] = *reinterpret_cast<const float4*>(q + INTER_ + j);
    }
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f), u = g;
#pragma unroll
    for (int s = 0; s < S_GU; ++s) {
      g.x += pg[s].x; g.y += pg[s].y; g.z += pg[s].z; g.w += pg[s].w;
      u.x += pu[s].x; u.y += pu[s].y; u.z += pu[s].z; u.w += pu[s].w;
    }
    const float gv[4] = {g.x, g.y, g.z, g.w}, uv[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gg = b_round(gv[e]);
      const float sg = b_round(__fdividef(gg, 1.0f + __expf(-gg)));
      o[e] = b_bits(sg * b_round(uv[e]));
    }
    *reinterpret_cast<uint2*>(p.mbuf + (size_t)b * INTER_ + j) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
  }
}

// ---- phase: LM head epilogue: bf16 logits, argmax (lowest index on ties); advances the lane's position (decode) -----------------
__device__ void head_phase(BCtx& c, float* scratch) {
  const BStepParams& p = c.p;
  float* s_v = scratch;
  int* s_i = reinterpret_cast<int*>(scratch + 8);
  for (int b = c.cta; b < p.B; b += c.G) {
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int r = c.tid; r < p.head_rows; r += NT) {
      float v6[S_HEAD];
#pragma unroll
      for (int s = 0; s < S_HEAD; ++s) v6[s] = p.partial[((size_t)s * p.N + b) * p.head_rows + r];
      float acc = 0.f;
#pragma unroll
      for (int s = 0; s < S_HEAD; ++s) acc += v6[s];
      const float v = b_round(acc);
      if (v > best) { best = v; best_i = r; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (c.lane == 0) { s_v[c.warp] = best; s_i[c.warp] = best_i; }
    cta_sync();
    if (c.tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (s_v[w] > best || (s_v[w] == best && s_i[w] < best_i)) { best = s_v[w]; best_i = s_i[w]; }
      if (best_i == 0x7fffffff) best_i = 0;
      p.tokens_out[b] = *c.s_abort ? -1001 : best_i;
      if (!p.prefill) p.positions[b] += 1;
    }
    cta_sync();
  }
}

// grid = one CTA per SM (cooperative launch), block = 288 (8 consumer warps + the weight producer warp)
__global__ void __launch_bounds__(NT_ALL, 1) qmk_bstep_kernel(const __grid_constant__ BStepParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ float s_scratch[2 * HD_ + 8 * 2 * HD_ + 32];
  __shared__ float s_red[16];
  __shared__ int s_abort;
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_consumed;
  BCtx c(p);
  c.sA = smem;
  c.sB = smem + NSLOT_A * A_TILE_BYTES;
  c.fullA = reinterpret_cast<uint64_t*>(smem + NSLOT_A * A_TILE_BYTES + NSLOT_B * B_TILE_BYTES);
  c.fullB = c.fullA + NSLOT_A;
  c.done = c.fullB + NSLOT_B;
  c.s_red = s_red;
  c.s_abort = &s_abort;
  c.s_consumed = &s_consumed;
  c.cta = blockIdx.x; c.G = gridDim.x; c.tid = threadIdx.x; c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
  c.a_issued = c.a_consumed = 0; c.pl = 0; c.pph = G_QKV; c.pkb = 0; c.b_uses = 0; c.done_uses = 0;
  c.bar_target = p.bar_base;
  c.trace_n = 0;
  c.t0 = clock64();
  if (c.tid == 0) {
    for (int i = 0; i < NSLOT_A; ++i) mbar_init(&c.fullA[i], 1);
    for (int i = 0; i < NSLOT_B; ++i) mbar_init(&c.fullB[i], 1);
    mbar_init(c.done, 1);
    s_abort = 0;
    s_consumed = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (c.warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  c.tmem = s_tmem;

  if (c.warp == NT / 32) {
    // ===== producer warp: the weight stream of the whole step, independent of every data dependency =====
    if (c.lane == 0) weights_producer(c, &s_consumed);
  } else {
    // ===== consumer warps =====
    input_phase(c);
    grid_barrier(c);
    for (int l = 0; l < p.L; ++l) {
      gemm_phase(c, l, G_QKV);
      grid_barrier(c);
      if (p.prefill) {
        qkv_attention_phase(c, l, 1, s_scratch);
        grid_barrier(c);
        qkv_attention_phase(c, l, 2, s_scratch);
      } else {
        qkv_attention_phase(c, l, 0, s_scratch);
      }
      grid_barrier(c);
      gemm_phase(c, l, G_O);
      grid_barrier(c);
      resid_norm_phase<S_O>(c, p.ln_post[l], nullptr);
      grid_barrier(c);
      gemm_phase(c, l, G_GU);
      grid_barrier(c);
      swiglu_phase(c);
      grid_barrier(c);
      gemm_phase(c, l, G_DOWN);
      grid_barrier(c);
      const bool last = (l == p.L - 1);
      resid_norm_phase<S_DOWN>(c, last ? p.final_norm : p.ln_in[l + 1], last ? p.hidden_out : nullptr);
      grid_barrier(c);
    }
    gemm_phase(c, p.L, G_HEAD);
    grid_barrier(c);
    head_phase(c, s_scratch);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (c.warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"(TMEM_COLS) : "memory");
}

}  // namespace qmkb
