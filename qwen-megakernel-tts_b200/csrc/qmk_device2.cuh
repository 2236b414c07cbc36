// qmk_device2.cuh — second-generation device side of the B200 decode engine (sm_100a): "group" kernel.
//
// Same contract, numerics and host ABI as qmk_device.cuh; what changes is how a layer is cut across SMs.
// The first kernel splits every projection by output rows over all CTAs, which makes five grid-wide exchanges per
// layer (x -> QKV -> attention on 8 CTAs -> O -> gate/up -> down) and leaves 140 SMs waiting while 8 of them do the
// attention.  Here the grid is 8 groups x 16 CTAs, one group per KV head:
//   * QKV rows of kv head g (2 q heads + k + v = 512 rows) live on group g; q/k/v are exchanged inside the group only.
//   * every CTA of the group evaluates the attention of its head pair itself (short context) or one position chunk
//     of it (long context, partials merged through a group-local exchange) -> `a` never crosses groups.
//   * O and down are K-SPLIT: group g multiplies its own 256 (O) / 384 (down) input columns, rows split over the
//     group's 16 CTAs, and the eight group partials of a row meet in L2: one red.add.u64 per row and group on a
//     word {arrival count : 8, fixed-point sum : 56}.  Integer adds commute, so the sum is order-independent
//     (deterministic), it is exact to 2^-24 absolute, and the arrival count rides in the same word: a consumer that
//     reads count == previous + 8 has the complete sum -- reduction and all-gather are ONE exchange, there is no
//     zeroing and no epoch (totals are cumulative; every consumer subtracts the total it read last time).
//   * gate/up rows are the ones whose SwiGLU output feeds the group's own down columns, so `m` is group-local too.
// Per layer: 2 grid-wide exchanges (h, h') + 2 group-local ones instead of 5 grid-wide ones, and no idle SMs.
// Measured (scripts/ubench/xchg7.cu, 128 CTAs): red-gather 2310 cycles, group exchange 1650, layer-shaped sequence
// of the four 6400 cycles.  Thread-block clusters with DSMEM pushes (215 cycles) were the first choice for the
// group exchanges, but this B200 hosts only seven co-resident 16-CTA clusters (one GPC is smaller), so the groups
// talk through L2 like everything else.
//
// Weight stream: per (CTA, layer) eight ring slots [QKV tile 0,1 | O | gate/up tile 0,1,2 | down half 0,1] of 32 KB
// (24 KB for the down halves); a tile is 16 rows x 1024 k, chunk-swizzled for ldmatrix and K-permuted so that a lane's
// B fragments are a contiguous run of the activation vector (same scheme as the first kernel, generalised to
// 2 / 3 / 8 tensor-core steps per warp).
#pragma once

#include "qmk_device.cuh"
#include "qmk_sample.cuh"

namespace qmk2 {
using namespace qmk;

constexpr int G2 = 128, NGRP = 8, GSZ = 16;
constexpr int SLOT2 = 32768, NSL = NSLOTS;       // 6 slots x 32 KB
static_assert(NSL == 6, "wait_full() of the first kernel is reused: same slot count");
constexpr int QKV_LOC = 32, O_LOC = 64, GU_LOC = 48, M_LOC = 24, KB_O = 256, KB_D = 384;
constexpr int DOWN_SLOT = 32 * KB_D * 2;         // 24576: 32 rows x 384 k
constexpr int LAYER_BYTES2 = 6 * SLOT2 + 2 * DOWN_SLOT;   // 245760
constexpr int ENT_PER_LAYER = 8;
constexpr int HEAD_TILES_MAX = 2;                // 3072 / 128 = 24 rows -> 2 tiles
constexpr int S2_MAX = GSZ;                      // attention position chunks per kv head
constexpr int FIX_SHIFT = 24;
constexpr u64 CNT_ONE = 1ull << 56;

// exchange buffer (bytes)
constexpr size_t XB_ACC = 0;                               // u64[2][1024]  accumulators A (h: down), B (h': O)
constexpr size_t XB_SNAP = XB_ACC + 2 * 1024 * 8;          // u64[2][1024]  totals at the end of the previous launch
constexpr size_t XB_LL = XB_SNAP + 2 * 1024 * 8;           // start of the epoch-tagged region (cleared on epoch wrap)
// Group buffers (q0 | q1 | k | v: 512 LL4 words; m: 384) are placed by CALIBRATION: the latency of a group exchange depends on
// where its 2 KB buffer lives in L2 (bimodal on B200: ~1950 vs ~2900 cycles per exchange, stable for an address), and the
// slowest group gates the next grid-wide exchange.  The engine measures a pool of candidate slots once and gives every group
// its two fastest (qmk2_calib_kernel; table behind the roles).
constexpr int XP_CAND = 48, XP_WORDS = 512;
constexpr size_t XB_POOL = XB_LL;                          // u32[48][512]
constexpr size_t XB_LOGITS = XB_POOL + (size_t)XP_CAND * XP_WORDS * 4;   // u32[3072]
constexpr size_t XB_PART = XB_LOGITS + MAX_HEAD_ROWS * 4;  // u64[8][2][16][PART_STRIDE]
constexpr size_t XB_ROLE = XB_PART + (size_t)NGRP * 2 * S2_MAX * PART_STRIDE * 8;   // int[128]: blockIdx -> role (group * 16 + rank)
constexpr size_t XB_SLOTS = XB_ROLE + G2 * 4;                // int[8][2]: pool slot of the group's q/k/v buffer and of its m buffer
constexpr size_t XBUF2_BYTES = XB_SLOTS + NGRP * 2 * 4;

// shared memory
constexpr int S2_RING = 0;
constexpr int S2_VEC = S2_RING + NSL * SLOT2;      // bf16[1024] normalised input of QKV / gate-up / head
constexpr int S2_A = S2_VEC + 2048;                // bf16[384]  attention output (256) or m (384): input of O / down
constexpr int S2_ACC = S2_A + 768;                 // float[8][2][128] attention cross-warp merge / sampling logits
constexpr int S2_SMALL = S2_ACC + NCW * 2 * HD * 4;  // float[1024] attention scratch
constexpr int S2_PART = S2_SMALL + 4096;           // float[64][8]
constexpr int S2_RED = S2_PART + 64 * NCW * 4;     // float[64]
constexpr int S2_BAR = S2_RED + 256;               // u64[8]
constexpr int S2_PREV = S2_BAR + 64;               // u64[2][1024]: totals of the accumulator words read last time (thread-private slots)
constexpr int S2_MISC = S2_PREV + 2 * 1024 * 8;    // abort, delays
constexpr int SMEM2_BYTES = S2_MISC + 256;
static_assert(SMEM2_BYTES <= 232448, "shared memory budget");

enum Kind2 { K2_QKV = 0, K2_ATTN = 1, K2_O = 2, K2_GU = 3, K2_DOWN = 4, K2_HEAD = 5, K2_ARGMAX = 6 };

__device__ __forceinline__ void red_add64(u64* p, u64 v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ld_acc2(const u64* p, u64& a, u64& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ bool acc_done(u64 now, u64 prev) { return (uint32_t)(((now - prev) + (1ull << 55)) >> 56) == (uint32_t)NGRP; }
__device__ __forceinline__ float acc_value(u64 now, u64 prev) {
  // signed 56-bit fixed point -> float through two 32-bit conversions (a 64-bit I2F is a multi-instruction slow path)
  const u64 d = (now - prev) - (u64)NGRP * CNT_ONE;
  const float hi = (float)(int)(uint32_t)(d >> 32), lo = (float)(uint32_t)d;
  return fmaf(hi, 4294967296.0f, lo) * (1.0f / (float)(1 << FIX_SHIFT));
}
__device__ __forceinline__ u64 acc_word(float v) { return CNT_ONE + (u64)__float2ll_rn(v * (float)(1 << FIX_SHIFT)); }

struct Ctx2 : Ctx {
  uint8_t* s_a;      // bf16[384]
  u64* acc;          // [2][1024]
  uint8_t* xb;       // exchange buffer base
  uint8_t* smem0;    // start of the CTA's shared memory
  int g, j;          // group (kv head), rank in group
  int slot_q, slot_m;   // pool slots of the group's exchange buffers
  __device__ Ctx2(const Params& pp) : Ctx(pp) {}
};

// Out of line, by-value arguments only (the totals stay in registers): returns once all four words are complete; the
// caller then re-reads them.
__device__ __noinline__ void acc_wait_slow(int* status, volatile int* s_abort, long long t0, long long timeout, int cta, int cur_idx,
                                           const u64* p, u64 p0, u64 p1, u64 p2, u64 p3) {
  uint32_t spins = 0;
  for (;;) {
    u64 n0, n1, n2, n3;
    ld_acc2(p, n0, n1);
    ld_acc2(p + 2, n2, n3);
    if (acc_done(n0, p0) & acc_done(n1, p1) & acc_done(n2, p2) & acc_done(n3, p3)) return;
    if ((++spins & 127u) == 0 && check_abort_slow(status, s_abort, t0, timeout, cta, cur_idx, ST_TIMEOUT_LL, -3)) return;
  }
}

// ---- weight producer: one lane refills the slots a phase has released ------------------------------------------
struct Prod2 {
  const uint8_t* layer_base;
  int step, l, e;
  uint32_t k;
  int left;
};
__device__ __forceinline__ const uint8_t* model_layers(const Params& p, int cta, int step) {
  const ModelDesc& md = p.models[p.steps[step].model];
  return md.packed_layers + (size_t)cta * md.L * LAYER_BYTES2;
}
__device__ __forceinline__ int head_tiles(const HeadDesc& h) { return h.rows > 0 ? (h.rows / G2 + 15) / 16 : 0; }
__device__ __forceinline__ void prod2_init(Ctx2& c, Prod2& pr) {
  const Params& p = c.p;
  pr.step = 0; pr.l = 0; pr.e = 0; pr.k = 0; pr.left = 0;
  for (int st = 0; st < p.n_steps; ++st) pr.left += p.models[p.steps[st].model].L * ENT_PER_LAYER + head_tiles(p.steps[st].head);
  if (p.frames.n_frames > 1) pr.left *= p.frames.n_frames;   // the frame loop repeats the step program
  pr.layer_base = model_layers(p, c.cta, 0);
}
// Weights are read exactly once per step and a step's weights exceed L2: stream them with an evict-first policy so
// that they do not push out what IS re-used (KV rows, embedding tables, norm weights, the exchange words).
__device__ __forceinline__ void tma_bulk_g2s_stream(void* dst_smem, const void* src_gmem, uint32_t bytes, u64* bar) {
  u64 policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void prod2_tma(Ctx2& c, uint32_t k, const uint8_t* src, uint32_t bytes) {
  const int slot = k % NSL;
  if (c.p.delay_o_idle == -1) bytes = 16;   // experiment (QMK_POLL_DELAY_O=-1): no weight traffic, exchanges on an idle memory system
  mbar_arrive_expect_tx(&c.full[slot], bytes);
  if (c.p.delay_o_idle == -2) tma_bulk_g2s(c.ring + (size_t)slot * SLOT2, src, bytes, &c.full[slot]);   // QMK_POLL_DELAY_O=-2: default L2 policy
  else tma_bulk_g2s_stream(c.ring + (size_t)slot * SLOT2, src, bytes, &c.full[slot]);
}
__device__ __forceinline__ void prod2_issue(Ctx2& c, Prod2& pr, int n) {
  const Params& p = c.p;
  if (n > pr.left) n = pr.left;
  pr.left -= n;
  // Called by warps 2..7 (the ones that neither finalise nor publish), all of which keep the cursor; stage i of the call
  // is issued by warp 2 + i % 6, so that a three-stage refill costs one mbarrier + one bulk copy of latency instead of three.
  for (int issuer = 2; n > 0; --n, ++pr.k, issuer = (issuer == NCW - 1) ? 2 : issuer + 1) {
    // skip steps without a head (their head phase has no stage); after the last step the program restarts (next frame)
    while (pr.l == p.models[p.steps[pr.step].model].L && pr.e >= head_tiles(p.steps[pr.step].head)) {
      pr.l = 0; pr.e = 0;
      if (++pr.step == p.n_steps) pr.step = 0;
      pr.layer_base = model_layers(p, c.cta, pr.step);
    }
    if (pr.l < p.models[p.steps[pr.step].model].L) {
      const uint32_t off = pr.e < 6 ? (uint32_t)pr.e * SLOT2 : 6u * SLOT2 + (uint32_t)(pr.e - 6) * DOWN_SLOT;
      if (c.lane == 0 && c.warp == issuer) prod2_tma(c, pr.k, pr.layer_base + off, pr.e < 6 ? SLOT2 : DOWN_SLOT);
      if (++pr.e == ENT_PER_LAYER) { pr.e = 0; ++pr.l; pr.layer_base += LAYER_BYTES2; }
    } else {
      const HeadDesc& h = p.steps[pr.step].head;
      if (c.lane == 0 && c.warp == issuer) prod2_tma(c, pr.k, h.packed + ((size_t)c.cta * head_tiles(h) + pr.e) * SLOT2, SLOT2);
      ++pr.e;
    }
  }
}

// ---- tensor-core stages, specialised by phase shape (NT tiles, NK steps per warp, first PRE tiles preloaded) --------
template <int NT, int NK, int PRE>
__device__ __forceinline__ void preload_a(uint32_t (&apre)[16][4], const uint32_t (&tb)[4], int warp, int a_khalf, int a_sw) {
  static_assert(PRE * NK <= 16, "preload registers");
#pragma unroll
  for (int t = 0; t < PRE; ++t)
#pragma unroll
    for (int jx = 0; jx < NK; ++jx)
      ldsm4(apre[t * NK + jx], tb[t] + ((uint32_t)((((warp * NK + jx) * 2 + a_khalf) ^ a_sw)) << 4));
}
// acc[t] = tile t x activation slice of this warp.  `bvec`: bf16 vector in natural order; lane q4 owns the contiguous run
// of 4 NK elements at  warp * 16 NK + q4 * 4 NK  (the weights are K-permuted to match, see pack_chunk2).
template <int NT, int NK, int PRE>
__device__ __forceinline__ void mma_tiles(float (&acc)[4][4], const uint32_t (&apre)[16][4], const uint32_t (&tb)[4], const uint8_t* bvec,
                                          int ntiles, int warp, int q4, int a_khalf, int a_sw) {
  uint32_t bfrag[NK][2];
  const uint8_t* bp = bvec + warp * (32 * NK) + q4 * (8 * NK);
  if (NK % 2 == 0) {
#pragma unroll
    for (int i = 0; i < NK / 2; ++i) {
      const uint4 v = *reinterpret_cast<const uint4*>(bp + i * 16);
      bfrag[2 * i][0] = v.x; bfrag[2 * i][1] = v.y; bfrag[2 * i + 1][0] = v.z; bfrag[2 * i + 1][1] = v.w;
    }
  } else {
#pragma unroll
    for (int jx = 0; jx < NK; ++jx) {
      const uint2 v = *reinterpret_cast<const uint2*>(bp + jx * 8);
      bfrag[jx][0] = v.x; bfrag[jx][1] = v.y;
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    if (t < PRE) {   // tiles below PRE always exist
      if (NK >= 4) {
        float acc2[4] = {0.f, 0.f, 0.f, 0.f};   // two accumulators halve the dependent HMMA chain
#pragma unroll
        for (int jx = 0; jx < NK; jx += 2) {
          mma16816(acc[t], apre[t * NK + jx], bfrag[jx]);
          if (jx + 1 < NK) mma16816(acc2, apre[t * NK + jx + 1], bfrag[jx + 1]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[t][e] += acc2[e];
      } else {
#pragma unroll
        for (int jx = 0; jx < NK; ++jx) mma16816(acc[t], apre[t * NK + jx], bfrag[jx]);
      }
    } else if (t < ntiles) {
      float acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int jb = 0; jb < NK; jb += 4) {
        uint32_t afrag[4][4];
#pragma unroll
        for (int jx = 0; jx < 4; ++jx)
          if (jb + jx < NK) ldsm4(afrag[jx], tb[t] + ((uint32_t)((((warp * NK + jb + jx) * 2 + a_khalf) ^ a_sw)) << 4));
        if (jb + 0 < NK) mma16816(acc[t], afrag[0], bfrag[jb]);
        if (jb + 1 < NK) mma16816(acc2, afrag[1], bfrag[jb + 1]);
        if (jb + 2 < NK) mma16816(acc[t], afrag[2], bfrag[jb + 2]);
        if (jb + 3 < NK) mma16816(acc2, afrag[3], bfrag[jb + 3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][e] += acc2[e];
    }
  }
}

// ---- attention ---------------------------------------------------------------------------------------------------
// A round = ATT2_PW cached positions per warp (ATT2_ROUND per CTA) scored together: their rows are one batch of loads and their
// 2 ATT2_PW partial dot products one transposed warp reduction.  Measured, us per talker step with 5 -> 6 positions per warp: p = 45: 323 -> 325,
// 90: 331 -> 337, 200: 345 -> 353, 500: 386 -> 390, 1000: 432 -> 437, 1500: 463 -> 456, 2000: 497 -> 487 (one round fewer beyond 1920
// positions): 5 is better over the headline's positions 9..509.
#ifndef QMK_ATT_PW
#define QMK_ATT_PW 5
#endif
constexpr int ATT2_PW = QMK_ATT_PW, ATT2_ROUND = NCW * ATT2_PW;
static_assert(ATT2_PW == 5 || ATT2_PW == 6, "the transposed reduction below folds 2 x 5 or 2 x 6 values");
struct KvRegs2 {
  uint2 k[ATT2_PW];
  uint2 v[ATT2_PW];
};
// Transposed reduction of 2 * ATT2_PW per-lane partial dot products; every lane ends up with all totals (12 reduction shuffles
// + the broadcasts instead of 5 per value).  Halves fold 6 -> 3 -> 2 (padded) -> 1.
__device__ __forceinline__ void warp_sum2pw_bcast(float (&v)[2 * ATT2_PW], int lane) {
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0;
  float u[6];
#pragma unroll
  for (int k = 0; k < ATT2_PW; ++k)
    u[k] = (b4 ? v[k + ATT2_PW] : v[k]) + __shfl_xor_sync(0xffffffffu, b4 ? v[k] : v[k + ATT2_PW], 16);
  if (ATT2_PW == 5) u[5] = 0.f;
  float w[4];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    w[k] = (b3 ? u[k + 3] : u[k]) + __shfl_xor_sync(0xffffffffu, b3 ? u[k] : u[k + 3], 8);
  w[3] = 0.f;
  float x[2];
#pragma unroll
  for (int k = 0; k < 2; ++k)
    x[k] = (b2 ? w[k + 2] : w[k]) + __shfl_xor_sync(0xffffffffu, b2 ? w[k] : w[k + 2], 4);
  float y = (b1 ? x[1] : x[0]) + __shfl_xor_sync(0xffffffffu, b1 ? x[0] : x[1], 2);
  y += __shfl_xor_sync(0xffffffffu, y, 1);
  // value index held by a lane: b4 * ATT2_PW + b3 * 3 + b2 * 2 + b1  (valid combinations only)
#pragma unroll
  for (int idx = 0; idx < 2 * ATT2_PW; ++idx) {
    const int r = idx % ATT2_PW, hi = idx / ATT2_PW;
    const int src = hi * 16 + (r >= 3 ? 8 : 0) + ((r % 3) >= 2 ? 4 : 0) + (((r % 3) % 2) ? 2 : 0);
    v[idx] = __shfl_sync(0xffffffffu, y, src);
  }
}
struct AttnItem2 {
  int S, p0, p1;
  bool has;
};
#ifndef QMK_ATT_SOLO_ROUNDS
#define QMK_ATT_SOLO_ROUNDS 2
#endif
constexpr int ATT_SOLO_ROUNDS = QMK_ATT_SOLO_ROUNDS;
__device__ __forceinline__ AttnItem2 attn_item2(int position, int j) {
  AttnItem2 it;
  const int n = position + 1;
  // Up to two rounds of 40 positions every CTA of the group evaluates the whole context itself; beyond that the context
  // is split over the group, one round per CTA.  Measured per talker step: a second round +33 us, a third +25 us more,
  // the split's partial exchange and merge +57 us.
  int S0 = n <= ATT2_ROUND * ATT_SOLO_ROUNDS ? 1 : (n + ATT2_ROUND - 1) / ATT2_ROUND;
  if (S0 > S2_MAX) S0 = S2_MAX;
  const int C = (n + S0 - 1) / S0;
  it.S = (n + C - 1) / C;
  if (it.S == 1) { it.p0 = 0; it.p1 = n; it.has = true; return it; }   // short context: every CTA of the group, redundantly
  it.has = j < it.S;
  it.p0 = j * C;
  it.p1 = it.p0 + C < n ? it.p0 + C : n;
  return it;
}
// Cached K/V rows are read with strong loads (see the append at the end of phase_attn2); like .cg they bypass L1.
__device__ __forceinline__ uint2 ld_strong_u2(const void* p) {
  uint2 v;
#ifdef QMK_EXP_KV_LD_CG
  asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
#elif defined(QMK_EXP_KV_LD_NOCLOBBER)
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
#else
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
#endif
  return v;
}
__device__ __forceinline__ void attn_l2_prefetch(const Ctx2& c, const ModelDesc& p, int l, int position, const AttnItem2& it) {
  const size_t base = ((size_t)(l * NKVH + c.g) * p.max_seq) * HD;
  for (int i = c.tid; i < (it.p1 - it.p0) * 4; i += NCT) {   // (row, K|V, half row): 128-byte lines
    const int pos = it.p0 + (i >> 2);
    if (pos != position) {
      const __nv_bfloat16* ptr = ((i & 2) ? p.v_cache : p.k_cache) + base + (size_t)pos * HD + (i & 1) * 64;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
    }
  }
}
__device__ __forceinline__ void attn_prefetch2(const Ctx2& c, const ModelDesc& p, int l, int position, const AttnItem2& it, int round, KvRegs2& r) {
  const size_t base = ((size_t)(l * NKVH + c.g) * p.max_seq) * HD;
#pragma unroll
  for (int i = 0; i < ATT2_PW; ++i) {
    const int pos = it.p0 + round * ATT2_ROUND + c.warp + NCW * i;
    if (pos < it.p1 && pos != position) {
      const size_t off = base + (size_t)pos * HD + c.lane * 4;
      r.k[i] = ld_strong_u2(p.k_cache + off);
      r.v[i] = ld_strong_u2(p.v_cache + off);
    }
  }
}

// q (2 heads), k, v of the group arrive as LL4 words; per-head RMSNorm + rotate-half RoPE; scores / online softmax / PV
// over this CTA's positions; cross-warp merge; (long context) cross-chunk merge through group-local LL8 words.
// Result: bf16 a[256] of the group's two q heads in c.s_a (every CTA of the group holds the same values).
template <bool TR>
__device__ void phase_attn2(Ctx2& c, const ModelDesc& md, int l, int position, uint32_t epoch, const AttnItem2& it, KvRegs2& kv, Prod2& prod,
                            int a_row, int a_khalf, int a_sw) {
  const Params& p = c.p;
  const uint32_t* x_q = reinterpret_cast<const uint32_t*>(c.xb + XB_POOL) + (size_t)c.slot_q * XP_WORDS;
  u64* x_part = reinterpret_cast<u64*>(c.xb + XB_PART) + (size_t)c.g * 2 * S2_MAX * PART_STRIDE;
  float* s_small = c.s_small;
  __nv_bfloat16* s_a = reinterpret_cast<__nv_bfloat16*>(c.s_a);

  bool retried = false;
  if (it.has) attn_prefetch2(c, md, l, position, it, 0, kv);   // L2 hits (prefetched during the QKV phase)
  uint2 nw_raw = make_uint2(0, 0);
  if (c.warp < 3) nw_raw = *reinterpret_cast<const uint2*>(md.aux_layers + ((size_t)l * 2) * AUX_BYTES + 2048 + (c.warp == 2 ? 256 : 0) + c.lane * 8);
  wait_window(c, c.s_delay[DL_ATTN]);
  if (c.warp < 4) {
    const uint4 w = ll4_wait(c, x_q + c.warp * HD + c.lane * 4, epoch, retried);
    float t[4] = {ll4_val(w.x), ll4_val(w.y), ll4_val(w.z), ll4_val(w.w)};
    if (c.warp == 3) {
      *reinterpret_cast<float4*>(s_small + SS_V + c.lane * 4) = make_float4(t[0], t[1], t[2], t[3]);
    } else {
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) ss = fmaf(t[e], t[e], ss);
      ss = warp_sum(ss);
      const float rms = sqrtf(ss * (1.0f / HD) + EPS);
      const int dbase = (c.lane * 4) & 63;
      const float nw[4] = {bf16_lo(nw_raw.x), bf16_hi(nw_raw.x), bf16_lo(nw_raw.y), bf16_hi(nw_raw.y)};
      float o4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float n = bf16_round((t[e] / rms) * nw[e]);
        const float o = __shfl_xor_sync(0xffffffffu, n, 16);
        const float cs = s_small[SS_CS + dbase + e], sn = s_small[SS_CS + 64 + dbase + e];
        const float a = bf16_round(n * cs), b = bf16_round(o * sn);
        o4[e] = bf16_round(c.lane < 16 ? a - b : a + b);
      }
      float* dst = (c.warp < 2) ? s_small + SS_QN + c.warp * HD : s_small + SS_KN;
      *reinterpret_cast<float4*>(dst + c.lane * 4) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
  }
  gather_bar(c, DL_ATTN, retried);
  if (c.tid == 0) c.s_delay[2 * DL_N + DL_ATTN] += (int)((clock64() - c.t_pub) >> 4);   // in-situ cost of this group's q/k/v buffer (autotuner)
  trace_sub<TR>(c, 1);

  float q0[4], q1[4], acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  {
    const float4 a = *reinterpret_cast<const float4*>(s_small + SS_QN + c.lane * 4);
    const float4 b = *reinterpret_cast<const float4*>(s_small + SS_QN + HD + c.lane * 4);
    q0[0] = a.x; q0[1] = a.y; q0[2] = a.z; q0[3] = a.w;
    q1[0] = b.x; q1[1] = b.y; q1[2] = b.z; q1[3] = b.w;
  }
  const int len = it.has ? it.p1 - it.p0 : 0;
  const int nrounds = (len + ATT2_ROUND - 1) / ATT2_ROUND;
  if (position < 2 * NCW) {
    // Tiny context (every code-predictor step, the first talker steps): this warp owns positions w and w + 8 at most.
    // Straight-line version of the loop below.
    const int n = position + 1, pa = c.warp, pb = c.warp + NCW;
    if (pa < n) {   // warp-uniform
      const bool hb = pb < n;
      float ka[4], kb[4], va[4], vb[4];
      if (pa == position) {
        const float4 kk = *reinterpret_cast<const float4*>(s_small + SS_KN + c.lane * 4), vv = *reinterpret_cast<const float4*>(s_small + SS_V + c.lane * 4);
        ka[0] = kk.x; ka[1] = kk.y; ka[2] = kk.z; ka[3] = kk.w; va[0] = vv.x; va[1] = vv.y; va[2] = vv.z; va[3] = vv.w;
      } else {
        ka[0] = bf16_lo(kv.k[0].x); ka[1] = bf16_hi(kv.k[0].x); ka[2] = bf16_lo(kv.k[0].y); ka[3] = bf16_hi(kv.k[0].y);
        va[0] = bf16_lo(kv.v[0].x); va[1] = bf16_hi(kv.v[0].x); va[2] = bf16_lo(kv.v[0].y); va[3] = bf16_hi(kv.v[0].y);
      }
      if (pb == position) {
        const float4 kk = *reinterpret_cast<const float4*>(s_small + SS_KN + c.lane * 4), vv = *reinterpret_cast<const float4*>(s_small + SS_V + c.lane * 4);
        kb[0] = kk.x; kb[1] = kk.y; kb[2] = kk.z; kb[3] = kk.w; vb[0] = vv.x; vb[1] = vv.y; vb[2] = vv.z; vb[3] = vv.w;
      } else if (hb) {
        kb[0] = bf16_lo(kv.k[1].x); kb[1] = bf16_hi(kv.k[1].x); kb[2] = bf16_lo(kv.k[1].y); kb[3] = bf16_hi(kv.k[1].y);
        vb[0] = bf16_lo(kv.v[1].x); vb[1] = bf16_hi(kv.v[1].x); vb[2] = bf16_lo(kv.v[1].y); vb[3] = bf16_hi(kv.v[1].y);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) { kb[e] = 0.f; vb[e] = 0.f; }
      }
      float d4[4] = {0.f, 0.f, 0.f, 0.f};   // {q0.ka, q0.kb, q1.ka, q1.kb}: same order as the general loop (sc[0], sc[1], sc[5], sc[6])
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        d4[0] = fmaf(q0[e], ka[e], d4[0]); d4[1] = fmaf(q0[e], kb[e], d4[1]);
        d4[2] = fmaf(q1[e], ka[e], d4[2]); d4[3] = fmaf(q1[e], kb[e], d4[3]);
      }
      const float tot = warp_sum4(d4, c.lane);
      const float s0a = __shfl_sync(0xffffffffu, tot, 0) * p.attn_scale, s0b = __shfl_sync(0xffffffffu, tot, 8) * p.attn_scale;
      const float s1a = __shfl_sync(0xffffffffu, tot, 16) * p.attn_scale, s1b = __shfl_sync(0xffffffffu, tot, 24) * p.attn_scale;
      m0 = hb ? fmaxf(s0a, s0b) : s0a;
      m1 = hb ? fmaxf(s1a, s1b) : s1a;
      const float e0a = __expf(s0a - m0), e0b = hb ? __expf(s0b - m0) : 0.f;
      const float e1a = __expf(s1a - m1), e1b = hb ? __expf(s1b - m1) : 0.f;
      l0 = e0a + e0b; l1 = e1a + e1b;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc0[e] = fmaf(e0b, vb[e], e0a * va[e]);
        acc1[e] = fmaf(e1b, vb[e], e1a * va[e]);
      }
    }
  } else
  for (int r = 0; r < nrounds; ++r) {
    if (r > 0) attn_prefetch2(c, md, l, position, it, r, kv);   // (double-buffering the rows in registers spills: +15 us per step, measured twice)
    float sc[2 * ATT2_PW];
    const int pos_first = it.p0 + r * ATT2_ROUND + c.warp;
    if (pos_first < it.p1) {  // warp-uniform
      const int nv = min(ATT2_PW, (it.p1 - pos_first + NCW - 1) / NCW);
#pragma unroll
      for (int i = 0; i < 2 * ATT2_PW; ++i) sc[i] = 0.f;
#pragma unroll
      for (int i = 0; i < ATT2_PW; ++i) {
        if (i >= nv) break;
        const int pos = pos_first + NCW * i;
        float kf[4];
        if (pos == position) {
          const float4 kk = *reinterpret_cast<const float4*>(s_small + SS_KN + c.lane * 4);
          kf[0] = kk.x; kf[1] = kk.y; kf[2] = kk.z; kf[3] = kk.w;
        } else {
          kf[0] = bf16_lo(kv.k[i].x); kf[1] = bf16_hi(kv.k[i].x);
          kf[2] = bf16_lo(kv.k[i].y); kf[3] = bf16_hi(kv.k[i].y);
        }
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          d0 = fmaf(q0[e], kf[e], d0);
          d1 = fmaf(q1[e], kf[e], d1);
        }
        sc[i] = d0;
        sc[ATT2_PW + i] = d1;
      }
      if (nv <= 2) {
        const float v4[4] = {sc[0], sc[1], sc[ATT2_PW], sc[ATT2_PW + 1]};
        const float tot = warp_sum4(v4, c.lane);
        sc[0] = __shfl_sync(0xffffffffu, tot, 0);
        sc[1] = __shfl_sync(0xffffffffu, tot, 8);
        sc[ATT2_PW] = __shfl_sync(0xffffffffu, tot, 16);
        sc[ATT2_PW + 1] = __shfl_sync(0xffffffffu, tot, 24);
      } else {
        warp_sum2pw_bcast(sc, c.lane);
      }
      float mx0 = m0, mx1 = m1;
#pragma unroll
      for (int i = 0; i < ATT2_PW; ++i) {
        if (i >= nv) break;
        sc[i] *= p.attn_scale;
        sc[ATT2_PW + i] *= p.attn_scale;
        mx0 = fmaxf(mx0, sc[i]);
        mx1 = fmaxf(mx1, sc[ATT2_PW + i]);
      }
      const float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - mx0);
      const float c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - mx1);
      l0 *= c0; l1 *= c1;
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc0[e] *= c0; acc1[e] *= c1; }
#pragma unroll
      for (int i = 0; i < ATT2_PW; ++i) {
        if (i >= nv) break;
        const int pos = pos_first + NCW * i;
        float vf[4];
        if (pos == position) {
          const float4 vv = *reinterpret_cast<const float4*>(s_small + SS_V + c.lane * 4);
          vf[0] = vv.x; vf[1] = vv.y; vf[2] = vv.z; vf[3] = vv.w;
        } else {
          vf[0] = bf16_lo(kv.v[i].x); vf[1] = bf16_hi(kv.v[i].x);
          vf[2] = bf16_lo(kv.v[i].y); vf[3] = bf16_hi(kv.v[i].y);
        }
        const float e0 = __expf(sc[i] - mx0), e1 = __expf(sc[ATT2_PW + i] - mx1);
        l0 += e0; l1 += e1;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc0[e] = fmaf(e0, vf[e], acc0[e]);
          acc1[e] = fmaf(e1, vf[e], acc1[e]);
        }
      }
      m0 = mx0; m1 = mx1;
    }
  }
  trace_sub<TR>(c, 2);
  // The O projection has no exchange in front of it whose latency could hide the shared-memory reads of its weight
  // tile (32 KB per CTA = 256 cycles of LDS bandwidth): move the A fragments into registers here, under the merge's
  // barriers (the score loop's registers are dead now).
  uint32_t apre[16][4];
  uint32_t tbo[4];
  {
    const uint32_t obase = smem_u32(c.ring + (size_t)(c.k % NSL) * SLOT2) + (uint32_t)a_row * 512u;
#pragma unroll
    for (int t = 0; t < 4; ++t) tbo[t] = obase + (uint32_t)t * 8192u;
    wait_full(c, c.k);
    preload_a<4, 2, 4>(apre, tbo, c.warp, a_khalf, a_sw);
  }
  // cross-warp merge: common max first, then plain sums in a fixed order
  if (c.lane == 0) {
    s_small[SS_M + c.warp * 2 + 0] = m0;
    s_small[SS_M + c.warp * 2 + 1] = m1;
  }
  consumer_bar();
  float M0 = -INFINITY, M1 = -INFINITY;
#pragma unroll
  for (int w = 0; w < NCW; ++w) {
    M0 = fmaxf(M0, s_small[SS_M + w * 2 + 0]);
    M1 = fmaxf(M1, s_small[SS_M + w * 2 + 1]);
  }
  {
    const float f0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - M0);
    const float f1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - M1);
    float* s_acc = c.s_acc;  // [NCW][2][128]
    *reinterpret_cast<float4*>(s_acc + (c.warp * 2 + 0) * HD + c.lane * 4) =
        make_float4(acc0[0] * f0, acc0[1] * f0, acc0[2] * f0, acc0[3] * f0);
    *reinterpret_cast<float4*>(s_acc + (c.warp * 2 + 1) * HD + c.lane * 4) =
        make_float4(acc1[0] * f1, acc1[1] * f1, acc1[2] * f1, acc1[3] * f1);
    if (c.lane == 0) {
      s_small[SS_L + c.warp * 2 + 0] = l0 * f0;
      s_small[SS_L + c.warp * 2 + 1] = l1 * f1;
    }
  }
  consumer_bar();
  {
    const int h = c.tid >> 7, d = c.tid & 127;
    float A = 0.f, Lsum = 0.f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) {
      A += c.s_acc[(w * 2 + h) * HD + d];
      Lsum += s_small[SS_L + w * 2 + h];
    }
    const float Mh = h ? M1 : M0;
    if (QMK_LIKELY(it.S == 1)) {
      s_a[c.tid] = __float2bfloat16_rn(A / Lsum);
    } else {
      if (it.has) {
        u64* part = x_part + ((size_t)h * S2_MAX + c.j) * PART_STRIDE;
        ll8_st(part + 2 + d, __float_as_uint(A), epoch);
        if (d == 0) {
          ll8_st(part + 0, __float_as_uint(Mh), epoch);
          ll8_st(part + 1, __float_as_uint(Lsum), epoch);
        }
      }
      // Every CTA of the group merges the chunks of both heads in the fixed order s = 0..S-1.  One L2 round trip, no
      // barrier: a thread polls its own accumulator words of up to eight chunks together with ONE of the warp's max / sum
      // words (lane s: max of chunk s, lane 16 + s: its sum; the warp shares them through shuffles).
      const int ms = c.lane & 15;
      const bool ml_valid = ms < it.S;
      const u64* pml = x_part + ((size_t)h * S2_MAX + ms) * PART_STRIDE + (c.lane >> 4);
      float Mx = -INFINITY, mlv = (c.lane >> 4) ? 0.f : -INFINITY, Lt = 0.f, B = 0.f;
      for (int s0 = 0; s0 < it.S; s0 += 8) {
        u64 w[8], mlw = 0;
        uint32_t spins = 0;
        for (;;) {
          bool all = true;
          if (s0 == 0 && ml_valid) mlw = ll8_ld(pml);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (s0 + i < it.S) w[i] = ll8_ld(x_part + ((size_t)h * S2_MAX + s0 + i) * PART_STRIDE + 2 + d);
          if (s0 == 0 && ml_valid) all = (uint32_t)(mlw >> 32) == epoch;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (s0 + i < it.S) all &= (uint32_t)(w[i] >> 32) == epoch;
          if (QMK_LIKELY(all)) break;
          if ((++spins & 63u) == 0 && check_abort(c, ST_TIMEOUT_LL, -1)) break;
        }
        __syncwarp();
        if (s0 == 0) {
          if (ml_valid) mlv = __uint_as_float((uint32_t)mlw);
          for (int s = 0; s < it.S; ++s) Mx = fmaxf(Mx, __shfl_sync(0xffffffffu, mlv, s));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (s0 + i < it.S) {   // warp-uniform
            const float f = __expf(__shfl_sync(0xffffffffu, mlv, s0 + i) - Mx);
            Lt = fmaf(__shfl_sync(0xffffffffu, mlv, 16 + s0 + i), f, Lt);
            B = fmaf(__uint_as_float((uint32_t)w[i]), f, B);
          }
        }
      }
      s_a[c.tid] = __float2bfloat16_rn(B / Lt);
    }
  }
  trace_sub<TR>(c, 3);
  consumer_bar();   // s_a complete
  // ---- O projection, K-split by group: rows 64 j .. 64 j + 63 of W_o x this group's 256 attention outputs.  It has no
  //      exchange in front of it, so it runs here instead of as a phase of its own; the eight groups' partials of a row
  //      meet in L2 (one fixed-point add per row). ----
  {
    float acc4[4][4];
    mma_tiles<4, 2, 4>(acc4, apre, tbo, c.s_a, 4, c.warp, c.lane & 3, a_khalf, a_sw);
    c.k += 1;
    if ((c.lane & 3) == 0) {
      const int g8 = c.lane >> 2;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        c.s_part[((t * 16 + g8) * NCW) + c.warp] = acc4[t][0];
        c.s_part[((t * 16 + g8 + 8) * NCW) + c.warp] = acc4[t][2];
      }
    }
    consumer_bar();
    trace_sub<TR>(c, 4);
    if (c.warp >= 2) prod2_issue(c, prod, 1);
    if (c.tid < O_LOC) red_add64(c.acc + 1024 + O_LOC * c.j + c.tid, acc_word(item_sum(c.s_part, c.tid)));
    c.t_pub = clock64();
    trace_sub<TR>(c, 5);
  }
  // Append the new K/V row (off the critical path).  EVERY CTA of the group writes it (they all hold the same bf16 values),
  // with strong (relaxed, gpu-scope) stores, and cached rows are read with strong loads: a CTA that reads the row in a later
  // step of the same launch is ordered after its OWN store by the CTA barriers in between, and whatever other store to that
  // address it observes instead carries the identical value -- no fence, no flag and no reliance on timing.  (A single writer
  // per group would need a release fence in front of its next publish, i.e. on the layer's critical path: +16 us per step.)
#ifdef QMK_EXP_KV_ONE_WRITER
  if (c.j == 0 && c.tid < 64) {
#else
  if (c.tid < 64) {
#endif
    const size_t off = ((size_t)(l * NKVH + c.g) * md.max_seq + position) * HD + c.tid * 2;
    const uint32_t kw = bf16_bits(s_small[SS_KN + c.tid * 2]) | (bf16_bits(s_small[SS_KN + c.tid * 2 + 1]) << 16);
    const uint32_t vw = bf16_bits(s_small[SS_V + c.tid * 2]) | (bf16_bits(s_small[SS_V + c.tid * 2 + 1]) << 16);
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(md.k_cache + off), "r"(kw) : "memory");
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(md.v_cache + off), "r"(vw) : "memory");
  }
}

// token sampling (sample_token2): csrc/qmk_sample.cuh

// ---- main loop -----------------------------------------------------------------------------------------------------
template <bool TR>
__device__ void consumer_loop2(Ctx2& c) {
  const Params& p = c.p;
  uint32_t* const x_q = reinterpret_cast<uint32_t*>(c.xb + XB_POOL) + (size_t)c.slot_q * XP_WORDS;
  uint32_t* const x_m = reinterpret_cast<uint32_t*>(c.xb + XB_POOL) + (size_t)c.slot_m * XP_WORDS;
  uint32_t* const x_logits = reinterpret_cast<uint32_t*>(c.xb + XB_LOGITS);
  u64* const accA = c.acc, * const accB = c.acc + 1024;
  // Frame loop (device-autonomous generation): a chained launch whose predecessor already saw EOS does nothing at all
  // (checked before the first weight copy is issued, so there is nothing to drain).
  // Loop state that is touched once per frame lives in shared memory, not in registers (the kernel sits at the register limit).
  int* const s_frames_base = reinterpret_cast<int*>(c.smem0 + S2_MISC + 192);
  if (p.frames.n_frames > 0 && p.frames.gen_state != nullptr) {
    if (p.frames.gen_state[1] != 0) return;
    if (c.tid == 0) *s_frames_base = p.frames.gen_state[0];
  } else if (c.tid == 0) {
    *s_frames_base = 0;
  }
  Prod2 prod;
  prod2_init(c, prod);
  if (c.warp >= 2) prod2_issue(c, prod, NSL);
  KvRegs2 kv;
#define QMK2_S_CODES (reinterpret_cast<int*>(c.smem0 + S2_MISC + 128))   /* int[16]: the codes of the current frame, selected by this CTA */
  const int gi0 = c.tid * 4;
  // totals of this thread's accumulator words at the end of the previous launch
  // (kept in shared memory, not in registers: 16 registers less across the whole kernel; they are read in the load shadow)
  u64* const s_prev = reinterpret_cast<u64*>(c.smem0 + S2_PREV);
  {
    const u64* snap = reinterpret_cast<const u64*>(c.xb + XB_SNAP);
#pragma unroll
    for (int e = 0; e < 4; ++e) { s_prev[gi0 + e] = snap[gi0 + e]; s_prev[1024 + gi0 + e] = snap[1024 + gi0 + e]; }
  }
  int next_token = 0;   // token selected at the end of the previous step of this launch
  float res[4] = {0.f, 0.f, 0.f, 0.f};   // residual stream, elements 4 tid .. 4 tid + 3 (every CTA holds all of it)
  // ldmatrix roles of this lane
  const int a_mi = c.lane >> 3;
  const int a_row = (c.lane & 7) + (a_mi & 1) * 8;
  const int a_sw = a_row & 7, a_khalf = a_mi >> 1;
  const int g8 = c.lane >> 2, q4 = c.lane & 3;

  int frame = 0;
  for (; frame < (p.frames.n_frames > 0 ? p.frames.n_frames : 1); ++frame) {
  if (p.frames.n_frames > 0) {
    // ---- frame prologue: EOS check on the token every CTA selected itself (first frame: the preceding launch's token) ----
    int tok0 = next_token;
    if (frame == 0) {
      tok0 = p.code0_ptr ? *p.code0_ptr : p.code0;
      const int hi = p.sum_rows0 - 1;
      tok0 = tok0 < 0 ? 0 : (tok0 > hi ? hi : tok0);
      next_token = tok0;
    }
    if (p.frames.eos_token >= 0 && tok0 == p.frames.eos_token) break;
    if (c.tid == 0) QMK2_S_CODES[0] = tok0;
    if (c.cta == 0 && c.tid == 0 && p.frames.codes_out != nullptr) p.frames.codes_out[(size_t)frame * 16] = (long long)tok0;
    c.t0 = clock64();   // the watchdog budget is per frame
  } else if (c.cta == 0 && c.tid == 0 && p.code0_out != nullptr) {
    *p.code0_out = (long long)(p.code0_ptr ? *p.code0_ptr : p.code0);
  }
  for (int step = 0; step < p.n_steps; ++step) {
    const StepDesc& sd = p.steps[step];
    const ModelDesc& md = p.models[sd.model];
    const int L = md.L;
    // epochs ebase + 1 .. ebase + L + 1 belong to this step
    const uint32_t ebase = p.epoch_base + (uint32_t)frame * (uint32_t)p.frames.epochs_per_frame + (uint32_t)sd.epoch_off;
    const int fadv = frame * sd.pos_per_frame;
    const int position = sd.position + fadv;
    const int in_mode = (frame > 0 && sd.in_mode_next >= 0) ? sd.in_mode_next : sd.in_mode;
    int in_token = sd.token;
    if (in_mode == IN_TABLE_TOKEN && sd.token_ptr != nullptr) {
      in_token = *sd.token_ptr;
      in_token = in_token < 0 ? 0 : (in_token > sd.token ? sd.token : in_token);
    }
    if (in_mode == IN_TABLE_PREV) {   // selected by THIS CTA at the end of an earlier step of this launch (see K2_ARGMAX)
      in_token = next_token;
      in_token = in_token < 0 ? 0 : ((sd.token > 0 && in_token > sd.token) ? sd.token : in_token);   // sd.token = last row of the table
    }
    // frame loop: the talker step's extra embedding is this frame's trailing-text row, or the pad embedding after the text
    const void* in_vec = sd.in_vec;
    if (p.frames.n_frames > 0 && sd.pos_per_frame != 0) {
      const int ti = p.frames.trailing_offset + *s_frames_base + frame;
      in_vec = (p.frames.trailing != nullptr && ti < p.frames.n_trailing) ? (const void*)(p.frames.trailing + (size_t)ti * H) : (const void*)p.frames.pad_embed;
    }
    const __nv_bfloat16* x_in = (in_mode == IN_TABLE_TOKEN || in_mode == IN_TABLE_PREV)
                                    ? sd.in_table + (size_t)in_token * H
                                    : reinterpret_cast<const __nv_bfloat16*>(in_vec);
    // the step's input row is requested before the RoPE row below, so that the two L2 / HBM round trips overlap
    uint2 xin_pre = make_uint2(0, 0);
    if (in_mode == IN_TABLE_TOKEN || in_mode == IN_TABLE_PREV || in_mode == IN_VEC_BF16)
      xin_pre = *reinterpret_cast<const uint2*>(x_in + gi0);
    else if (in_mode == IN_PREV_NORM)   // this thread's own four elements of the previous step's normalised hidden
      xin_pre = *reinterpret_cast<const uint2*>(c.s_vec + gi0 * 2);
    const AttnItem2 item = attn_item2(position, c.j);
    const int hrows_loc = sd.head.rows > 0 ? sd.head.rows / G2 : 0;
    if (c.tid < 128) {  // RoPE row of this step; with M-RoPE every rotary frequency takes the table row of its axis' position
      const int d = c.tid & 63;
      const int axis = (int)((md.rope_axis[d >> 5] >> (2 * (d & 31))) & 3ull);
      const __nv_bfloat16* t = (c.tid < 64) ? md.cos_t : md.sin_t;
      c.s_small[SS_CS + c.tid] = __bfloat162float(t[(size_t)(sd.rope_pos[axis] + fadv) * HD + d]);
    }
    consumer_bar();

    const int n_idx = L * PH_PER_LAYER + 2;
    for (int idx = 0; idx < n_idx; ++idx) {
      c.cur_idx = idx;
      trace_sub<TR>(c, 0);
      const int l = idx / PH_PER_LAYER;
      const int kind = idx < L * PH_PER_LAYER ? idx % PH_PER_LAYER : (idx == L * PH_PER_LAYER ? K2_HEAD : K2_ARGMAX);
      const uint32_t epoch = (ebase + 1u + (uint32_t)l) & 0xffffu;

      if (kind == K2_ATTN) {
        phase_attn2<TR>(c, md, l, position, epoch, item, kv, prod, a_row, a_khalf, a_sw);
        ++idx;   // the O projection (index K2_O) ran inside phase_attn2
        continue;
      }
      if (QMK_UNLIKELY(kind == K2_ARGMAX)) {
        // Token selection.  Inside a multi-step launch EVERY CTA gathers the logits and selects the token itself
        // (same words, same code, fixed summation orders -> the same token everywhere): the next step's input is
        // known without a gather-to-one-CTA plus a broadcast, i.e. one exchange per step instead of two.  The last step
        // of a launch is selected by CTA 0 alone.
        if (sd.head.rows <= 0 || (c.cta != 0 && step + 1 >= p.n_steps && frame + 1 >= p.frames.n_frames)) continue;
        const int hrows = sd.head.rows;
        const uint32_t epoch_head = (ebase + (uint32_t)L + 1u) & 0xffffu;
        const bool sample = sd.select != 0 && hrows <= NCW * 2 * HD && (hrows % (NCT * 4)) == 0;
        float* s_log = c.s_acc;
        float best = -INFINITY;
        int best_i = 0x7fffffff;
        bool retried = false;
        wait_window(c, c.s_delay[DL_ARGMAX]);
        for (int i = c.tid * 4; i < hrows; i += NCT * 4) {
          const uint4 w = ll4_wait(c, x_logits + i, epoch_head, retried);
          const float4 v = make_float4(ll4_val(w.x), ll4_val(w.y), ll4_val(w.z), ll4_val(w.w));
          if (sd.logits_out != nullptr && c.cta == 0) *reinterpret_cast<float4*>(sd.logits_out + i) = v;
          if (sample) *reinterpret_cast<float4*>(s_log + i) = v;
          const float v4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (v4[e] > best) { best = v4[e]; best_i = i + e; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
          if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        if (c.lane == 0) {
          c.s_red[c.warp * 2] = best;
          c.s_red[c.warp * 2 + 1] = __int_as_float(best_i);
        }
        gather_bar(c, DL_ARGMAX, retried);
#pragma unroll
        for (int w = 0; w < NCW; ++w) {
          const float ov = c.s_red[w * 2];
          const int oi = __float_as_int(c.s_red[w * 2 + 1]);
          if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        if (QMK_UNLIKELY(best_i == 0x7fffffff)) best_i = 0;   // all-NaN logits: never hand an out-of-range index to a table lookup
        int chosen = best_i;
        if (sample)
          chosen = sample_token2(s_log, reinterpret_cast<unsigned*>(c.s_part), c.s_red, c.tid, c.warp, c.lane, hrows,
                                p.sample_top_k, p.sample_temperature, p.sample_seed, p.sample_counter + (unsigned long long)frame, sd.group, best, best_i);
        next_token = chosen;
        if (p.forced_tokens != nullptr && sd.group >= 0) next_token = p.forced_tokens[sd.group];   // teacher forcing
        if (c.tid == 0 && sd.group >= 0) QMK2_S_CODES[sd.group + 1] = chosen;
        if (c.cta == 0 && c.tid == 0) {
          const int st = *((volatile int*)p.status);
          const int out = (st != 0 || *c.s_abort) ? -1000 - st : chosen;
          if (sd.out_token != nullptr) *sd.out_token = out;
          if (p.frames.n_frames > 0) {
            if (sd.group >= 0 && p.frames.codes_out != nullptr) p.frames.codes_out[(size_t)frame * 16 + sd.group + 1] = (long long)out;
            if (sd.pos_per_frame != 0 && p.frames.tokens_out != nullptr) p.frames.tokens_out[frame] = out;
          } else if (sd.out_code != nullptr) {
            *sd.out_code = (long long)out;
          }
        }
        consumer_bar();
        continue;
      }

      // ---- GEMV-shaped phases ----------------------------------------------------------------------------------
      // Shape of a phase: `ntiles` tiles of 16 rows; each warp owns a K slice of 16 nk (nk tensor-core steps per tile).
      // Everything that does not depend on the gathered activations runs BEFORE they are checked (the A fragments move
      // from the ring into registers while the exchange is in flight): a single warp executes dependent instructions
      // at ~5+ cycles each, so every instruction between "data arrived" and "published" is on the layer's critical path.
      int ntiles, nst;
      uint32_t stride;
      if (kind == K2_QKV) { ntiles = 2; nst = 2; stride = 2048; }
      else if (kind == K2_GU) { ntiles = 3; nst = 3; stride = 2048; }
      else if (kind == K2_DOWN) { ntiles = 4; nst = 2; stride = 768; }
      else { ntiles = (hrows_loc + 15) / 16; nst = ntiles; stride = 2048; }
      const bool norm = (kind == K2_QKV || kind == K2_GU || kind == K2_HEAD);
      const bool from_input = (kind == K2_QKV && l == 0);
      const bool useB = (kind == K2_GU);
      u64* const acc = useB ? accB : accA;
      const int dslot = kind == K2_GU ? DL_GU : (kind == K2_QKV ? DL_QKV : (kind == K2_DOWN ? DL_DOWN : DL_HEAD));
      const uint8_t* nw_ptr = (kind == K2_HEAD) ? sd.head.aux : md.aux_layers + ((size_t)l * 2 + (kind == K2_GU ? 1 : 0)) * AUX_BYTES;

      // ---- 1. weights of this phase -> registers, BEFORE the wait window: the warps that neither finalise nor publish sit
      //         at the window's barrier for ~600 cycles anyway, and the 64-96 KB of shared-memory reads (500-750 cycles) would
      //         otherwise outlast the round trip of the gather loads they are meant to hide behind ----
      uint32_t tb[4];   // shared-memory address of this lane's ldmatrix row in tile t
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int slot_i = (kind == K2_DOWN) ? (t >> 1) : (t < nst ? t : 0);
        const uint32_t toff = (kind == K2_DOWN) ? (uint32_t)(t & 1) * 12288u : 0u;
        tb[t] = smem_u32(c.ring + (size_t)((c.k + slot_i) % NSL) * SLOT2) + toff + (uint32_t)a_row * stride;
      }
      {   // three phase tests back to back (their latencies overlap); a stage that is not resident yet is waited for
        const uint32_t k0 = c.k, k1 = c.k + (nst > 1 ? 1 : 0), k2 = c.k + (nst > 2 ? 2 : 0);
        const uint32_t ready = mbar_test_wait3(&c.full[k0 % NSL], (k0 / NSL) & 1u, &c.full[k1 % NSL], (k1 / NSL) & 1u,
                                               &c.full[k2 % NSL], (k2 / NSL) & 1u);
        if (QMK_UNLIKELY(ready != 7u)) {
#pragma unroll 1
          for (int sidx = 0; sidx < nst; ++sidx) wait_full(c, c.k + sidx);
        }
      }
      uint32_t apre[16][4];   // A fragments (weights) held in registers across the exchange wait
      if (kind == K2_DOWN) preload_a<4, 3, 4>(apre, tb, c.warp, a_khalf, a_sw);
      else preload_a<3, 8, 2>(apre, tb, c.warp, a_khalf, a_sw);   // QKV / head: both tiles; gate/up: tiles 0, 1
      // ---- 2. wait window, then issue the gather loads ----
      u64 now[4];
      uint4 mw = make_uint4(0, 0, 0, 0);
      if (!from_input) {
        wait_window(c, c.s_delay[dslot]);
        trace_sub<TR>(c, 1);
        if (norm) {
          ld_acc2(acc + gi0, now[0], now[1]);
          ld_acc2(acc + gi0 + 2, now[2], now[3]);
        } else if (c.lane < 12) {   // down: warp w gathers its own K slice m[48 w .. 48 w + 48) of the group's 384 values
          mw = ll4_ld4(x_m + c.warp * 48 + c.lane * 4);
        }
      }
      // shadow of the load latency: norm weights, L2 prefetches
      uint2 wv = make_uint2(0, 0);
      if (norm) wv = *reinterpret_cast<const uint2*>(nw_ptr + c.tid * 8);
      // Older KV rows of this CTA's attention item -> L2 now (fire and forget).  The register loads themselves are issued at
      // the start of the attention phase: loads that miss L2 share hardware scoreboards with the exchange loads when
      // they are in flight together, and the data check below would wait for them too.
      if (kind == K2_QKV) {
        if (item.has) attn_l2_prefetch(c, md, l, position, item);
        if (c.cta == ((l + 1) & (G2 - 1)) && c.tid < 40 && l + 1 < L)   // next layer's norm weights (one CTA per layer; 40 lines)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(md.aux_layers + ((size_t)(l + 1) * 2) * AUX_BYTES + c.tid * 128));
      }
      trace_sub<TR>(c, 2);

      // ---- 3. data ----
      if (norm) {
        bool retried = false;
        if (from_input) {
          if (in_mode == IN_CODES_SUM) {
            // 16 code indices, then 16 embedding rows: two rounds of independent loads (a serial chain of 31 dependent
            // global loads costs ~10 us per talker step), then the bf16 adds in the upstream order.  The indices come from
            // device memory (a preceding launch's codes) or, inside the frame loop, from this CTA's own selections; they are
            // clamped to the tables' rows (an aborted predecessor writes negative sentinels).
            int code[16];
#pragma unroll
            for (int g = 0; g < 16; ++g) {
              const int v = sd.codes != nullptr ? (int)sd.codes[g] : QMK2_S_CODES[g];
              const int hi = (g == 0 ? p.sum_rows0 : p.sum_rows) - 1;
              code[g] = v < 0 ? 0 : (v > hi ? hi : v);
            }
            uint2 row[16];
            row[0] = *reinterpret_cast<const uint2*>(sd.in_table + (size_t)code[0] * H + gi0);
#pragma unroll
            for (int g = 0; g < 15; ++g) row[g + 1] = *reinterpret_cast<const uint2*>(p.sum_tables[g] + (size_t)code[g + 1] * H + gi0);
            const uint2 vx = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(in_vec) + gi0);
            float e4[4] = {bf16_lo(row[0].x), bf16_hi(row[0].x), bf16_lo(row[0].y), bf16_hi(row[0].y)};
#pragma unroll
            for (int g = 1; g < 16; ++g) {
              e4[0] = bf16_round(e4[0] + bf16_lo(row[g].x)); e4[1] = bf16_round(e4[1] + bf16_hi(row[g].x));
              e4[2] = bf16_round(e4[2] + bf16_lo(row[g].y)); e4[3] = bf16_round(e4[3] + bf16_hi(row[g].y));
            }
            res[0] = bf16_round(e4[0] + bf16_lo(vx.x)); res[1] = bf16_round(e4[1] + bf16_hi(vx.x));
            res[2] = bf16_round(e4[2] + bf16_lo(vx.y)); res[3] = bf16_round(e4[3] + bf16_hi(vx.y));
          } else if (in_mode == IN_VEC_F32) {
            const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in_vec) + gi0);
            res[0] = bf16_round(v.x); res[1] = bf16_round(v.y); res[2] = bf16_round(v.z); res[3] = bf16_round(v.w);
          } else {
            const uint2 v = xin_pre;
            res[0] = bf16_lo(v.x); res[1] = bf16_hi(v.x); res[2] = bf16_lo(v.y); res[3] = bf16_hi(v.y);
          }
        } else {
          u64 prev[4];
          {
            const ulonglong2 q0 = *reinterpret_cast<const ulonglong2*>(s_prev + (useB ? 1024 : 0) + gi0);
            const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(s_prev + (useB ? 1024 : 0) + gi0 + 2);
            prev[0] = q0.x; prev[1] = q0.y; prev[2] = q1.x; prev[3] = q1.y;
          }
          if (QMK_UNLIKELY(!(acc_done(now[0], prev[0]) & acc_done(now[1], prev[1]) & acc_done(now[2], prev[2]) & acc_done(now[3], prev[3])))) {
            retried = true;
            acc_wait_slow(p.status, c.s_abort, c.t0, p.timeout_cycles, c.cta, idx, acc + gi0, prev[0], prev[1], prev[2], prev[3]);
            ld_acc2(acc + gi0, now[0], now[1]);
            ld_acc2(acc + gi0 + 2, now[2], now[3]);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float o = bf16_round(acc_value(now[e], prev[e]));
            res[e] = md.residual_fp32 ? res[e] + o : bf16_round(res[e] + o);
          }
          *reinterpret_cast<ulonglong2*>(s_prev + (useB ? 1024 : 0) + gi0) = make_ulonglong2(now[0], now[1]);
          *reinterpret_cast<ulonglong2*>(s_prev + (useB ? 1024 : 0) + gi0 + 2) = make_ulonglong2(now[2], now[3]);
        }
        const float r0 = bf16_round(res[0]), r1 = bf16_round(res[1]), r2 = bf16_round(res[2]), r3 = bf16_round(res[3]);
        float ss = fmaf(r0, r0, r1 * r1) + fmaf(r2, r2, r3 * r3);
        ss = warp_sum(ss);
        if (c.lane == 0) c.s_red[c.warp] = ss;
        if (from_input) consumer_bar(); else gather_bar(c, dslot, retried);
        trace_sub<TR>(c, 3);
        const float4 s0 = *reinterpret_cast<const float4*>(c.s_red), s1 = *reinterpret_cast<const float4*>(c.s_red + 4);
        const float inv = rsqrtf((((s0.x + s0.y) + (s0.z + s0.w)) + ((s1.x + s1.y) + (s1.z + s1.w))) * (1.0f / H) + EPS);
        const __nv_bfloat162 n01 = __floats2bfloat162_rn((r0 * inv) * bf16_lo(wv.x), (r1 * inv) * bf16_hi(wv.x));
        const __nv_bfloat162 n23 = __floats2bfloat162_rn((r2 * inv) * bf16_lo(wv.y), (r3 * inv) * bf16_hi(wv.y));
        const uint2 packed = make_uint2(*reinterpret_cast<const uint32_t*>(&n01), *reinterpret_cast<const uint32_t*>(&n23));
        *reinterpret_cast<uint2*>(c.s_vec + gi0 * 2) = packed;
        if (kind == K2_HEAD && c.cta == 0) {
          if (sd.hidden_out != nullptr) {
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(r0, r1), h23 = __floats2bfloat162_rn(r2, r3);
            *reinterpret_cast<uint2*>(sd.hidden_out + gi0) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
          }
          if (sd.out_norm != nullptr)
            *reinterpret_cast<float4*>(sd.out_norm + gi0) = make_float4(bf16_lo(packed.x), bf16_hi(packed.x), bf16_lo(packed.y), bf16_hi(packed.y));
        }
      } else if (kind == K2_DOWN) {
        bool retried = false;
        if (c.lane < 12) {
          if (QMK_UNLIKELY(!ll4_ok(mw, epoch))) {
            retried = true;
            mw = ll4_wait_slow(p.status, c.s_abort, c.t0, p.timeout_cycles, c.cta, idx, x_m + c.warp * 48 + c.lane * 4, epoch, c.lane);
          }
          *reinterpret_cast<uint2*>(c.s_a + (c.warp * 48 + c.lane * 4) * 2) = make_uint2((mw.x & 0xffffu) | (mw.y << 16), (mw.z & 0xffffu) | (mw.w << 16));
        }
        gather_note(c, DL_DOWN, retried);
        if (c.tid == 0) c.s_delay[2 * DL_N + DL_DOWN] += (int)((clock64() - c.t_pub) >> 4);   // ... and of its m buffer
      }
      __syncwarp();   // warp w consumes exactly the K slice its own lanes wrote 
      trace_sub<TR>(c, 4);
      if (kind == K2_HEAD && ntiles == 0) continue;   // step without an LM head: only the final norm / outputs

      // ---- 4. tensor-core stages ----
      float acc4[4][4];
      if (kind == K2_DOWN) mma_tiles<4, 3, 4>(acc4, apre, tb, c.s_a, 4, c.warp, q4, a_khalf, a_sw);
      else mma_tiles<3, 8, 2>(acc4, apre, tb, c.s_vec, ntiles, c.warp, q4, a_khalf, a_sw);
      c.k += nst;
      if (TR) { if (__float_as_uint(acc4[0][0]) == 0x7fc12345u) c.s_red[40] = 1.f; }   // stamp 5 after the tensor-core results exist
      trace_sub<TR>(c, 5);
      // per-warp (K-slice) partials of rows 16 t + g8 and 16 t + g8 + 8 (every B column is the same vector: take column 0)
      if (q4 == 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (t < ntiles) {
            c.s_part[((t * 16 + g8) * NCW) + c.warp] = acc4[t][0];
            c.s_part[((t * 16 + g8 + 8) * NCW) + c.warp] = acc4[t][2];
          }
        }
      }
      consumer_bar();
      trace_sub<TR>(c, 6);
      if (c.warp >= 2) prod2_issue(c, prod, nst);

      // ---- 5. finalize + publish: the cross-warp sum is common to all kinds (rows beyond the phase's count hold stale
      //         partials and are not published), the kind only selects the tail ----
      if (c.warp < 2) {
        const float v = item_sum(c.s_part, c.tid);
        if (kind == K2_DOWN) {          // one fixed-point add per row: the eight groups' K-split partials meet in L2
          red_add64(accA + O_LOC * c.j + c.tid, acc_word(v));
        } else if (kind == K2_GU) {     // gate in even lanes, up in the next lane
          const float u = bf16_round(__shfl_down_sync(0xffffffffu, v, 1));
          if (c.tid < GU_LOC && (c.tid & 1) == 0) {
            const float gt = bf16_round(v);
            const float sg = bf16_round(__fdividef(gt, 1.0f + __expf(-gt)));
            ll4_st(x_m + M_LOC * c.j + (c.tid >> 1), sg * u, epoch);
          }
        } else if (kind == K2_QKV) {    // local rows: 0..15 q (16 j + r of the head pair), 16..23 k, 24..31 v
          const int r = c.tid;
          const int w = r < 16 ? GSZ * c.j + r : (r < 24 ? 256 + 8 * c.j + (r - 16) : 384 + 8 * c.j + (r - 24));
          if (r < QKV_LOC) ll4_st(x_q + w, v, epoch);
        } else {                        // head: logits as LL4 words
          if (c.tid < hrows_loc) ll4_st(x_logits + hrows_loc * c.cta + c.tid, v, (ebase + (uint32_t)L + 1u) & 0xffffu);
        }
      }
      c.t_pub = clock64();
      trace_sub<TR>(c, 8);
    }
  }
  if (p.frames.n_frames > 0 && c.cta == 0 && c.tid == 0 && p.frames.gen_state != nullptr) {
    // progress word for a polling host (the buffers may be mapped host memory): codes and tokens first, then the count
    __threadfence_system();
    *((volatile int*)p.frames.gen_state) = *s_frames_base + frame + 1;
    ((volatile int*)p.frames.gen_state)[2] = next_token;
  }
  }
  if (p.frames.n_frames > 0 && frame < p.frames.n_frames) {
    // EOS: the weight copies already issued for the next frame must land before the CTA may exit
    consumer_bar();
    if (c.warp >= 2) {
      for (uint32_t k = c.k; k != prod.k; ++k) wait_full(c, k);
    }
    if (c.cta == 0 && c.tid == 0 && p.frames.gen_state != nullptr) ((volatile int*)p.frames.gen_state)[1] = 1;
  }
  // totals for the next launch
  if (c.cta == 0) {
    u64* snap = reinterpret_cast<u64*>(c.xb + XB_SNAP);
#pragma unroll
    for (int e = 0; e < 4; ++e) { snap[gi0 + e] = s_prev[gi0 + e]; snap[1024 + gi0 + e] = s_prev[1024 + gi0 + e]; }
  }
}

template <bool TR>
__device__ __forceinline__ void decode2_body(const Params& p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctx2 c(p);
  c.smem0 = smem;
  c.ring = smem + S2_RING;
  c.s_vec = smem + S2_VEC;
  c.s_a = smem + S2_A;
  c.s_acc = reinterpret_cast<float*>(smem + S2_ACC);
  c.s_small = reinterpret_cast<float*>(smem + S2_SMALL);
  c.s_part = reinterpret_cast<float*>(smem + S2_PART);
  c.s_red = reinterpret_cast<float*>(smem + S2_RED);
  c.full = reinterpret_cast<u64*>(smem + S2_BAR);
  c.s_abort = reinterpret_cast<volatile int*>(smem + S2_MISC);
  c.s_delay = reinterpret_cast<int*>(smem + S2_MISC + 16);
  c.s_tbl = nullptr;
  c.xb = p.xbuf;
  c.x32 = reinterpret_cast<uint32_t*>(p.xbuf + XB_LL);
  c.acc = reinterpret_cast<u64*>(p.xbuf + XB_ACC);
  c.tid = threadIdx.x;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  // Role of this CTA (which group, which rank): a host-built permutation (identity by default; see qmk_engine_create for the
  // measured effect of SM-sorted groups).  Any permutation is correct; it only moves latency.
  c.cta = reinterpret_cast<const int*>(p.xbuf + XB_ROLE)[blockIdx.x];
  c.g = c.cta / GSZ;
  c.j = c.cta % GSZ;
  c.slot_q = reinterpret_cast<const int*>(p.xbuf + XB_SLOTS)[2 * c.g];
  c.slot_m = reinterpret_cast<const int*>(p.xbuf + XB_SLOTS)[2 * c.g + 1];
  c.k = 0;
  c.t0 = clock64();
  c.t_pub = c.t0;
  c.cur_idx = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSL; ++i) mbar_init(&c.full[i], 1);
    *c.s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 3 * DL_N) c.s_delay[threadIdx.x] = p.delays[blockIdx.x * 3 * DL_N + threadIdx.x];
  __syncthreads();
  consumer_loop2<TR>(c);
  __syncthreads();
  if (TR && p.trace != nullptr && threadIdx.x == 0) {   // end stamp + the SM this CTA ran on
    const int n_idx = p.lay.L * PH_PER_LAYER + 2;
    if ((n_idx + 1) * TRACE_SUBS <= p.trace_stride) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[(size_t)c.cta * p.trace_stride + n_idx * TRACE_SUBS + 0] = clock64();
      p.trace[(size_t)c.cta * p.trace_stride + n_idx * TRACE_SUBS + 1] = (long long)smid;
    }
  }
  if (threadIdx.x >= DL_N && threadIdx.x < 3 * DL_N) p.delays[blockIdx.x * 3 * DL_N + threadIdx.x] = c.s_delay[threadIdx.x];
  if (*c.s_abort && threadIdx.x == 0) {
    const long long t = clock64();
    while (clock64() - t < 2000000) {
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) qmk2_decode_kernel(const __grid_constant__ Params p) { decode2_body<false>(p); }
__global__ void __launch_bounds__(NTHREADS, 1) qmk2_decode_kernel_traced(const __grid_constant__ Params p) { decode2_body<true>(p); }

// Calibration of the group-buffer placement: in step t group g exchanges `rounds` times through pool slot (t + 5 g) % 48
// (all groups on different slots, like in the decode kernel); out[g][slot] = cycles.  A grid barrier separates the steps.
__global__ void __launch_bounds__(NTHREADS, 1) qmk2_calib_kernel(uint32_t* pool, int rounds, const int* role, unsigned* bar, long long* out) {
  extern __shared__ __align__(16) uint8_t smem_calib[];
  const int cta = role[blockIdx.x], g = cta / GSZ, j = cta % GSZ, tid = threadIdx.x;
  unsigned sink = 0;
  for (int t = 0; t < XP_CAND; ++t) {
    const int slot = (t + 5 * g) % XP_CAND;
    uint32_t* buf = pool + (size_t)slot * XP_WORDS;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      const uint32_t epoch = (uint32_t)(1 + t * rounds + r);
      if (tid < 32) ll4_st(buf + 32 * j + tid, 1.0f, epoch);
      if (tid < 128) {
        uint4 w;
        unsigned spins = 0;
        do { w = ll4_ld4(buf + 4 * tid); } while (!ll4_ok(w, epoch) && ++spins < (1u << 10));
        if (spins >= (1u << 10)) atomicAdd(bar + 1, 1u);   // give up (counted): a stuck poll must not hang engine creation
        sink += w.x;
      }
      __syncthreads();
    }
    if (tid == 0 && j == 0) out[g * XP_CAND + slot] = clock64() - t0;
    // grid barrier (all 128 CTAs are co-resident: cooperative launch)
    if (tid == 0) {
      __threadfence();
      atomicAdd(bar, 1u);
      unsigned spins = 0;
      while (*((volatile unsigned*)bar) < (unsigned)(G2 * (t + 1)) && ++spins < (1u << 16)) {
      }
    }
    __syncthreads();
  }
  if (sink == 0x12345u) smem_calib[0] = 1;
}

// records the SM id of every CTA of a launch with the decode kernel's shape (grid 128, 256 threads, same shared memory)
__global__ void __launch_bounds__(NTHREADS, 1) qmk2_probe_smid_kernel(int* out) {
  extern __shared__ __align__(16) uint8_t smem_probe[];
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    out[blockIdx.x] = (int)smid + (smem_probe[0] & 0);
  }
}

// ---- weight re-packing for the group layout -----------------------------------------------------------------------
// 16-byte chunk `ch` of a packed row whose K block starts at 32-bit word `kw0` of the source row: slice w = ch / (2 nk),
// m = ch % (2 nk); the chunk holds, for q = 0..3, source words  w * 8 nk + 2 nk q + m.
__device__ __forceinline__ uint4 pack_chunk2(const uint32_t* row_words, int kw0, int nk, int ch) {
  const int w = ch / (2 * nk), m = ch % (2 * nk);
  const uint32_t* s = row_words + kw0 + w * 8 * nk + m;
  return make_uint4(s[0], s[2 * nk], s[4 * nk], s[6 * nk]);
}
// grid = (128, L), block = 128
__global__ void qmk2_pack_layers_kernel(const LayerPtrs* layers, int L, uint8_t* packed, uint8_t* aux_layers) {
  const int cta = blockIdx.x, l = blockIdx.y, t = threadIdx.x;
  const int g = cta / GSZ, j = cta % GSZ;
  const LayerPtrs lp = layers[l];
  uint8_t* dst = packed + ((size_t)cta * L + l) * LAYER_BYTES2;
  const uint32_t* wq = reinterpret_cast<const uint32_t*>(lp.w[W_Q]);
  const uint32_t* wk = reinterpret_cast<const uint32_t*>(lp.w[W_K]);
  const uint32_t* wv = reinterpret_cast<const uint32_t*>(lp.w[W_V]);
  const uint32_t* wo = reinterpret_cast<const uint32_t*>(lp.w[W_O]);
  const uint32_t* wg = reinterpret_cast<const uint32_t*>(lp.w[W_GATE]);
  const uint32_t* wu = reinterpret_cast<const uint32_t*>(lp.w[W_UP]);
  const uint32_t* wd = reinterpret_cast<const uint32_t*>(lp.w[W_DOWN]);
  // QKV: 32 rows x 2 KB (128 chunks per row)
  for (int r = 0; r < QKV_LOC; ++r) {
    const uint32_t* src = r < 16 ? wq + (size_t)(2 * g * HD + GSZ * j + r) * 512
                                 : (r < 24 ? wk + (size_t)(g * HD + 8 * j + (r - 16)) * 512 : wv + (size_t)(g * HD + 8 * j + (r - 24)) * 512);
    uint4* d = reinterpret_cast<uint4*>(dst + (size_t)(r >> 4) * SLOT2 + (size_t)(r & 15) * 2048);
    d[t ^ (r & 7)] = pack_chunk2(src, 0, 8, t);
  }
  // O: 64 rows x 512 B (32 chunks per row), K block = columns 256 g ..
  for (int r = 0; r < O_LOC; ++r) {
    const uint32_t* src = wo + (size_t)(O_LOC * j + r) * 1024;
    uint4* d = reinterpret_cast<uint4*>(dst + 2 * (size_t)SLOT2 + (size_t)r * 512);
    if (t < 32) d[t ^ (r & 7)] = pack_chunk2(src, KB_O * g / 2, 2, t);
  }
  // gate/up: 48 rows, gate_i and up_i interleaved
  for (int r = 0; r < GU_LOC; ++r) {
    const int row = KB_D * g + M_LOC * j + (r >> 1);
    const uint32_t* src = ((r & 1) ? wu : wg) + (size_t)row * 512;
    uint4* d = reinterpret_cast<uint4*>(dst + (size_t)(3 + (r >> 4)) * SLOT2 + (size_t)(r & 15) * 2048);
    d[t ^ (r & 7)] = pack_chunk2(src, 0, 8, t);
  }
  // down: 64 rows x 768 B (48 chunks per row), K block = columns 384 g ..; two 24 KB slots of 32 rows
  for (int r = 0; r < O_LOC; ++r) {
    const uint32_t* src = wd + (size_t)(O_LOC * j + r) * 1536;
    uint4* d = reinterpret_cast<uint4*>(dst + 6 * (size_t)SLOT2 + (size_t)(r >> 5) * DOWN_SLOT + (size_t)(r & 31) * 768);
    if (t < 48) d[t ^ (r & 7)] = pack_chunk2(src, KB_D * g / 2, 3, t);
  }
  if (cta == 0) {  // shared norm-weight blocks, same format as the first kernel
    uint4* a0 = reinterpret_cast<uint4*>(aux_layers + ((size_t)l * 2 + 0) * AUX_BYTES);
    uint4* a1 = reinterpret_cast<uint4*>(aux_layers + ((size_t)l * 2 + 1) * AUX_BYTES);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    a0[t] = lp.w[W_IN][t];
    a1[t] = lp.w[W_POST][t];
    if (t < 16) {
      a0[128 + t] = lp.w[W_QN][t];
      a0[144 + t] = lp.w[W_KN][t];
      a1[128 + t] = zero;
      a1[144 + t] = zero;
    }
  }
}
// grid = 128, block = 128: rows_loc = rows / 128 rows of this CTA, padded to whole 16-row tiles
__global__ void qmk2_pack_head_kernel(const uint4* head_w, int rows, uint8_t* packed) {
  const int cta = blockIdx.x, t = threadIdx.x;
  const int rows_loc = rows / G2, tiles = (rows_loc + 15) / 16;
  uint8_t* dst = packed + (size_t)cta * tiles * SLOT2;
  for (int r = 0; r < tiles * 16; ++r) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < rows_loc) v = pack_chunk2(reinterpret_cast<const uint32_t*>(head_w) + (size_t)(rows_loc * cta + r) * 512, 0, 8, t);
    reinterpret_cast<uint4*>(dst + (size_t)(r >> 4) * SLOT2 + (size_t)(r & 15) * 2048)[t ^ (r & 7)] = v;
  }
}

}  // namespace qmk2
