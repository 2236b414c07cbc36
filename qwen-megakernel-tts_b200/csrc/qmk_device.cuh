// qmk_device.cuh — device side of the B200 decode engine (sm_100a).
//
// One persistent cooperative kernel, one CTA per SM, 8 warps per CTA:
//   * this CTA's slice of the re-packed weights streams through a 6-slot x 30.5 KB shared-memory ring with 1-D TMA
//     bulk copies (cp.async.bulk ... mbarrier::complete_tx).  Weight addresses do not depend on activations, so
//     the stream runs ~0.75 layer ahead through every data dependency and HBM stays busy while the CTAs wait
//     on each other; a slot is refilled by one lane as soon as the phase that read it has passed its barrier.
//   * a ring stage is a 14-row x 1024-k weight tile; each warp owns a 128-wide K slice of it and multiplies it
//     with the activation vector on the tensor cores (mma.sync m16n8k16, fp32 accumulate); the K-slice partials
//     are summed through shared memory in a fixed order.
// CTAs exchange activations through "LL" words in global memory: every exchanged vector is bf16-valued
// (the reference rounds to bf16 at exactly these points), so a 32-bit word carries {bf16 payload, 16-bit
// epoch}.  Data and flag arrive in one single-copy-atomic access: there is no grid barrier, no fence and no
// separate flag on the critical path, and a consumer gathers a whole vector with 16-byte loads (4 words).
// Measured on B200 (scripts/ubench/xchg*.cu): polling a line before it is written more than doubles the
// exchange latency (4500 vs 1700 cycles for 148 CTAs x 1024 words), so each CTA waits an ADAPTIVE delay
// after its own publish before the first poll; the delay grows when a poll had to be repeated and decays
// otherwise, and the wait window is used to test the weight-ring barriers of the next phase.
//
// Numerics follow the upstream *PyTorch* path, not upstream kernel.cu (see oracle/tts_oracle.py for the
// rounding points and the upstream file:line of each): bf16 rounding after every projection / norm /
// RoPE product / SwiGLU factor, fp32 accumulation, fp32 (talker) or bf16 (code predictor) residual.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define QMK_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define QMK_LIKELY(x) __builtin_expect(!!(x), 1)

namespace qmk {

typedef unsigned long long u64;

constexpr int H = 1024, INTER = 3072, QSZ = 2048, KVSZ = 1024, HD = 128, NQH = 16, NKVH = 8;
constexpr int QKV_ROWS = QSZ + 2 * KVSZ;  // 4096
constexpr int SEG_ELEMS = 1024, SEG_BYTES = 2048;
constexpr int NCW = 8;               // consumer warps: each owns 8 of the 64 k16-steps of a 1024-wide segment
constexpr int KSTEPS = 64 / NCW;     // (exactly 2 warps per SM sub-partition = 255 registers/thread: with ~10 KB of L1 left
                                     //  beside 200+ KB of shared memory, a single spilled register costs an L2 round trip)
constexpr int NCT = NCW * 32;        // 256 consumer threads
constexpr int NTHREADS = NCT;        // no dedicated producer warp: see Prod below
constexpr int STAGE_ITEMS = 14;      // 2 KB row segments per ring stage = rows of one m16 tensor-core tile
constexpr int AUX_BYTES = 2560;      // norm weights (2048) + q_norm / k_norm (2 x 256)
constexpr int SLOT_BYTES = AUX_BYTES + STAGE_ITEMS * SEG_BYTES;  // 31232: [aux][14 row segments, 16-byte chunks swizzled]
constexpr int NSLOTS = 6;
constexpr int MAX_ST = 3;            // stages per phase (42 gate/up segments on 148 CTAs = 3 stages)
constexpr int MAX_ITEMS = MAX_ST * STAGE_ITEMS;
constexpr int ATT_PER_WARP = 5;
constexpr int ATT_ROUND = NCW * ATT_PER_WARP;  // 40 cached positions per round per item
constexpr int S_MAX = 18;                      // max KV splits per kv head (8 * 18 = 144 CTAs)
constexpr int PART_STRIDE = 132;               // u64 words per (q head, split) partial: m, l, acc[128], pad
constexpr int MAX_HEAD_ROWS = 3072;
constexpr int MAX_STEPS = 17;              // a code-predictor frame (16 steps) + the talker step of the same frame
constexpr float EPS = 1e-6f;

enum Phase { PH_QKV = 0, PH_ATTN = 1, PH_O = 2, PH_GU = 3, PH_DOWN = 4, PH_PER_LAYER = 5 };
enum DelaySlot { DL_QKV = 0, DL_ATTN = 1, DL_O = 2, DL_GU = 3, DL_DOWN = 4, DL_HEAD = 5, DL_ARGMAX = 6, DL_TOKEN = 7, DL_N = 8 };
enum Status { ST_OK = 0, ST_TIMEOUT_LL = 1, ST_TIMEOUT_FULL = 2, ST_TIMEOUT_EMPTY = 3, ST_BAD_CONFIG = 4 };

// Exchange buffer: 32-bit LL words first, then the 64-bit LL words of the split-attention partials.
constexpr int XW_RES = 0, XW_QKV = XW_RES + H, XW_A = XW_QKV + QKV_ROWS, XW_RES2 = XW_A + QSZ, XW_M = XW_RES2 + H,
              XW_LOGITS = XW_M + INTER, XW_TOKEN = XW_LOGITS + MAX_HEAD_ROWS, XW_END32 = XW_TOKEN + 32;
constexpr int XW_PART64 = NQH * S_MAX * PART_STRIDE;
constexpr size_t XBUF_BYTES = (size_t)XW_END32 * 4 + (size_t)XW_PART64 * 8;
static_assert((XW_END32 * 4) % 16 == 0, "64-bit region must stay aligned");

// Segment layout of one (cta, layer) block of the packed weight stream, in 2 KB segments:
//   [qkv rows] | [o items (row,kb) kb<2] | [gate_j, up_j pairs] | [down items (row,kb) kb<3]
// The norm weights live once per layer in `aux_layers` ([L][2][AUX_BYTES]) and are fetched into the aux
// part of the first stage of the QKV / gate-up phases.
struct Layout {
  int G;        // CTAs (= SMs)
  int L;        // layers
  int qkv_max;  // ceil(4096 / G)
  int o_max;    // ceil(1024 / G)
  int gu_max;   // ceil(3072 / G)
  int off_o, off_gu, off_down, layer_segs;
};

struct HeadDesc {
  const uint8_t* packed;  // [G][segs_max][2048]: rows of this CTA; rows==0 -> no LM head
  const uint8_t* aux;     // AUX_BYTES: final RMSNorm weight
  int rows;
  int segs_max;
};

enum InMode { IN_TABLE_TOKEN = 0, IN_VEC_BF16 = 1, IN_TABLE_PREV = 2, IN_VEC_F32 = 3, IN_CODES_SUM = 4,
              IN_PREV_NORM = 5 };   // group kernel only: the previous step's post-final-norm hidden (kept on chip)
struct StepDesc {
  const __nv_bfloat16* in_table;  // IN_TABLE_TOKEN: row `token`; IN_TABLE_PREV: row = token selected by the previous step
  const void* in_vec;             // IN_VEC_BF16: bf16[1024]; IN_VEC_F32: f32[1024] (rounded to bf16 on load)
  const long long* codes;         // IN_CODES_SUM: int64[16] (device) -> in_table[codes[0]] + sum_g sum_tables[g][codes[g+1]] + in_vec
  int in_mode;
  int token;
  const int* token_ptr;           // IN_TABLE_TOKEN: if non-null the row index is read from device memory (clamped)
  int position;
  HeadDesc head;
  int select;                     // 0 = argmax, 1 = temperature / top-k / multinomial (Params::sample_*)
  int group;                      // code-predictor group of this head (index into forced_tokens), or -1
  int model;                      // group kernel: index into Params::models (0 = the launch's primary model)
  int in_mode_next;               // group kernel, frame loop: input mode of this step in frames after the first (-1 = in_mode)
  int pos_per_frame;              // group kernel, frame loop: position / rope_pos advance by this much per frame
  int rope_pos[3];                // RoPE table rows of the three M-RoPE axes (standard RoPE: all = position)
  int epoch_off;                  // group kernel: epochs used by the steps before this one (sum of L + 2)
  int* out_token;                 // int32[1] (head.rows > 0) or null
  long long* out_code;            // int64[1] or null
  float* logits_out;              // f32[head.rows] or null
  float* out_norm;                // f32[1024] post-final-norm hidden (bf16-rounded values) or null
  __nv_bfloat16* hidden_out;      // bf16[1024] last-layer output or null
};

// A second (or first) layer stack a step can run on: the group kernel executes a whole codec frame -- 16 code-predictor
// steps and the talker step -- in one launch, so a launch carries two models.  models[0] mirrors the fields below.
struct ModelDesc {
  const uint8_t* packed_layers;
  const uint8_t* aux_layers;
  const __nv_bfloat16* cos_t;
  const __nv_bfloat16* sin_t;
  __nv_bfloat16* k_cache;
  __nv_bfloat16* v_cache;
  int max_seq;
  int L;
  int residual_fp32;
  unsigned long long rope_axis[2]; // M-RoPE: 2 bits per rotary frequency index 0..63 = axis (0..2) whose position that frequency
                                   // uses (StepDesc::rope_pos); all zero = standard RoPE
};
// Device-autonomous frame loop (group kernel): the launch repeats its step program `n_frames` times without the host
// (upstream analogue: launch_ldg_generate_nosync, kernel.cu:1555-1613).  Frame f: stop if the talker's last token is EOS;
// code-predictor steps write codes_out[f][0..15]; the talker step (pos_per_frame = 1) runs at position + f on
// embed-sum(codes) + (f0 + f < n_trailing ? trailing[f0 + f] : pad_embed) and writes tokens_out[f].
struct FrameLoop {
  int n_frames;                    // 0 = plain launch (no loop)
  int eos_token;                   // < 0: never stop early
  const __nv_bfloat16* trailing;   // bf16[n_trailing][1024] or null
  int n_trailing, trailing_offset;
  const __nv_bfloat16* pad_embed;  // bf16[1024]
  long long* codes_out;            // int64[n_frames][16]
  int* tokens_out;                 // int32[n_frames] or null: talker token produced by frame f
  int* gen_state;                  // int32[4] {frames done (cumulative over chained launches), eos seen, last token, -}
  int epochs_per_frame;            // sum of L + 2 over the step program
};

struct Params {
  Layout lay;
  const uint8_t* packed_layers;    // [G][L][layer_segs][2048]
  const uint8_t* aux_layers;       // [L][2][AUX_BYTES]
  const __nv_bfloat16* cos_t;      // [max_seq][128]
  const __nv_bfloat16* sin_t;
  __nv_bfloat16* k_cache;          // [L][8][max_seq][128]
  __nv_bfloat16* v_cache;
  int max_seq;
  float attn_scale;
  int residual_fp32;
  uint8_t* xbuf;                   // exchange words (XBUF_BYTES)
  float* res_spill;                // f32[1024]: fp32 residual of every row (read back only by staged launches)
  int* delays;                     // [G][3][DL_N]: poll delays (cycles after own publish; in), repeated-poll counts and
                                   // cycles/16 spent waiting for weights per exchange kind (in/out)
  int o_sentinel;                  // O phase of idle CTAs: one warp polls sample words before the CTA-wide gather
  int delay_o_idle;                // O-phase delay of CTAs without an attention item (they wait for the attention CTAs)
  uint32_t epoch_base;             // epochs base+1 .. base+n_steps*(L+2) are used by this launch (16-bit, never 0)
  int* status;                     // int[4]: code, cta, phase index, aux
  int phase_begin, phase_end;      // half-open range in the linear phase index space of step 0 (staged mode)
  long long timeout_cycles;
  long long* trace;                // optional [G][trace_stride] clock64 stamps
  int trace_stride;
  float sample_temperature;        // code-predictor sampling (upstream model_tts.py:756-762)
  int sample_top_k;
  unsigned long long sample_seed, sample_counter;
  const int* forced_tokens;        // optional int32[15] (device): token fed to the next step instead of the selected one
  const __nv_bfloat16* sum_tables[15];   // IN_CODES_SUM: the 15 code-predictor embedding tables [2048, 1024]
  const int* code0_ptr;            // optional: code0 comes from device memory (the talker's out_token)
  long long* code0_out;            // optional: receives code0 (the talker's token, first entry of the frame's codes)
  int code0;
  ModelDesc models[2];
  FrameLoop frames;
  int sum_rows0, sum_rows;         // IN_CODES_SUM clamp bounds: rows of the talker table / of the 15 group tables
  int n_steps;
  StepDesc steps[MAX_STEPS];
};
static_assert(sizeof(Params) <= 4000, "kernel parameter space");

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int row_begin(int cta, int n_rows, int G) {
  return (int)(((unsigned)cta * (unsigned)n_rows) / (unsigned)G);   // cta <= 1024, n_rows <= 4096: fits 32 bits
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
// Transposed reduction of four per-lane partial sums: afterwards every lane of the 8-lane group q holds
// the warp-wide total of v[q].  6 shuffles instead of 20.
__device__ __forceinline__ float warp_sum4(const float (&v)[4], int lane) {
  const bool hi = (lane & 16) != 0, mid = (lane & 8) != 0;
  const float u0 = (hi ? v[2] : v[0]) + __shfl_xor_sync(0xffffffffu, hi ? v[0] : v[2], 16);
  const float u1 = (hi ? v[3] : v[1]) + __shfl_xor_sync(0xffffffffu, hi ? v[1] : v[3], 16);
  float w = (mid ? u1 : u0) + __shfl_xor_sync(0xffffffffu, mid ? u0 : u1, 8);
  w += __shfl_xor_sync(0xffffffffu, w, 4);
  w += __shfl_xor_sync(0xffffffffu, w, 2);
  w += __shfl_xor_sync(0xffffffffu, w, 1);
  return w;
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
__device__ __forceinline__ bool mbar_test_wait(u64* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Three phase tests issued back to back (their ~150-cycle latencies overlap); bit s of the result = barrier s complete.
__device__ __forceinline__ uint32_t mbar_test_wait3(u64* b0, uint32_t p0, u64* b1, uint32_t p1, u64* b2, uint32_t p2) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred q0, q1, q2;\n\t.reg .u32 r0, r1, r2;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q0, [%1], %2;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q1, [%3], %4;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q2, [%5], %6;\n\t"
      "selp.u32 r0, 1, 0, q0;\n\t"
      "selp.u32 r1, 2, 0, q1;\n\t"
      "selp.u32 r2, 4, 0, q2;\n\t"
      "or.b32 r0, r0, r1;\n\t"
      "or.b32 %0, r0, r2;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b0)), "r"(p0), "r"(smem_u32(b1)), "r"(p1), "r"(smem_u32(b2)), "r"(p2)
      : "memory");
  return ok;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, u64* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// LL4 words: {epoch (high 16), bf16 payload (low 16)} in one 32-bit relaxed gpu-scope access.
__device__ __forceinline__ void ll4_st(uint32_t* p, float value, uint32_t epoch) {   // rounds to bf16 (RNE)
  const uint32_t v = (epoch << 16) | bf16_bits(value);
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ll4_ld4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ bool ll4_ok(const uint4& w, uint32_t epoch) {
  return ((w.x >> 16) == epoch) & ((w.y >> 16) == epoch) & ((w.z >> 16) == epoch) & ((w.w >> 16) == epoch);
}
__device__ __forceinline__ float ll4_val(uint32_t w) { return __uint_as_float(w << 16); }
// LL8 words: {epoch (high 32), payload (low 32)}; used for the fp32 split-attention partials and the token.
__device__ __forceinline__ void ll8_st(u64* p, uint32_t payload, uint32_t epoch) {
  u64 v = ((u64)epoch << 32) | payload;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ll8_ld(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_cg_u2(const void* p) {
  uint2 v;
  asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }
// Barrier over the consumer warps that also ORs a predicate across them.
__device__ __forceinline__ bool consumer_bar_or(bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 p, %1, 0;\n\t"
      "bar.red.or.pred q, 1, %2, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(out)
      : "r"((uint32_t)pred), "n"(NCT)
      : "memory");
  return out != 0;
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up
// ------------------------------------------------------------------------------------------------
constexpr int SM_RING = 0;
constexpr int SM_VEC = SM_RING + NSLOTS * SLOT_BYTES;     // bf16[3072]: activation vector in weight-segment order
constexpr int SM_ACC = SM_VEC + INTER * 2;                // float[NCW][2][128]: attention cross-warp merge
constexpr int SM_SMALL = SM_ACC + NCW * 2 * HD * 4;              // float[1024]  attention scratch
constexpr int SM_PART = SM_SMALL + 1024 * 4;              // float[56][NCW] per-item, per-warp (K slice) partial dot products
constexpr int SM_RED = SM_PART + MAX_ITEMS * NCW * 4;     // float[64]    cross-warp reductions
constexpr int SM_BAR = SM_RED + 64 * 4;                   // u64 full[NSLOTS]
constexpr int SM_TBL = SM_BAR + 2 * 8 * 8;                // uint4[16]: per-layer stage table of the weight stream
constexpr int SM_MISC = SM_TBL + 16 * 16;                 // int abort; int delays[DL_N]; ...
constexpr int SMEM_BYTES = SM_MISC + 256;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

// s_small sub-offsets (floats)
constexpr int SS_V = 0, SS_QN = 128, SS_KN = 384, SS_M = 512, SS_L = 544, SS_CS = 576;  // CS: cos[64], sin[64]

struct Ctx {
  const Params& p;
  uint8_t* ring;
  uint8_t* s_vec;   // bf16[3072]
  float* s_acc;
  float* s_small;
  float* s_part;
  float* s_red;
  u64* full;
  volatile int* s_abort;
  int* s_delay;
  uint4* s_tbl;   // [kind 0..4]: {first stage entry, #stages, #items, -}; [5 + j]: {src offset, item bytes, aux offset + 1, -}
  uint32_t* x32;
  int tid, warp, lane, cta;
  uint32_t k;     // stage counter (same sequence in producer and consumers)
  long long t0;
  long long t_pub;  // clock64 at this CTA's latest publish
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

__device__ __noinline__ void wait_full_slow(int* status, volatile int* s_abort, long long t0, long long timeout, int cta,
                                           int cur_idx, u64* bar, uint32_t parity, int k) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 63u) == 0 && check_abort_slow(status, s_abort, t0, timeout, cta, cur_idx, ST_TIMEOUT_FULL, k)) return;
  }
}
__device__ __forceinline__ void wait_full(Ctx& c, uint32_t k) {
  u64* bar = &c.full[k % NSLOTS];
  const uint32_t parity = (k / NSLOTS) & 1u;
  if (QMK_LIKELY(mbar_try_wait(bar, parity))) return;
  wait_full_slow(c.p.status, c.s_abort, c.t0, c.p.timeout_cycles, c.cta, c.cur_idx, bar, parity, (int)k);
}
__device__ __noinline__ uint4 ll4_wait_slow(int* status, volatile int* s_abort, long long t0, long long timeout, int cta,
                                            int cur_idx, const uint32_t* p, uint32_t epoch, int aux) {
  uint32_t spins = 0;
  for (;;) {
    uint4 w = ll4_ld4(p);
    if (ll4_ok(w, epoch)) return w;
    if ((++spins & 127u) == 0 && check_abort_slow(status, s_abort, t0, timeout, cta, cur_idx, ST_TIMEOUT_LL, aux)) return w;
  }
}
__device__ __forceinline__ uint4 ll4_wait(Ctx& c, const uint32_t* p, uint32_t epoch, bool& retried) {
  uint4 w = ll4_ld4(p);
  if (!ll4_ok(w, epoch)) {
    retried = true;
    w = ll4_wait_slow(c.p.status, c.s_abort, c.t0, c.p.timeout_cycles, c.cta, c.cur_idx, p, epoch, (int)(p - c.x32));
  }
  return w;
}
__device__ __forceinline__ uint32_t ll8_wait(Ctx& c, const u64* p, uint32_t epoch) {
  u64 w = ll8_ld(p);
  uint32_t spins = 0;
  while ((uint32_t)(w >> 32) != epoch) {
    if ((++spins & 127u) == 0 && check_abort(c, ST_TIMEOUT_LL, -1)) break;
    w = ll8_ld(p);
  }
  return (uint32_t)w;
}

// Debug trace: trace[cta][(idx - phase_begin) * 8 + sub] = clock64() (thread 0 of the CTA only).
constexpr int TRACE_SUBS = 24;   // per phase: 8 coarse 64-bit stamps + 16 fine 32-bit stamps of the GEMV path
template <bool TR>
__device__ __forceinline__ void trace_sub(const Ctx& c, int sub) {
  if (TR && c.p.trace != nullptr && c.tid == 0) {
    const int slot = (c.cur_idx - c.p.phase_begin) * TRACE_SUBS + sub;
    // sub-slot 7 (row published) records the GLOBAL timer in ns so that publish times compare across SMs
    long long t;
    if (sub == 7) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); else t = clock64();
    if (slot >= 0 && slot < c.p.trace_stride) c.p.trace[(size_t)c.cta * c.p.trace_stride + slot] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// phase descriptors (identical arithmetic in producer and consumers)
// ------------------------------------------------------------------------------------------------
struct PhaseDesc {
  const uint8_t* src;   // this CTA's items of the phase (contiguous 2 KB segments)
  const uint8_t* aux;   // AUX_BYTES block fetched with the first stage, or null
  int n_items;
};
struct CtaRows {  // per-CTA row ranges, computed once per kernel
  int q_row0, q_rows, o_row0, o_rows, gu_row0, gu_rows;
};
__device__ __forceinline__ CtaRows cta_rows(const Layout& y, int cta) {
  CtaRows r;
  r.q_row0 = row_begin(cta, QKV_ROWS, y.G);
  r.q_rows = row_begin(cta + 1, QKV_ROWS, y.G) - r.q_row0;
  r.o_row0 = row_begin(cta, H, y.G);
  r.o_rows = row_begin(cta + 1, H, y.G) - r.o_row0;
  r.gu_row0 = row_begin(cta, INTER, y.G);
  r.gu_rows = row_begin(cta + 1, INTER, y.G) - r.gu_row0;
  return r;
}
__device__ __forceinline__ PhaseDesc layer_phase_desc(const Params& p, const CtaRows& r, int l, int ph, int cta) {
  const Layout& y = p.lay;
  PhaseDesc d;
  const uint8_t* base = p.packed_layers + ((size_t)((size_t)cta * y.L + l) * y.layer_segs) * SEG_BYTES;
  d.aux = nullptr;
  if (ph == PH_QKV) {
    d.n_items = r.q_rows;
    d.src = base;
    d.aux = p.aux_layers + ((size_t)l * 2 + 0) * AUX_BYTES;
  } else if (ph == PH_O) {
    d.n_items = 2 * r.o_rows;
    d.src = base + (size_t)y.off_o * SEG_BYTES;
  } else if (ph == PH_GU) {
    d.n_items = 2 * r.gu_rows;
    d.src = base + (size_t)y.off_gu * SEG_BYTES;
    d.aux = p.aux_layers + ((size_t)l * 2 + 1) * AUX_BYTES;
  } else {  // PH_DOWN
    d.n_items = 3 * r.o_rows;
    d.src = base + (size_t)y.off_down * SEG_BYTES;
  }
  return d;
}
__device__ __forceinline__ PhaseDesc head_phase_desc(const Params& p, const HeadDesc& h, int cta) {
  PhaseDesc d;
  d.aux = h.aux;
  if (h.rows > 0) {
    d.n_items = row_begin(cta + 1, h.rows, p.lay.G) - row_begin(cta, h.rows, p.lay.G);
    d.src = h.packed + ((size_t)cta * h.segs_max) * SEG_BYTES;
  } else {
    d.n_items = 0;
    d.src = h.aux;
  }
  return d;
}
// A phase with an aux block always has at least one stage.
__device__ __forceinline__ int n_stages_of(const PhaseDesc& d) {
  const int n = (d.n_items + STAGE_ITEMS - 1) / STAGE_ITEMS;
  return (n == 0 && d.aux != nullptr) ? 1 : n;
}

// ------------------------------------------------------------------------------------------------
// producer
// ------------------------------------------------------------------------------------------------
// Weight producer.  There is no producer warp: the ring is refilled by one lane of the last consumer warp right
// after the barrier that ends a phase's tensor-core stages (every warp has finished reading those slots by then),
// so the stage that is NSLOTS ahead is always issued as soon as its slot is free, without empty-barriers and
// without a ninth warp (which would cap the kernel at 168 registers per thread).
struct Prod {
  const uint8_t* layer_base;  // this CTA's packed block of layer `l`
  const uint8_t* aux_base;    // aux blocks of layer `l`
  int step, l, j;     // next stage to issue: entry j of the layer table (l == L: stage j of the head phase)
  int j_end;          // entries per layer for this launch (staged launches cover a sub-range of one layer)
  uint32_t k;         // its global stage number (slot = k % NSLOTS)
  int pending;        // stages whose slots have been released but not refilled yet
  int left;           // stages still to issue in this launch
};
// Stage table of one layer (identical for every layer of this CTA), built once per launch by thread 0.
//   s_tbl[kind 0..4] = {first entry, #stages, #items, -};  s_tbl[5 + j] = {src offset, item bytes, aux offset + 1, -}
__device__ __noinline__ void prod_build_table_slow(uint4* s_tbl, int q_rows, int o_rows, int gu_rows, int off_o, int off_gu,
                                                   int off_down) {
  const int n_items[5] = {q_rows, 0, 2 * o_rows, 2 * gu_rows, 3 * o_rows};
  const int off_seg[5] = {0, 0, off_o, off_gu, off_down};
  int e = 5;
#pragma unroll 1
  for (int kind = 0; kind < 5; ++kind) {
    const int nst = (n_items[kind] + STAGE_ITEMS - 1) / STAGE_ITEMS;
    s_tbl[kind] = make_uint4((uint32_t)e, (uint32_t)nst, (uint32_t)n_items[kind], 0u);
#pragma unroll 1
    for (int s = 0; s < nst; ++s, ++e) {
      int items = n_items[kind] - s * STAGE_ITEMS;
      if (items > STAGE_ITEMS) items = STAGE_ITEMS;
      const uint32_t aux = (s == 0 && (kind == PH_QKV || kind == PH_GU)) ? (uint32_t)((kind == PH_GU) ? AUX_BYTES : 0) + 1u : 0u;
      s_tbl[e] = make_uint4((uint32_t)((off_seg[kind] + s * STAGE_ITEMS) * SEG_BYTES), (uint32_t)(items * SEG_BYTES), aux, 0u);
    }
  }
}
__device__ __forceinline__ void prod_build_table(Ctx& c, const CtaRows& rows) {
  if (c.tid == 0)
    prod_build_table_slow(c.s_tbl, rows.q_rows, rows.o_rows, rows.gu_rows, c.p.lay.off_o, c.p.lay.off_gu, c.p.lay.off_down);
}
__device__ __forceinline__ bool step_has_final(const StepDesc& sd) {
  return sd.head.rows > 0 || sd.out_norm != nullptr || sd.hidden_out != nullptr;
}
__device__ __forceinline__ int head_stages(const Params& p, const HeadDesc& h, int cta) {
  return h.rows > 0 ? (row_begin(cta + 1, h.rows, p.lay.G) - row_begin(cta, h.rows, p.lay.G) + STAGE_ITEMS - 1) / STAGE_ITEMS : 1;
}
// Call after prod_build_table + barrier.
__device__ __forceinline__ void prod_init(Ctx& c, Prod& pr) {
  const Params& p = c.p;
  const int L = p.lay.L, nlayer_idx = L * PH_PER_LAYER;
  const int per_layer = (int)(c.s_tbl[PH_DOWN].x + c.s_tbl[PH_DOWN].y) - 5;
  pr.step = 0; pr.k = 0; pr.pending = 0;
  if (p.n_steps > 1 || (p.phase_begin == 0 && p.phase_end == nlayer_idx + 2)) {   // whole steps
    pr.l = 0; pr.j = 0; pr.j_end = per_layer;
    pr.left = 0;
    for (int st = 0; st < p.n_steps; ++st) pr.left += L * per_layer + head_stages(p, p.steps[st].head, c.cta);
  } else {   // staged launch: exactly one phase
    const int idx = p.phase_begin;
    if (idx < nlayer_idx) {
      const uint4 t = c.s_tbl[idx % PH_PER_LAYER];
      pr.l = idx / PH_PER_LAYER; pr.j = (int)t.x - 5; pr.j_end = pr.j + (int)t.y; pr.left = (int)t.y;
    } else {
      pr.l = L; pr.j = 0; pr.j_end = 0;
      pr.left = (idx == nlayer_idx) ? head_stages(p, p.steps[0].head, c.cta) : 0;
    }
  }
  pr.layer_base = p.packed_layers + ((size_t)((size_t)c.cta * L + pr.l) * p.lay.layer_segs) * SEG_BYTES;
  pr.aux_base = p.aux_layers + (size_t)pr.l * 2 * AUX_BYTES;
}
__device__ __forceinline__ void prod_tma(Ctx& c, uint32_t k, const uint8_t* src, uint32_t item_bytes, const uint8_t* aux) {
  // No proxy fence: every generic-proxy read of the slot (ldmatrix / aux loads) has completed before the barrier
  // that precedes the refill, exactly like the consumer-release of a TMA pipeline stage.
  const int slot = k % NSLOTS;
  uint8_t* dst = c.ring + (size_t)slot * SLOT_BYTES;
  mbar_arrive_expect_tx(&c.full[slot], item_bytes + (aux ? AUX_BYTES : 0));
  if (aux) tma_bulk_g2s(dst, aux, AUX_BYTES, &c.full[slot]);
  if (item_bytes) tma_bulk_g2s(dst + AUX_BYTES, src, item_bytes, &c.full[slot]);
}
// Issue the next `n` stages of the weight stream (lane 0 of the calling warp; all lanes keep the cursor).
__device__ __forceinline__ void prod_issue(Ctx& c, Prod& pr, int n) {
  const Params& p = c.p;
  if (n > pr.left) n = pr.left;
  pr.left -= n;
  for (; n > 0; --n, ++pr.k) {
    if (QMK_LIKELY(pr.l < p.lay.L)) {
      if (c.lane == 0) {
        const uint4 e = c.s_tbl[5 + pr.j];
        prod_tma(c, pr.k, pr.layer_base + e.x, e.y, e.z ? pr.aux_base + (e.z - 1u) : nullptr);
      }
      if (++pr.j == pr.j_end) {
        pr.j = 0;
        ++pr.l;
        pr.layer_base += (size_t)p.lay.layer_segs * SEG_BYTES;
        pr.aux_base += 2 * AUX_BYTES;
      }
    } else {   // head phase of step pr.step
      const PhaseDesc d = head_phase_desc(p, p.steps[pr.step].head, c.cta);
      int items = d.n_items - pr.j * STAGE_ITEMS;
      if (items > STAGE_ITEMS) items = STAGE_ITEMS;
      if (items < 0) items = 0;
      if (c.lane == 0)
        prod_tma(c, pr.k, d.src + (size_t)pr.j * STAGE_ITEMS * SEG_BYTES, (uint32_t)items * SEG_BYTES, pr.j == 0 ? d.aux : nullptr);
      if (++pr.j >= n_stages_of(d)) {   // next step starts again at layer 0
        pr.j = 0;
        pr.l = 0;
        ++pr.step;
        pr.layer_base = p.packed_layers + ((size_t)((size_t)c.cta * p.lay.L) * p.lay.layer_segs) * SEG_BYTES;
        pr.aux_base = p.aux_layers;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// consumer building blocks
// ------------------------------------------------------------------------------------------------
// GEMV on the legacy tensor path (mma.sync m16n8k16, bf16 x bf16 -> fp32).  Measured on B200
// (scripts/ubench/mma_rate.cu): the CUDA-core version spends one conversion + one FMA issue slot per weight and
// is issue-bound at ~4x the shared-memory time of a stage; HMMA sustains one 16x16 weight tile per ~2 cycles per
// SM, so the stage becomes shared-memory-bandwidth bound.  A stage is a 14-row x 1024-k tile; the tile's rows
// are the A operand (16-byte chunk c of row r is stored at chunk c ^ (r & 7): conflict-free ldmatrix), the
// B operand's eight columns are the activation segments (column n = x[n*1024 ...]), so a row whose k-block is
// kb reads its result from column kb.  Each consumer warp owns 4 of the 64 k16-steps.
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}

// The wait window before a gather: warp 0 (which issued the publish) spins until t_pub + delay while the other
// consumer warps sleep at a hardware barrier -- spinning warps would steal issue slots from the warp that is
// still finalising the previous phase.  Returns a bit mask of the next phase's ring stages already full.
// `sentinel` (optional): after the delay warp 0 alone polls 32 sample words of the vector (one per 64) until they
// carry `epoch`, so that a long wait (the O phase waiting for the attention CTAs) does not have every thread of 140
// CTAs re-reading the lines the producers are about to write.
__device__ __forceinline__ void wait_window(Ctx& c, int delay, const uint32_t* sentinel = nullptr, uint32_t epoch = 0) {
  if (c.warp == 0) {
    const long long t_ready = c.t_pub + delay;
    while (clock64() < t_ready) {
    }
    if (sentinel != nullptr) {
      uint32_t spins = 0;
      for (;;) {
        uint32_t w;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w) : "l"(sentinel + c.lane * 64) : "memory");
        if (__all_sync(0xffffffffu, (w >> 16) == epoch)) break;
        if ((++spins & 63u) == 0 && check_abort(c, ST_TIMEOUT_LL, -2)) break;
      }
    }
  }
  asm volatile("bar.sync 2, %0;" ::"n"(NCT) : "memory");
}
// After a gather: consumer barrier that also counts the gathers in which some poll had to be repeated
// (polling a line before it is written is what makes an exchange slow; the counts guide the delay tuning).
__device__ __forceinline__ void gather_bar(Ctx& c, int dslot, bool retried) {
  const bool any = consumer_bar_or(retried);
  if (any && c.tid == 0) c.s_delay[DL_N + dslot] += 1;
}
// barrier-free variant for gathers whose consumers are the gathering warp itself
__device__ __forceinline__ void gather_note(Ctx& c, int dslot, bool retried) {
  if (QMK_UNLIKELY(__any_sync(0xffffffffu, retried)) && c.lane == 0) atomicAdd(&c.s_delay[DL_N + dslot], 1);
}

// ------------------------------------------------------------------------------------------------
// attention work items: (kv head g, split s).  Item (g, s) runs on CTA s*8 + g.
// ------------------------------------------------------------------------------------------------
struct AttnItem {
  int g, s, S, p0, p1;
  bool owner;  // holds the new position (= position) -> takes k,v from the exchange and appends to the cache
};
__device__ __forceinline__ bool attn_item(const Params& p, int position, int cta, AttnItem& it) {
  const int n = position + 1;
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
__device__ __forceinline__ void attn_prefetch(const Ctx& c, int l, int position, const AttnItem& it, int round, KvRegs& r) {
  const Params& p = c.p;
  const size_t base = ((size_t)(l * NKVH + it.g) * p.max_seq) * HD;
#pragma unroll
  for (int i = 0; i < ATT_PER_WARP; ++i) {
    const int pos = it.p0 + round * ATT_ROUND + c.warp + NCW * i;
    if (pos < it.p1 && pos != position) {
      const size_t off = base + (size_t)pos * HD + c.lane * 4;
      r.k[i] = ld_cg_u2(p.k_cache + off);
      r.v[i] = ld_cg_u2(p.v_cache + off);
    }
  }
}

// Transposed reduction of 2*ATT_PER_WARP (= 10) per-lane partial dot products; every lane ends up with all
// ten totals (12 reduction shuffles + 10 broadcasts instead of 50 shuffles).
__device__ __forceinline__ void warp_sum10_bcast(float (&v)[10], int lane) {
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0;
  float u[6];
#pragma unroll
  for (int k = 0; k < 5; ++k)
    u[k] = (b4 ? v[k + 5] : v[k]) + __shfl_xor_sync(0xffffffffu, b4 ? v[k] : v[k + 5], 16);
  u[5] = 0.f;
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
  // value index held by a lane: b4*5 + b3*3 + b2*2 + b1  (valid combinations only)
#pragma unroll
  for (int idx = 0; idx < 10; ++idx) {
    const int r = idx % 5, hi = idx / 5;
    const int src = hi * 16 + (r >= 3 ? 8 : 0) + ((r % 3) >= 2 ? 4 : 0) + (((r % 3) % 2) ? 2 : 0);
    v[idx] = __shfl_sync(0xffffffffu, y, src);
  }
}

struct AttnPre {  // per-layer constants a norm/rope warp needs, fetched while the QKV weights are resident
  float nw[4];
};

template <bool TR>
__device__ void phase_attn(Ctx& c, int l, int position, uint32_t epoch, const AttnItem& it, KvRegs& kv,
                           const AttnPre& pre) {
  const Params& p = c.p;
  const uint32_t* x_qkv = c.x32 + XW_QKV;
  uint32_t* x_a = c.x32 + XW_A;
  u64* x_part = reinterpret_cast<u64*>(p.xbuf + (size_t)XW_END32 * 4);
  float* s_small = c.s_small;

  // 1) gather q (2 heads), k, v of this kv group; per-head RMSNorm + rotate-half RoPE in bf16 steps.
  //    warp 0,1: q heads; warp 2: k; warp 3: v (owner only).  Lane owns dims 4*lane .. 4*lane+3.
  bool retried = false;
  wait_window(c, c.s_delay[DL_ATTN]);
  if (c.warp < 2 || (it.owner && c.warp < 4)) {
    const int row0 = (c.warp < 2) ? (2 * it.g + c.warp) * HD : (c.warp == 2 ? QSZ + it.g * HD : QSZ + KVSZ + it.g * HD);
    const uint4 w = ll4_wait(c, x_qkv + row0 + c.lane * 4, epoch, retried);
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
      float o4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float n = bf16_round((t[e] / rms) * pre.nw[e]);
        const float o = __shfl_xor_sync(0xffffffffu, n, 16);
        const float cs = s_small[SS_CS + dbase + e], sn = s_small[SS_CS + 64 + dbase + e];
        const float a = bf16_round(n * cs), b = bf16_round(o * sn);
        o4[e] = bf16_round(c.lane < 16 ? a - b : a + b);
      }
      float* dst = (c.warp < 2) ? s_small + SS_QN + c.warp * HD : s_small + SS_KN;
      *reinterpret_cast<float4*>(dst + c.lane * 4) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
  }
  gather_bar(c, DL_ATTN, retried);   // barrier: q / k / v are in shared memory
  trace_sub<TR>(c, 1);

  // 2) scores / online softmax / PV over this item's positions
  float q0[4], q1[4], acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  {
    const float4 a = *reinterpret_cast<const float4*>(s_small + SS_QN + c.lane * 4);
    const float4 b = *reinterpret_cast<const float4*>(s_small + SS_QN + HD + c.lane * 4);
    q0[0] = a.x; q0[1] = a.y; q0[2] = a.z; q0[3] = a.w;
    q1[0] = b.x; q1[1] = b.y; q1[2] = b.z; q1[3] = b.w;
  }
  const int len = it.p1 - it.p0;
  const int nrounds = (len + ATT_ROUND - 1) / ATT_ROUND;
  for (int r = 0; r < nrounds; ++r) {
    if (r > 0) attn_prefetch(c, l, position, it, r, kv);
    float sc[10];
    const int pos_first = it.p0 + r * ATT_ROUND + c.warp;
    if (pos_first < it.p1) {  // warp-uniform
      // positions of this warp in this round (warp-uniform): a short context (every code-predictor step, the first
      // talker steps) has 1-2, so the loops below stop early instead of predicating five iterations off
      const int nv = min(ATT_PER_WARP, (it.p1 - pos_first + NCW - 1) / NCW);
#pragma unroll
      for (int i = 0; i < 10; ++i) sc[i] = 0.f;
#pragma unroll
      for (int i = 0; i < ATT_PER_WARP; ++i) {
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
        sc[5 + i] = d1;
      }
      if (nv <= 2) {
        const float v4[4] = {sc[0], sc[1], sc[5], sc[6]};
        const float tot = warp_sum4(v4, c.lane);      // 8-lane group q holds the total of v4[q]
        sc[0] = __shfl_sync(0xffffffffu, tot, 0);
        sc[1] = __shfl_sync(0xffffffffu, tot, 8);
        sc[5] = __shfl_sync(0xffffffffu, tot, 16);
        sc[6] = __shfl_sync(0xffffffffu, tot, 24);
      } else {
        warp_sum10_bcast(sc, c.lane);
      }
      float mx0 = m0, mx1 = m1;
#pragma unroll
      for (int i = 0; i < ATT_PER_WARP; ++i) {
        if (i >= nv) break;
        sc[i] *= p.attn_scale;
        sc[5 + i] *= p.attn_scale;
        mx0 = fmaxf(mx0, sc[i]);
        mx1 = fmaxf(mx1, sc[5 + i]);
      }
      const float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - mx0);
      const float c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - mx1);
      l0 *= c0; l1 *= c1;
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc0[e] *= c0; acc1[e] *= c1; }
#pragma unroll
      for (int i = 0; i < ATT_PER_WARP; ++i) {
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
        const float e0 = __expf(sc[i] - mx0), e1 = __expf(sc[5 + i] - mx1);
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
  // 3) cross-warp merge: common max first, then plain sums (s_vec is free during attention)
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
  trace_sub<TR>(c, 3);
  if (c.tid < 256) {
    const int h = c.tid >> 7, d = c.tid & 127;
    float A = 0.f, Lsum = 0.f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) {
      A += c.s_acc[(w * 2 + h) * HD + d];
      Lsum += s_small[SS_L + w * 2 + h];
    }
    const float Mh = h ? M1 : M0;
    const int hq = 2 * it.g + h;
    if (QMK_LIKELY(it.S == 1)) {
      ll4_st(x_a + hq * HD + d, bf16_round(A / Lsum), epoch);
    } else {
      u64* part = x_part + ((size_t)hq * S_MAX + it.s) * PART_STRIDE;
      ll8_st(part + 2 + d, __float_as_uint(A), epoch);
      if (d == 0) {
        ll8_st(part + 0, __float_as_uint(Mh), epoch);
        ll8_st(part + 1, __float_as_uint(Lsum), epoch);
      }
      if (it.s == 0) {  // this CTA merges the splits of its two q heads, fixed order s = 0..S-1
        float Mx = -INFINITY;
        for (int s = 0; s < it.S; ++s) {
          const u64* ps = x_part + ((size_t)hq * S_MAX + s) * PART_STRIDE;
          Mx = fmaxf(Mx, __uint_as_float(ll8_wait(c, ps, epoch)));
        }
        float Lt = 0.f, B = 0.f;
        for (int s = 0; s < it.S; ++s) {
          const u64* ps = x_part + ((size_t)hq * S_MAX + s) * PART_STRIDE;
          const float f = __expf(__uint_as_float(ll8_wait(c, ps, epoch)) - Mx);
          Lt = fmaf(__uint_as_float(ll8_wait(c, ps + 1, epoch)), f, Lt);
          B = fmaf(__uint_as_float(ll8_wait(c, ps + 2 + d, epoch)), f, B);
        }
        ll4_st(x_a + hq * HD + d, bf16_round(B / Lt), epoch);
      }
    }
  }
  c.t_pub = clock64();
  trace_sub<TR>(c, 4);
  // 4) append the new K/V row (off the critical path: after `a` has been published)
  if (it.owner && c.tid < 128) {
    const size_t off = ((size_t)(l * NKVH + it.g) * p.max_seq + position) * HD + c.tid;
    p.k_cache[off] = __float2bfloat16_rn(s_small[SS_KN + c.tid]);
    p.v_cache[off] = __float2bfloat16_rn(s_small[SS_V + c.tid]);
  }
  consumer_bar();  // s_small / s_vec are reused by the next phase
}

// ------------------------------------------------------------------------------------------------
// consumer main loop.  All five GEMV-shaped phases (QKV, O, gate/up, down, LM head) run through ONE copy of
// the gather / norm / stage / reduce code: the kernel's instruction footprint has to stay inside the SM's
// instruction cache (a per-phase inlined version measured 2-3x slower on every sub-step).
// ------------------------------------------------------------------------------------------------
// Temperature / top-k (ties kept) / multinomial selection over `hrows` bf16 logits held as floats in s_log; every
// consumer thread of the CTA calls it (it contains consumer barriers) and every thread returns the same token.
// Out of line: it runs on one CTA once per step and would otherwise sit in the middle of the hot loop's code.
__device__ __noinline__ int sample_token(const float* s_log, unsigned* hist, float* s_red, int tid, int warp, int lane,
                                         int hrows, int top_k, float temperature, unsigned long long seed,
                                         unsigned long long counter, int group, float best, int best_i) {
  // (1) k-th largest logit by a two-pass radix select on the order-preserving 16-bit key of the bf16 value
  unsigned* s_sel = reinterpret_cast<unsigned*>(s_red) + 40;  // [0] bin, [1] remaining rank, [2] chosen
  const int per = hrows / NCT;                               // contiguous elements per thread (CDF in index order)
  const int i0 = tid * per;
  int k_rank = top_k;
  unsigned thr_key = 0;
  if (k_rank > 0 && k_rank < hrows) {
    unsigned prefix_hi = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      hist[tid] = 0;   // NCT == 256 bins
      consumer_bar();
#pragma unroll 1
      for (int e = 0; e < per; ++e) {
        const unsigned bits = __float_as_uint(s_log[i0 + e]) >> 16;
        const unsigned key = (bits & 0x8000u) ? (~bits & 0xffffu) : (bits | 0x8000u);
        if (pass == 0) atomicAdd(&hist[key >> 8], 1u);
        else if ((key >> 8) == prefix_hi) atomicAdd(&hist[key & 0xffu], 1u);
      }
      consumer_bar();
      if (warp == 0) {   // lane l owns bins 255-8l .. 248-8l; suffix counts from the top
        unsigned cnt[8], tot = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) { cnt[e] = hist[255 - 8 * lane - e]; tot += cnt[e]; }
        unsigned incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const unsigned before = incl - tot;   // elements in higher bins
        if (before < (unsigned)k_rank && incl >= (unsigned)k_rank) {
          unsigned run = before;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (run < (unsigned)k_rank && run + cnt[e] >= (unsigned)k_rank) { s_sel[0] = 255 - 8 * lane - e; s_sel[1] = (unsigned)k_rank - run; }
            run += cnt[e];
          }
        }
      }
      consumer_bar();
      if (pass == 0) { prefix_hi = s_sel[0]; k_rank = (int)s_sel[1]; }
      else thr_key = (prefix_hi << 8) | s_sel[0];
      consumer_bar();
    }
  }
  // (2) softmax over the kept logits / temperature, inverse-CDF draw in index order
  const float inv_t = 1.0f / temperature;
  const float zmax = best * inv_t;
  float pe[8];
  float local = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    pe[e] = 0.f;
    if (e < per) {
      const float v = s_log[i0 + e];
      const unsigned bits = __float_as_uint(v) >> 16;
      const unsigned key = (bits & 0x8000u) ? (~bits & 0xffffu) : (bits | 0x8000u);
      if (key >= thr_key) pe[e] = __expf(v * inv_t - zmax);
      local += pe[e];
    }
  }
  float incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_red[16 + warp] = incl;
  if (tid == 0) s_sel[2] = (unsigned)best_i;   // fallback if rounding leaves the target beyond the last element
  consumer_bar();
  float wbase = 0.f, total = 0.f;
#pragma unroll
  for (int w = 0; w < NCW; ++w) {
    const float t = s_red[16 + w];
    if (w < warp) wbase += t;
    total += t;
  }
  // counter-based uniform in [0, 1): splitmix64 of (seed, frame counter, group)
  unsigned long long zr = seed + 0x9E3779B97F4A7C15ull * (counter * 16ull + (unsigned long long)(group + 1));
  zr = (zr ^ (zr >> 30)) * 0xBF58476D1CE4E5B9ull;
  zr = (zr ^ (zr >> 27)) * 0x94D049BB133111EBull;
  zr ^= zr >> 31;
  const float target = (float)(zr >> 40) * (1.0f / 16777216.0f) * total;
  const float lo = wbase + incl - local;
  if (target >= lo && target < lo + local) {
    float run = lo;
    int pick = -1;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (e < per && pe[e] > 0.f) {
        if (pick < 0 && target < run + pe[e]) pick = i0 + e;
        run += pe[e];
      }
    }
    if (pick >= 0) s_sel[2] = (unsigned)pick;
  }
  consumer_bar();
  return (int)s_sel[2];
}

// sum of the NCW per-warp K-slice partials of one item, fixed order
__device__ __forceinline__ float item_sum(const float* s_part, int it) {
  static_assert(NCW == 8 && KSTEPS == 8, "item_sum reads 8 partials; pack_chunk assumes 8 K-slices of 128");
  const float4* v = reinterpret_cast<const float4*>(s_part + it * NCW);
  const float4 a = v[0], b = v[1];
  return ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
}

enum Kind { K_QKV = PH_QKV, K_ATTN = PH_ATTN, K_O = PH_O, K_GU = PH_GU, K_DOWN = PH_DOWN, K_HEAD = 5, K_ARGMAX = 6 };

template <bool TR>
__device__ void consumer_loop(Ctx& c) {
  const Params& p = c.p;
  const Layout& y = p.lay;
  const int nlayer_idx = y.L * PH_PER_LAYER;
  uint32_t* const x_res = c.x32 + XW_RES;
  uint32_t* const x_qkv = c.x32 + XW_QKV;
  uint32_t* const x_a = c.x32 + XW_A;
  uint32_t* const x_res2 = c.x32 + XW_RES2;
  uint32_t* const x_m = c.x32 + XW_M;
  uint32_t* const x_logits = c.x32 + XW_LOGITS;
  const CtaRows rows = cta_rows(y, c.cta);
  Prod prod;
  prod_build_table(c, rows);
  consumer_bar();
  prod_init(c, prod);
  if (c.warp == NCW - 1) prod_issue(c, prod, NSLOTS);   // fill the ring
  KvRegs kv;
  AttnPre pre = {{0.f, 0.f, 0.f, 0.f}};
  float res_mine = 0.f;  // fp32 residual of row o_row0 + tid (tid < o_rows)
  // ldmatrix row address of this lane inside a stage tile; rows 14, 15 of the m16 tile alias rows 6, 7
  // (their results are unused and the alias keeps the 8 rows of a matrix on distinct bank groups)
  const int a_mi = c.lane >> 3;
  int a_row = (c.lane & 7) + (a_mi & 1) * 8;
  if (a_row >= STAGE_ITEMS) a_row -= 8;
  const uint32_t a_off = (uint32_t)(AUX_BYTES + a_row * SEG_BYTES);
  const int a_sw = a_row & 7, a_khalf = a_mi >> 1;
  // down-projection: k-block of the item in tile row r of stage s is (14 s + r) % 3; 2 bits per (s, half)
  uint32_t kb3_pack = 0;
#pragma unroll
  for (int sh = 0; sh < 2 * MAX_ST; ++sh)
    kb3_pack |= (uint32_t)(((sh >> 1) * STAGE_ITEMS + (c.lane >> 2) + (sh & 1) * 8) % 3) << (2 * sh);

  if (c.cta == 0 && c.tid == 0 && p.code0_out != nullptr) *p.code0_out = (long long)(p.code0_ptr ? *p.code0_ptr : p.code0);
  for (int step = 0; step < p.n_steps; ++step) {
    const StepDesc& sd = p.steps[step];
    const uint32_t ebase = p.epoch_base + (uint32_t)step * (uint32_t)(y.L + 2);
    const int position = sd.position;
    int in_token = sd.token;
    if (sd.in_mode == IN_TABLE_TOKEN && sd.token_ptr != nullptr) {   // written by an earlier launch on the same stream
      in_token = *sd.token_ptr;
      in_token = in_token < 0 ? 0 : (in_token > sd.token ? sd.token : in_token);   // sd.token = last valid row
    }
    if (sd.in_mode == IN_TABLE_PREV) {
      // the token chosen by CTA 0 at the end of the previous step (one LL8 word, epoch = that step's head epoch)
      if (c.warp == 0) {
        const long long t_ready = c.t_pub + c.s_delay[DL_TOKEN];
        while (clock64() < t_ready) {
        }
        if (c.lane == 0) {
          const u64* x_tok = reinterpret_cast<const u64*>(c.x32 + XW_TOKEN);
          c.s_red[32] = __int_as_float((int)ll8_wait(c, x_tok + (step & 15), (ebase - 1u) & 0xffffu));
        }
      }
      consumer_bar();
      in_token = __float_as_int(c.s_red[32]);
    }
    const __nv_bfloat16* x_in = (sd.in_mode == IN_TABLE_TOKEN || sd.in_mode == IN_TABLE_PREV)
                                    ? sd.in_table + (size_t)in_token * H
                                    : reinterpret_cast<const __nv_bfloat16*>(sd.in_vec);
    AttnItem item;
    const bool has_item = attn_item(p, position, c.cta, item);
    const int head_row0 = sd.head.rows > 0 ? row_begin(c.cta, sd.head.rows, y.G) : 0;
    if (has_item && c.tid < 128) {  // RoPE row of this step: cos[0..63], sin[0..63]
      const int d = c.tid & 63;
      const __nv_bfloat16* t = (c.tid < 64) ? p.cos_t : p.sin_t;
      c.s_small[SS_CS + c.tid] = __bfloat162float(t[(size_t)position * HD + d]);
    }
    consumer_bar();
    int pre_layer = -1;  // layer whose KV rows / qk-norm weights were fetched during its QKV phase
    const int begin = (step == 0) ? p.phase_begin : 0;
    const int end = (p.n_steps == 1) ? p.phase_end : nlayer_idx + 2;
    if (begin > 0 && c.tid < rows.o_rows) res_mine = p.res_spill[rows.o_row0 + c.tid];  // staged launches only

    for (int idx = begin; idx < end; ++idx) {
      c.cur_idx = idx;
      trace_sub<TR>(c, 0);
      const int l = idx / PH_PER_LAYER;
      const int kind = idx < nlayer_idx ? idx % PH_PER_LAYER : (idx == nlayer_idx ? K_HEAD : K_ARGMAX);
      const uint32_t epoch = (ebase + 1u + (uint32_t)l) & 0xffffu;  // layer l; for K_HEAD (l == L) the logits epoch

      if (kind == K_ATTN) {
        if (has_item) {
          if (QMK_UNLIKELY(pre_layer != l)) {  // staged launch: the QKV phase ran in an earlier launch
            attn_prefetch(c, l, position, item, 0, kv);
            if (c.warp < 3) {
              const uint2 wv = *reinterpret_cast<const uint2*>(p.aux_layers + ((size_t)l * 2) * AUX_BYTES + 2048 +
                                                                (c.warp == 2 ? 256 : 0) + c.lane * 8);
              pre.nw[0] = bf16_lo(wv.x); pre.nw[1] = bf16_hi(wv.x); pre.nw[2] = bf16_lo(wv.y); pre.nw[3] = bf16_hi(wv.y);
            }
          }
          phase_attn<TR>(c, l, position, epoch, item, kv, pre);
        }
        continue;
      }
      if (QMK_UNLIKELY(kind == K_ARGMAX)) {
        // token selection on CTA 0: argmax over the bf16 logits (lowest index wins ties), or temperature / top-k /
        // multinomial sampling (upstream model_tts.py:756-762: ties with the k-th value are kept)
        if (c.cta != 0 || sd.head.rows <= 0) continue;
        const int hrows = sd.head.rows;
        const uint32_t epoch_head = (ebase + (uint32_t)y.L + 1u) & 0xffffu;
        const bool sample = sd.select != 0 && hrows <= NCW * 2 * HD && (hrows % (NCT * 4)) == 0;
        float* s_log = c.s_acc;   // up to 2048 logits
        float best = -INFINITY;
        int best_i = 0x7fffffff;
        bool retried = false;
        wait_window(c, c.s_delay[DL_ARGMAX]);
        for (int i = c.tid * 4; i < hrows; i += NCT * 4) {   // indices ascend per thread
          const uint4 w = ll4_wait(c, x_logits + i, epoch_head, retried);
          const float4 v = make_float4(ll4_val(w.x), ll4_val(w.y), ll4_val(w.z), ll4_val(w.w));
          if (sd.logits_out != nullptr) *reinterpret_cast<float4*>(sd.logits_out + i) = v;
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
        for (int w = 0; w < NCW; ++w) {   // every thread: same result
          const float ov = c.s_red[w * 2];
          const int oi = __float_as_int(c.s_red[w * 2 + 1]);
          if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        int chosen = best_i;
        if (sample)
          chosen = sample_token(s_log, reinterpret_cast<unsigned*>(c.s_part), c.s_red, c.tid, c.warp, c.lane, hrows,
                                p.sample_top_k, p.sample_temperature, p.sample_seed, p.sample_counter, sd.group, best, best_i);
        if (c.tid == 0) {
          // a failed launch must not look like a token: encode the watchdog code as a negative id
          const int st = *((volatile int*)p.status);
          const int out = (st != 0 || *c.s_abort) ? -1000 - st : chosen;
          if (sd.out_token != nullptr) *sd.out_token = out;
          if (sd.out_code != nullptr) *sd.out_code = (long long)out;
          if (step + 1 < p.n_steps) {   // hand the (possibly teacher-forced) token to every CTA's next step
            int fed = chosen;
            if (p.forced_tokens != nullptr && sd.group >= 0) fed = p.forced_tokens[sd.group];
            ll8_st(reinterpret_cast<u64*>(c.x32 + XW_TOKEN) + ((step + 1) & 15), (uint32_t)fed, epoch_head);
          }
        }
        consumer_bar();
        continue;
      }

      // ---- GEMV-shaped phases -------------------------------------------------------------------------
      // (traced build only) fine-grained 32-bit clock stamps of thread 0, flushed to trace slots 8..23 of the phase
      uint32_t fs[16];
#define QMK_FS(i) do { if (TR) fs[i] = (uint32_t)clock(); } while (0)
      if (TR) {
#pragma unroll
        for (int i = 0; i < 16; ++i) fs[i] = 0;
      }
      QMK_FS(0);
      // Everything that does not depend on the gathered activations is computed BEFORE they are checked: a single
      // warp executes dependent instructions at ~5 cycles each, so every instruction between "data arrived" and
      // "row published" is on the critical path of the layer.
      const uint32_t* xw = x_res;
      uint32_t* xout = x_qkv + rows.q_row0;     // where thread tid publishes (finalize role depends on the kind)
      int n_words = H, dslot = DL_QKV, xr_mod = 1;
      uint32_t ep = epoch;
      bool norm = true;
      if (kind == K_QKV) { ep = (ebase + (uint32_t)l) & 0xffffu; }
      else if (kind == K_O) { xw = x_a; n_words = QSZ; dslot = DL_O; xr_mod = 2; norm = false; xout = x_res2 + rows.o_row0; }
      else if (kind == K_GU) { xw = x_res2; dslot = DL_GU; xout = x_m + rows.gu_row0; }
      else if (kind == K_DOWN) { xw = x_m; n_words = INTER; dslot = DL_DOWN; xr_mod = 3; norm = false; xout = x_res + rows.o_row0; }
      else { ep = (ebase + (uint32_t)y.L) & 0xffffu; dslot = DL_HEAD; xout = x_logits + head_row0; }
      const bool from_input = (kind == K_QKV && l == 0);

      // gather: thread tid owns elements [4 tid, 4 tid + 4) of each 1024-wide segment, i.e. warp w gathers exactly
      // the K slice [128 w, 128 w + 128) that its own tensor-core steps consume -> no CTA barrier is needed
      // between the gather and the B-fragment loads (only the RMS statistic crosses warps).
      uint4 gw[3];
      const int gi0 = c.tid * 4;
      if (from_input) {
        if (sd.in_mode == IN_CODES_SUM) {
          // the frame loop's 16-way embedding sum + trailing-text embedding (upstream tts_engine.py:319-333), bf16 adds
          // in the upstream order, fused into the step that consumes it
          const uint2 v0 = *reinterpret_cast<const uint2*>(sd.in_table + (size_t)sd.codes[0] * H + gi0);
          float e4[4] = {bf16_lo(v0.x), bf16_hi(v0.x), bf16_lo(v0.y), bf16_hi(v0.y)};
#pragma unroll 1
          for (int g = 0; g < 15; ++g) {
            const uint2 v = *reinterpret_cast<const uint2*>(p.sum_tables[g] + (size_t)sd.codes[g + 1] * H + gi0);
            e4[0] = bf16_round(e4[0] + bf16_lo(v.x)); e4[1] = bf16_round(e4[1] + bf16_hi(v.x));
            e4[2] = bf16_round(e4[2] + bf16_lo(v.y)); e4[3] = bf16_round(e4[3] + bf16_hi(v.y));
          }
          const uint2 vx = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(sd.in_vec) + gi0);
          gw[0] = make_uint4(bf16_bits(e4[0] + bf16_lo(vx.x)), bf16_bits(e4[1] + bf16_hi(vx.x)),
                             bf16_bits(e4[2] + bf16_lo(vx.y)), bf16_bits(e4[3] + bf16_hi(vx.y)));
          *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(c.s_acc) + gi0 * 2) = make_uint2(gw[0].x | (gw[0].y << 16), gw[0].z | (gw[0].w << 16));
        } else if (sd.in_mode == IN_VEC_F32) {
          const float* xf = reinterpret_cast<const float*>(sd.in_vec);
          const float4 v = *reinterpret_cast<const float4*>(xf + gi0);
          gw[0] = make_uint4(bf16_bits(v.x), bf16_bits(v.y), bf16_bits(v.z), bf16_bits(v.w));
          if (c.tid < rows.o_rows) res_mine = bf16_round(xf[rows.o_row0 + c.tid]);
        } else {
          const uint2 v = *reinterpret_cast<const uint2*>(x_in + gi0);
          gw[0] = make_uint4(v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16);
          if (c.tid < rows.o_rows) res_mine = __bfloat162float(x_in[rows.o_row0 + c.tid]);
        }
        if (sd.in_mode != IN_CODES_SUM && c.tid < rows.o_rows) p.res_spill[rows.o_row0 + c.tid] = res_mine;
      } else {
        if (kind == K_O && !has_item) wait_window(c, p.delay_o_idle, p.o_sentinel ? xw : nullptr, ep);
        else wait_window(c, c.s_delay[dslot]);
        trace_sub<TR>(c, 1);
        QMK_FS(1);
        gw[0] = ll4_ld4(xw + gi0);
        if (n_words > H) gw[1] = ll4_ld4(xw + gi0 + H);
        if (n_words > 2 * H) gw[2] = ll4_ld4(xw + gi0 + 2 * H);
      }
      QMK_FS(2);
      // -- shadow of the load latency --
      PhaseDesc d;
      if (kind == K_HEAD) d = head_phase_desc(p, sd.head, c.cta);
      else d = layer_phase_desc(p, rows, l, kind, c.cta);
      const int nst = n_stages_of(d);
      uint32_t ready;
      uint32_t sbase[MAX_ST];
      {
        static_assert(MAX_ST == 3, "mbar_test_wait3");
        u64* fb[MAX_ST];
        uint32_t fp[MAX_ST];
#pragma unroll
        for (int s = 0; s < MAX_ST; ++s) {
          const uint32_t k = c.k + (s < nst ? s : 0);
          const int slot = k % NSLOTS;
          sbase[s] = smem_u32(c.ring + (size_t)slot * SLOT_BYTES);
          fb[s] = &c.full[slot];
          fp[s] = (k / NSLOTS) & 1u;
        }
        ready = mbar_test_wait3(fb[0], fp[0], fb[1], fp[1], fb[2], fp[2]) & ((1u << nst) - 1u);
      }
      // A fragments of the first two stages move from the ring into registers while the activations are in flight:
      // the shared-memory reads (~280 cycles per 28 KB stage, the floor of the stage loop) leave the critical path
      constexpr int PRE_ST = 2;
      uint32_t apre[PRE_ST][KSTEPS][4];
      const uint32_t preloaded = ready;   // stages whose tile was resident at this point
#pragma unroll
      for (int s = 0; s < PRE_ST; ++s) {
        if (s < nst && ((preloaded >> s) & 1u)) {
#pragma unroll
          for (int j = 0; j < KSTEPS; ++j)
            ldsm4(apre[s][j], sbase[s] + a_off + ((uint32_t)((((c.warp * KSTEPS + j) * 2 + a_khalf) ^ a_sw)) << 4));
        }
      }
      const uint8_t* aux = c.ring + (size_t)(c.k % NSLOTS) * SLOT_BYTES;
      uint2 wv = make_uint2(0, 0);
      if (norm && (ready & 1u)) wv = *reinterpret_cast<const uint2*>(aux + c.tid * 8);
      if (kind == K_QKV && has_item) attn_prefetch(c, l, position, item, 0, kv);  // older KV rows do not depend on this layer
      const uint4* xs = reinterpret_cast<const uint4*>(c.s_vec + ((c.lane >> 2) < xr_mod ? (c.lane >> 2) : 0) * SEG_BYTES +
                                                       c.warp * 256 + (c.lane & 3) * 64);
      uint32_t* const my_x = reinterpret_cast<uint32_t*>(c.s_vec + gi0 * 2);
      // which of this lane's accumulator columns carry an item's K-slice partial: bit (2 s + half) of pw_mask set ->
      // row r = lane/4 + 8 half of stage s is an item whose k-block column kb lives in this lane; pw_sel picks odd kb
      uint32_t pw_mask = 0, pw_sel = 0;
#pragma unroll
      for (int sh = 0; sh < 2 * MAX_ST; ++sh) {
        const int r = (c.lane >> 2) + (sh & 1) * 8;
        const int it = (sh >> 1) * STAGE_ITEMS + r;
        const int kb = (xr_mod == 3) ? (int)((kb3_pack >> (2 * sh)) & 3u) : (xr_mod == 2 ? (r & 1) : 0);
        if (r < STAGE_ITEMS && it < d.n_items && (kb >> 1) == (c.lane & 3)) pw_mask |= 1u << sh;
        if (kb & 1) pw_sel |= 1u << sh;
      }
      float* const pw_base = c.s_part + ((c.lane >> 2) * NCW + c.warp);
      QMK_FS(3);
      // -- data --
      bool retried = false;
      if (!from_input) {
        if (QMK_UNLIKELY(!ll4_ok(gw[0], ep))) { retried = true; gw[0] = ll4_wait_slow(p.status, c.s_abort, c.t0, p.timeout_cycles, c.cta, idx, xw + gi0, ep, gi0); }
        if (QMK_UNLIKELY(n_words > H && !ll4_ok(gw[1], ep))) { retried = true; gw[1] = ll4_wait_slow(p.status, c.s_abort, c.t0, p.timeout_cycles, c.cta, idx, xw + gi0 + H, ep, gi0); }
        if (QMK_UNLIKELY(n_words > 2 * H && !ll4_ok(gw[2], ep))) { retried = true; gw[2] = ll4_wait_slow(p.status, c.s_abort, c.t0, p.timeout_cycles, c.cta, idx, xw + gi0 + 2 * H, ep, gi0); }
      }
      trace_sub<TR>(c, 2);
      QMK_FS(4);
      if (norm) {
        // n = r( r(x) * rsqrt(mean(r(x)^2) + eps) * w ); the norm weights arrive with the first stage of the phase
        const float r0 = ll4_val(gw[0].x), r1 = ll4_val(gw[0].y), r2 = ll4_val(gw[0].z), r3 = ll4_val(gw[0].w);
        float ss = fmaf(r0, r0, r1 * r1) + fmaf(r2, r2, r3 * r3);
        ss = warp_sum(ss);
        if (c.lane == 0) c.s_red[c.warp] = ss;
        QMK_FS(5);
        if (from_input) {
          consumer_bar();
          if (sd.in_mode == IN_CODES_SUM && c.tid < rows.o_rows) {   // fp32 residual of this CTA's rows = the summed input
            res_mine = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(c.s_acc)[rows.o_row0 + c.tid]);
            p.res_spill[rows.o_row0 + c.tid] = res_mine;
          }
        } else {
          gather_bar(c, dslot, retried);
        }
        trace_sub<TR>(c, 3);
        QMK_FS(6);
        if (QMK_UNLIKELY(!(ready & 1u))) {
          wait_full(c, c.k);
          ready |= 1u;
          wv = *reinterpret_cast<const uint2*>(aux + c.tid * 8);
        }
        const float4 q0 = *reinterpret_cast<const float4*>(c.s_red), q1 = *reinterpret_cast<const float4*>(c.s_red + 4);
        const float inv = rsqrtf((((q0.x + q0.y) + (q0.z + q0.w)) + ((q1.x + q1.y) + (q1.z + q1.w))) * (1.0f / H) + EPS);
        const __nv_bfloat162 n01 = __floats2bfloat162_rn((r0 * inv) * bf16_lo(wv.x), (r1 * inv) * bf16_hi(wv.x));
        const __nv_bfloat162 n23 = __floats2bfloat162_rn((r2 * inv) * bf16_lo(wv.y), (r3 * inv) * bf16_hi(wv.y));
        const uint2 packed = make_uint2(*reinterpret_cast<const uint32_t*>(&n01), *reinterpret_cast<const uint32_t*>(&n23));
        *reinterpret_cast<uint2*>(my_x) = packed;
        if (kind == K_HEAD && c.cta == 0) {
          if (sd.hidden_out != nullptr)
            *reinterpret_cast<uint2*>(sd.hidden_out + gi0) = make_uint2((gw[0].x & 0xffffu) | (gw[0].y << 16), (gw[0].z & 0xffffu) | (gw[0].w << 16));
          if (sd.out_norm != nullptr)
            *reinterpret_cast<float4*>(sd.out_norm + gi0) = make_float4(bf16_lo(packed.x), bf16_hi(packed.x), bf16_lo(packed.y), bf16_hi(packed.y));
        }
        if (kind == K_QKV && has_item) {
          if (c.warp < 3) {  // q_norm / k_norm weights of this layer (aux of the resident first stage)
            const uint2 nv = *reinterpret_cast<const uint2*>(aux + 2048 + (c.warp == 2 ? 256 : 0) + c.lane * 8);
            pre.nw[0] = bf16_lo(nv.x); pre.nw[1] = bf16_hi(nv.x); pre.nw[2] = bf16_lo(nv.y); pre.nw[3] = bf16_hi(nv.y);
          }
          pre_layer = l;
        }
      } else {
        gather_note(c, dslot, retried);
        *reinterpret_cast<uint2*>(my_x) = make_uint2((gw[0].x & 0xffffu) | (gw[0].y << 16), (gw[0].z & 0xffffu) | (gw[0].w << 16));
        *reinterpret_cast<uint2*>(my_x + H / 2) = make_uint2((gw[1].x & 0xffffu) | (gw[1].y << 16), (gw[1].z & 0xffffu) | (gw[1].w << 16));
        if (n_words > 2 * H)
          *reinterpret_cast<uint2*>(my_x + H) = make_uint2((gw[2].x & 0xffffu) | (gw[2].y << 16), (gw[2].z & 0xffffu) | (gw[2].w << 16));
      }
      __syncwarp();
      trace_sub<TR>(c, 4);
      QMK_FS(7);

      // stages: each warp multiplies its KSTEPS k16-steps of every stage tile (rows = items) by the activation columns
      uint32_t bfrag[KSTEPS][2];
#pragma unroll
      for (int i = 0; i < KSTEPS / 2; ++i) {
        const uint4 v = xs[i];
        bfrag[2 * i][0] = v.x; bfrag[2 * i][1] = v.y; bfrag[2 * i + 1][0] = v.z; bfrag[2 * i + 1][1] = v.w;
      }
      QMK_FS(8);
      float acc[MAX_ST][4];
#pragma unroll
      for (int s = 0; s < MAX_ST; ++s) {
        acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.f;
        if (s < nst) {
          float acc2[4] = {0.f, 0.f, 0.f, 0.f};   // two accumulators halve the dependent HMMA chain
          if (s < PRE_ST && QMK_LIKELY((preloaded >> s) & 1u)) {
#pragma unroll
            for (int j = 0; j < KSTEPS; j += 2) {
              mma16816(acc[s], apre[s][j], bfrag[j]);
              mma16816(acc2, apre[s][j + 1], bfrag[j + 1]);
            }
          } else {
            if (QMK_UNLIKELY(!((ready >> s) & 1u))) wait_full(c, c.k + s);
#pragma unroll
            for (int jb = 0; jb < KSTEPS; jb += 4) {  // batches of 4 ldmatrix keep the live A fragments at 16 registers
              uint32_t afrag[4][4];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                ldsm4(afrag[j], sbase[s] + a_off + ((uint32_t)((((c.warp * KSTEPS + jb + j) * 2 + a_khalf) ^ a_sw)) << 4));
              mma16816(acc[s], afrag[0], bfrag[jb]);
              mma16816(acc2, afrag[1], bfrag[jb + 1]);
              mma16816(acc[s], afrag[2], bfrag[jb + 2]);
              mma16816(acc2, afrag[3], bfrag[jb + 3]);
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[s][e] += acc2[e];
        }
        if (TR) { if (s == 0) fs[9] = __float_as_uint(acc[0][0]) * 0u + (uint32_t)clock(); else if (s == 1) fs[10] = __float_as_uint(acc[1][0]) * 0u + (uint32_t)clock(); else fs[11] = __float_as_uint(acc[2][0]) * 0u + (uint32_t)clock(); }
      }
      c.k += nst;
      trace_sub<TR>(c, 5);
      // per-warp (K-slice) partial of item it -> s_part[it][warp]
#pragma unroll
      for (int sh = 0; sh < 2 * MAX_ST; ++sh) {
        if ((pw_mask >> sh) & 1u)
          pw_base[((sh >> 1) * STAGE_ITEMS + (sh & 1) * 8) * NCW] =
              ((pw_sel >> sh) & 1u) ? acc[sh >> 1][(sh & 1) * 2 + 1] : acc[sh >> 1][(sh & 1) * 2];
      }
      QMK_FS(12);
      consumer_bar();
      QMK_FS(13);
      trace_sub<TR>(c, 6);
      if (c.warp == NCW - 1) prod_issue(c, prod, nst);   // refill the slots this phase released (warps 0-1 finalize meanwhile)

      // finalize + publish: thread t sums the NCW K-slice partials of item t; the items of one output row sit in
      // neighbouring lanes (an output row never straddles a warp: 2 | 32, and 3 * rows <= 30)
      if (kind != K_O && kind != K_DOWN) {
        if (c.warp < 2) {   // n_items <= 42: warps 0 and 1
          const bool mine = c.tid < d.n_items;
          const float v = mine ? item_sum(c.s_part, c.tid) : 0.f;
          if (kind == K_GU) {  // gate in even lanes, up in the next lane
            const float u = bf16_round(__shfl_down_sync(0xffffffffu, v, 1));
            if (mine && (c.tid & 1) == 0) {
              const float g = bf16_round(v);
              const float sg = bf16_round(__fdividef(g, 1.0f + __expf(-g)));
              ll4_st(xout + (c.tid >> 1), sg * u, epoch);
            }
          } else if (mine) {
            ll4_st(xout + c.tid, v, epoch);
          }
        }
      } else if (c.warp == 0) {   // 2 or 3 items per output row, all inside warp 0
        const float v = c.lane < d.n_items ? item_sum(c.s_part, c.lane) : 0.f;
        const int per = (kind == K_O) ? 2 : 3;
        const int src = c.lane * per;
        float tot = __shfl_sync(0xffffffffu, v, src & 31) + __shfl_sync(0xffffffffu, v, (src + 1) & 31);
        const float v2 = __shfl_sync(0xffffffffu, v, (src + 2) & 31);
        if (per == 3) tot += v2;
        if (c.lane < rows.o_rows) {
          const float o = bf16_round(tot);
          res_mine = p.residual_fp32 ? res_mine + o : bf16_round(res_mine + o);
          ll4_st(xout + c.lane, res_mine, epoch);
          p.res_spill[rows.o_row0 + c.lane] = res_mine;
        }
      }
      c.t_pub = clock64();
      QMK_FS(14);
      if (TR && c.p.trace != nullptr && c.tid == 0) {
        const int slot0 = (c.cur_idx - c.p.phase_begin) * TRACE_SUBS + 8;
        if (slot0 + 16 <= c.p.trace_stride) {
#pragma unroll
          for (int i = 0; i < 16; ++i) c.p.trace[(size_t)c.cta * c.p.trace_stride + slot0 + i] = (long long)fs[i];
        }
      }
      trace_sub<TR>(c, 7);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <bool TR>
__device__ __forceinline__ void decode_kernel_body(const Params& p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctx c(p);
  c.ring = smem + SM_RING;
  c.s_vec = smem + SM_VEC;
  c.s_acc = reinterpret_cast<float*>(smem + SM_ACC);
  c.s_small = reinterpret_cast<float*>(smem + SM_SMALL);
  c.s_part = reinterpret_cast<float*>(smem + SM_PART);
  c.s_red = reinterpret_cast<float*>(smem + SM_RED);
  c.full = reinterpret_cast<u64*>(smem + SM_BAR);
  c.s_abort = reinterpret_cast<volatile int*>(smem + SM_MISC);
  c.s_delay = reinterpret_cast<int*>(smem + SM_MISC + 16);
  c.s_tbl = reinterpret_cast<uint4*>(smem + SM_TBL);
  c.x32 = reinterpret_cast<uint32_t*>(p.xbuf);
  c.tid = threadIdx.x;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.cta = blockIdx.x;
  c.k = 0;
  c.t0 = clock64();
  c.t_pub = c.t0;
  c.cur_idx = p.phase_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOTS; ++i) {
      mbar_init(&c.full[i], 1);
    }
    *c.s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 3 * DL_N) c.s_delay[threadIdx.x] = p.delays[blockIdx.x * 3 * DL_N + threadIdx.x];
  __syncthreads();

  consumer_loop<TR>(c);
  __syncthreads();
  c.cur_idx = p.phase_end;
  trace_sub<TR>(c, 0);
  if (threadIdx.x >= DL_N && threadIdx.x < 3 * DL_N) p.delays[blockIdx.x * 3 * DL_N + threadIdx.x] = c.s_delay[threadIdx.x];
  if (*c.s_abort && threadIdx.x == 0) {
    // failure path only: give in-flight bulk copies time to land before the CTA's shared memory is released
    const long long t = clock64();
    while (clock64() - t < 2000000) {
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) qmk_decode_kernel(const __grid_constant__ Params p) {
  decode_kernel_body<false>(p);
}
// same kernel with the per-phase clock trace compiled in (qmk_engine_trace_enable)
__global__ void __launch_bounds__(NTHREADS, 1) qmk_decode_kernel_traced(const __grid_constant__ Params p) {
  decode_kernel_body<true>(p);
}

// ------------------------------------------------------------------------------------------------
// weight re-packing (one-time): upstream [out,in] row-major bf16 -> per-CTA segment streams
// ------------------------------------------------------------------------------------------------
struct LayerPtrs {  // = upstream LDGLayerWeights (kernel.cu:78-90)
  const uint4* w[11];
};
enum { W_IN = 0, W_Q, W_K, W_V, W_QN, W_KN, W_O, W_POST, W_GATE, W_UP, W_DOWN };

// Packed order of one 2 KB row segment (1024 k): the segment is 8 K-slices of 128 (one per consumer warp); inside a
// slice, 16-byte chunk m (= tensor-core step j = m / 2, k-half b = m % 2) holds, for q = 0..3, the two elements
// k = 32 q + 2 m + {0, 1} of the slice.  With this order the B fragment of lane q is the CONTIGUOUS run
// x[32 q .. 32 q + 32) of the activation slice (4 x LDS.128 instead of 16 x LDS.32).  Chunk t of stage row r is
// stored at chunk t ^ (r & 7) (conflict-free ldmatrix of the 16 x 16 tiles).
__device__ __forceinline__ uint4 pack_chunk(const uint32_t* row_words, int t) {
  const int w = t >> 4, m = t & 15;
  const uint32_t* s = row_words + w * 64 + m;
  return make_uint4(s[0], s[16], s[32], s[48]);
}

// grid = (G, L), block = 128 threads: thread t builds 16-byte chunk t of each 2 KB segment.
__global__ void qmk_pack_layers_kernel(const LayerPtrs* layers, Layout y, uint8_t* packed, uint8_t* aux_layers) {
  const int cta = blockIdx.x, l = blockIdx.y, t = threadIdx.x;
  const LayerPtrs lp = layers[l];
  uint4* dst = reinterpret_cast<uint4*>(packed + ((size_t)((size_t)cta * y.L + l) * y.layer_segs) * SEG_BYTES);
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const int q0 = row_begin(cta, QKV_ROWS, y.G), nq = row_begin(cta + 1, QKV_ROWS, y.G) - q0;
  const int o0 = row_begin(cta, H, y.G), no = row_begin(cta + 1, H, y.G) - o0;
  const int g0 = row_begin(cta, INTER, y.G), ng = row_begin(cta + 1, INTER, y.G) - g0;
  for (int seg = 0; seg < y.layer_segs; ++seg) {
    const uint32_t* src = nullptr;   // first 32-bit word of the 1024-element source segment
    int s_in_phase;
    if (seg < y.off_o) {
      const int s = seg;
      s_in_phase = s;
      if (s < nq) {
        const int row = q0 + s;
        if (row < QSZ) src = reinterpret_cast<const uint32_t*>(lp.w[W_Q]) + (size_t)row * 512;
        else if (row < QSZ + KVSZ) src = reinterpret_cast<const uint32_t*>(lp.w[W_K]) + (size_t)(row - QSZ) * 512;
        else src = reinterpret_cast<const uint32_t*>(lp.w[W_V]) + (size_t)(row - QSZ - KVSZ) * 512;
      }
    } else if (seg < y.off_gu) {
      const int s = seg - y.off_o;
      s_in_phase = s;
      if (s < 2 * no) src = reinterpret_cast<const uint32_t*>(lp.w[W_O]) + ((size_t)(o0 + s / 2) * 2 + (s % 2)) * 512;
    } else if (seg < y.off_down) {
      const int s = seg - y.off_gu;
      s_in_phase = s;
      if (s < 2 * ng) {
        const int row = g0 + s / 2;
        src = reinterpret_cast<const uint32_t*>((s % 2 == 0) ? lp.w[W_GATE] : lp.w[W_UP]) + (size_t)row * 512;
      }
    } else {
      const int s = seg - y.off_down;
      s_in_phase = s;
      if (s < 3 * no) src = reinterpret_cast<const uint32_t*>(lp.w[W_DOWN]) + ((size_t)(o0 + s / 3) * 3 + (s % 3)) * 512;
    }
    dst[(size_t)seg * 128 + (t ^ ((s_in_phase % STAGE_ITEMS) & 7))] = src ? pack_chunk(src, t) : zero;
  }
  if (cta == 0) {  // shared aux blocks: [input_ln | q_norm | k_norm] and [post_ln | 0]
    uint4* a0 = reinterpret_cast<uint4*>(aux_layers + ((size_t)l * 2 + 0) * AUX_BYTES);
    uint4* a1 = reinterpret_cast<uint4*>(aux_layers + ((size_t)l * 2 + 1) * AUX_BYTES);
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

// grid = G, block = 128: rows of this CTA
__global__ void qmk_pack_head_kernel(const uint4* head_w, int rows, int G, int segs_max, uint8_t* packed) {
  const int cta = blockIdx.x, t = threadIdx.x;
  uint4* dst = reinterpret_cast<uint4*>(packed + ((size_t)cta * segs_max) * SEG_BYTES);
  const int r0 = row_begin(cta, rows, G), n = row_begin(cta + 1, rows, G) - r0;
  for (int seg = 0; seg < segs_max; ++seg) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (seg < n) v = pack_chunk(reinterpret_cast<const uint32_t*>(head_w) + (size_t)(r0 + seg) * 512, t);
    dst[(size_t)seg * 128 + (t ^ ((seg % STAGE_ITEMS) & 7))] = v;
  }
}

}  // namespace qmk
