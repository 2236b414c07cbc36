// qmk_batched.cu — batched multi-stream decode (B = 16 .. 64 concurrent utterances), SURVEY.md section 8a row 18.
//
// At B >= 16 the projections are real dense contractions ([rows, K] x [K, B]); they run on the 5th-generation tensor
// cores (tcgen05.mma with TMEM accumulators, qmk_bgemm.cuh), every weight byte is read once per step for all B
// streams, and the work between the GEMMs (split-K reduction, bf16 rounding points, RMSNorm, per-head norm + RoPE,
// KV append, GQA attention, SwiGLU, residual, argmax) is a handful of small fused kernels.  Each stream is numerically
// the B = 1 path: same rounding points (oracle/tts_oracle.py), own position, own KV cache [B][L][8][S][128].
// There is no upstream counterpart (upstream is strictly B = 1); weights are read in the upstream [out, in] layout.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "qmk_b200.h"
#include "qmk_bgemm.cuh"
#include "qmk_bstep.cuh"
#include "qmk_sample.cuh"

namespace {

constexpr int H = 1024, INTER = 3072, QSZ = 2048, KVSZ = 1024, HD = 128, NKVH = 8;
constexpr int QKV_ROWS = 4096, GU_ROWS = 6144;
constexpr float EPS = 1e-6f;

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t bf16_bits(float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Split-K partials are staged through shared memory with cp.async: a thread's 8-16 independent 16-byte reads are then ALL in
// flight at once (one L2 round trip).  With ordinary loads the compiler interleaves loads and dependent adds three at a time
// -- five round trips, 1.4 us of a 2.1-us kernel (scripts/chain_trace.py) -- whatever the source order.  A thread only reads
// back what it copied itself, so cp.async.wait_all is all the synchronisation needed.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// sum over the 256 threads of a block (8 warps); every thread gets the total
__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += s_red[w];
  __syncthreads();
  return t;
}
// n = r( r(x) * rsqrt(mean(r(x)^2) + eps) * w ) for the 4 elements of this thread (same arithmetic as the B = 1 kernel)
__device__ __forceinline__ uint2 rmsnorm4(const float (&x)[4], const __nv_bfloat16* w, float* s_red) {
  float r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) r[e] = bf16_round(x[e]);
  const float tot = block_sum_256(fmaf(r[0], r[0], r[1] * r[1]) + fmaf(r[2], r[2], r[3] * r[3]), s_red);
  const float inv = rsqrtf(tot * (1.0f / H) + EPS);
  const uint2 wv = *reinterpret_cast<const uint2*>(w + threadIdx.x * 4);
  const __nv_bfloat162 n01 = __floats2bfloat162_rn((r[0] * inv) * bf16_lo(wv.x), (r[1] * inv) * bf16_hi(wv.x));
  const __nv_bfloat162 n23 = __floats2bfloat162_rn((r[2] * inv) * bf16_lo(wv.y), (r[3] * inv) * bf16_hi(wv.y));
  return make_uint2(*reinterpret_cast<const uint32_t*>(&n01), *reinterpret_cast<const uint32_t*>(&n23));
}

// ---- step input: embedding row or caller vector -> fp32 residual + layer-0 input norm -------------------------------
// grid = B, block = 256
__global__ void kb_input(const int* token_ids, const __nv_bfloat16* embed_table, int vocab, const __nv_bfloat16* embeds,
                         const float* embeds_f32, float* res, const __nv_bfloat16* w_in, __nv_bfloat16* xn, int n_valid) {
  __shared__ float s_red[8];
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  const int b = blockIdx.x, t = threadIdx.x;
  const int bs = b < n_valid ? b : n_valid - 1;   // prefill with fewer positions than lanes: the spare lanes repeat the last one
  float x[4];
  if (embeds_f32 != nullptr) {      // fp32 vector rounded to bf16 on load (the talker's hidden state entering the code predictor)
    const float4 v = *reinterpret_cast<const float4*>(embeds_f32 + (size_t)bs * H + t * 4);
    x[0] = bf16_round(v.x); x[1] = bf16_round(v.y); x[2] = bf16_round(v.z); x[3] = bf16_round(v.w);
  } else {
    int tok = token_ids ? token_ids[bs] : -1;
    if (tok >= vocab) tok = vocab - 1;             // device-side ids are not trusted: clamp instead of reading out of bounds
    if (tok < 0 && embeds == nullptr) tok = 0;     // sentinel without an embedding buffer
    const __nv_bfloat16* src = tok >= 0 ? embed_table + (size_t)tok * H : embeds + (size_t)bs * H;
    const uint2 v = *reinterpret_cast<const uint2*>(src + t * 4);
    x[0] = bf16_lo(v.x); x[1] = bf16_hi(v.x); x[2] = bf16_lo(v.y); x[3] = bf16_hi(v.y);
  }
  *reinterpret_cast<float4*>(res + (size_t)b * H + t * 4) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<uint2*>(xn + (size_t)b * H + t * 4) = rmsnorm4(x, w_in, s_red);
}

// ---- O / down epilogue: split-K sum -> bf16 -> residual -> next RMSNorm ----------------------------------------------
// One CLUSTER of C CTAs per stream, 256 threads in all (thread v owns elements 4 v .. 4 v + 3).  partial: [SPLITS][B][1024].
// hidden_out (optional, final norm only): f32[B][1024].  The kernel is nothing but the latency of its loads: all SPLITS partial
// loads of a thread are in flight together (one L2 round trip), and the 64 KB a stream reads are spread over C SMs (one SM
// sustains ~45 GB/s of such loads: 1.4 us for 64 KB).  The sum of squares keeps the single-CTA order: warp totals travel
// through distributed shared memory into every CTA's s_red[8] and are added in warp order.
template <int SPLITS, int C>
__global__ void __launch_bounds__(256 / C) kb_resid_norm(const float* partial, int B, float* res, int residual_fp32,
                                                         const __nv_bfloat16* w_norm, __nv_bfloat16* xn, float* hidden_out, int* advance_positions) {
  __shared__ float s_red[8];
  __shared__ float4 s_pp[SPLITS][256 / C];
  qmkb::KTrace kt;
  kt.mark(0);
  unsigned rank = 0;
  if (C > 1) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");   // every CTA of the cluster runs (ahead of the dependency)
  }
  const int b = blockIdx.x / C, t = (int)rank * (256 / C) + threadIdx.x;
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  kt.mark(1);
#pragma unroll
  for (int s = 0; s < SPLITS; ++s) cp_async16(&s_pp[s][threadIdx.x], partial + ((size_t)s * B + b) * H + t * 4);
  float4 r = *reinterpret_cast<const float4*>(res + (size_t)b * H + t * 4);
  const uint2 wv = *reinterpret_cast<const uint2*>(w_norm + t * 4);
  cp_async_wait_all();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int s = 0; s < SPLITS; ++s) { const float4 pv = s_pp[s][threadIdx.x]; acc.x += pv.x; acc.y += pv.y; acc.z += pv.z; acc.w += pv.w; }
  if (kt.on && acc.x == 1.2345e-30f) kt.t[0] = 0;   // (trace only) the stamp below waits for the loads
  kt.mark(2);
  const float o[4] = {bf16_round(acc.x), bf16_round(acc.y), bf16_round(acc.z), bf16_round(acc.w)};
  float x[4] = {r.x + o[0], r.y + o[1], r.z + o[2], r.w + o[3]};
  if (!residual_fp32) {
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = bf16_round(x[e]);
  }
  *reinterpret_cast<float4*>(res + (size_t)b * H + t * 4) = make_float4(x[0], x[1], x[2], x[3]);
  // n = r( r(x) * rsqrt(mean(r(x)^2) + eps) * w ): same arithmetic and summation order as rmsnorm4
  float rr[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) rr[e] = bf16_round(x[e]);
  const float ws = warp_sum(fmaf(rr[0], rr[0], rr[1] * rr[1]) + fmaf(rr[2], rr[2], rr[3] * rr[3]));
  if ((threadIdx.x & 31) == 0) {
    if (C > 1) {
      const uint32_t local = (uint32_t)__cvta_generic_to_shared(&s_red[t >> 5]);
#pragma unroll
      for (int peer = 0; peer < C; ++peer) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(peer));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(ws) : "memory");
      }
    } else {
      s_red[t >> 5] = ws;
    }
  }
  if (C > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  else __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += s_red[w];
  const float inv = rsqrtf(tot * (1.0f / H) + EPS);
  const __nv_bfloat162 n01 = __floats2bfloat162_rn((rr[0] * inv) * bf16_lo(wv.x), (rr[1] * inv) * bf16_hi(wv.x));
  const __nv_bfloat162 n23 = __floats2bfloat162_rn((rr[2] * inv) * bf16_lo(wv.y), (rr[3] * inv) * bf16_hi(wv.y));
  const uint2 n = make_uint2(*reinterpret_cast<const uint32_t*>(&n01), *reinterpret_cast<const uint32_t*>(&n23));
  *reinterpret_cast<uint2*>(xn + (size_t)b * H + t * 4) = n;
  if (hidden_out)
    *reinterpret_cast<float4*>(hidden_out + (size_t)b * H + t * 4) = make_float4(bf16_lo(n.x), bf16_hi(n.x), bf16_lo(n.y), bf16_hi(n.y));
  if (advance_positions != nullptr && t == 0) advance_positions[b] += 1;   // a step without an LM head ends here
  kt.flush(3);
}

// ---- QKV epilogue + decode attention, one CTA per (stream, kv head) --------------------------------------------------
// Warps 0-3 first finish the projection for this group: split-K sum -> bf16 -> per-head RMSNorm + rotate-half RoPE in
// bf16 steps (q heads 2g, 2g+1 and k) -> q, and the new k / v row, into shared memory; k / v are also appended to the cache row
// positions[b].  Then all ATT_NW warps stride over positions 0 .. pos (lane owns 4 dims), fp32 online softmax, fixed-order
// cross-warp merge.  The cached rows of the first ATT_PRE rounds (positions < 32) are requested BEFORE the grid
// dependency resolves: they were written by earlier steps, so their L2 / HBM latency overlaps the projection that precedes
// this kernel; deeper contexts continue in batches of ATT_PRE independent row loads.
// Grid shape, measured (us per B = 16 / B = 64 step, graph replay): one item per 256-thread CTA 656 / 864 (kept); four warps per
// item with eight rows each in flight 694 / 866; two / four items per 512- / 1024-thread CTA - / 903, - / 930.  An EMPTY
// 512 x 256 grid costs 4.8 us in a dependent chain (scripts/ubench/pdl_gap.cu), but programmatic launch hides that behind the
// preceding projection, and fat CTAs keep the next projection's CTAs from becoming resident early.
// grid = (B, 8), block = 256.  partial: [SPLITS][B][4096] (q rows 0..2047, k 2048..3071, v 3072..4095).
#ifndef QMK_BATT_PRE
#define QMK_BATT_PRE 4
#endif
constexpr int ATT_PRE = QMK_BATT_PRE, ATT_NW = 8;
template <int SPLITS>
__global__ void __launch_bounds__(32 * ATT_NW) kb_qkv_attention(const float* partial, int B, const int* positions, const __nv_bfloat16* q_norm,
                                 const __nv_bfloat16* k_norm, const __nv_bfloat16* cos_t, const __nv_bfloat16* sin_t,
                                 __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, __nv_bfloat16* a_out, int layer, int L,
                                 int max_seq, float scale) {
  __shared__ float s_q[2][HD];
  __shared__ float s_kv[2][HD];   // the new row: k (after norm + RoPE), v
  __shared__ float s_acc[ATT_NW][2][HD];
  __shared__ float s_m[ATT_NW][2], s_l[ATT_NW][2];
  qmkb::KTrace kt;
  kt.mark(0);
  const int b = blockIdx.x, g = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // positions[] was last written two kernels ago at the latest (the previous step's final kernel; every kernel of the chain
  // resolves its own dependency before it lets its successor start), so it may be read ahead of the dependency.
  int pos = positions[b];
  pos = pos < 0 ? 0 : (pos >= max_seq ? max_seq - 1 : pos);   // a stream that ran past its cache keeps rewriting the last row
  const size_t base = (((size_t)b * L + layer) * NKVH + g) * max_seq * HD;
  uint2 kk[ATT_PRE], vv[ATT_PRE];
#pragma unroll
  for (int i = 0; i < ATT_PRE; ++i) {
    const int p = warp + ATT_NW * i;
    if (p < pos) {
      kk[i] = __ldcg(reinterpret_cast<const uint2*>(k_cache + base + (size_t)p * HD + lane * 4));
      vv[i] = __ldcg(reinterpret_cast<const uint2*>(v_cache + base + (size_t)p * HD + lane * 4));
    }
  }
  const __nv_bfloat16* wn = warp < 2 ? q_norm : k_norm;
  const uint2 wn_raw = warp < 3 ? *reinterpret_cast<const uint2*>(wn + lane * 4) : make_uint2(0u, 0u);
  const int dbase = (lane * 4) & 63;
  const uint2 cs_raw = *reinterpret_cast<const uint2*>(cos_t + (size_t)pos * HD + dbase), sn_raw = *reinterpret_cast<const uint2*>(sin_t + (size_t)pos * HD + dbase);
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  kt.mark(1);
  if (warp < 4) {
    // warp 0, 1: q heads 2g, 2g+1; warp 2: k head g; warp 3: v head g
    const int row0 = warp < 2 ? (2 * g + warp) * HD : (warp == 2 ? QSZ + g * HD : QSZ + KVSZ + g * HD);
    float4 pp[SPLITS];
#pragma unroll
    for (int s = 0; s < SPLITS; ++s) pp[s] = __ldcg(reinterpret_cast<const float4*>(partial + ((size_t)s * B + b) * QKV_ROWS + row0 + lane * 4));
    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < SPLITS; ++s) { t[0] += pp[s].x; t[1] += pp[s].y; t[2] += pp[s].z; t[3] += pp[s].w; }
#pragma unroll
    for (int e = 0; e < 4; ++e) t[e] = bf16_round(t[e]);
    if (warp == 3) {
      *reinterpret_cast<float4*>(&s_kv[1][lane * 4]) = make_float4(t[0], t[1], t[2], t[3]);
      *reinterpret_cast<uint2*>(v_cache + base + (size_t)pos * HD + lane * 4) =
          make_uint2(bf16_bits(t[0]) | (bf16_bits(t[1]) << 16), bf16_bits(t[2]) | (bf16_bits(t[3]) << 16));
    } else {
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) ss = fmaf(t[e], t[e], ss);
      ss = warp_sum(ss);
      const float rms = sqrtf(ss * (1.0f / HD) + EPS);
      const float wf[4] = {bf16_lo(wn_raw.x), bf16_hi(wn_raw.x), bf16_lo(wn_raw.y), bf16_hi(wn_raw.y)};
      const float cf[4] = {bf16_lo(cs_raw.x), bf16_hi(cs_raw.x), bf16_lo(cs_raw.y), bf16_hi(cs_raw.y)};
      const float sf[4] = {bf16_lo(sn_raw.x), bf16_hi(sn_raw.x), bf16_lo(sn_raw.y), bf16_hi(sn_raw.y)};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float n = bf16_round((t[e] / rms) * wf[e]);
        const float other = __shfl_xor_sync(0xffffffffu, n, 16);
        const float x = bf16_round(n * cf[e]), y = bf16_round(other * sf[e]);
        o[e] = bf16_round(lane < 16 ? x - y : x + y);
      }
      if (warp < 2) {
        *reinterpret_cast<float4*>(&s_q[warp][lane * 4]) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        *reinterpret_cast<float4*>(&s_kv[0][lane * 4]) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint2*>(k_cache + base + (size_t)pos * HD + lane * 4) =
            make_uint2(bf16_bits(o[0]) | (bf16_bits(o[1]) << 16), bf16_bits(o[2]) | (bf16_bits(o[3]) << 16));
      }
    }
  }
  __syncthreads();   // q and the new k / v row are in shared memory
  kt.mark(2);
  const int n = pos + 1;
  const float4 qa = *reinterpret_cast<const float4*>(&s_q[0][lane * 4]), qb = *reinterpret_cast<const float4*>(&s_q[1][lane * 4]);
  const float q0[4] = {qa.x, qa.y, qa.z, qa.w}, q1[4] = {qb.x, qb.y, qb.z, qb.w};
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f, acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
  for (int p0 = warp; p0 < n; p0 += ATT_NW * ATT_PRE) {   // warp-uniform; positions in ascending order per warp, as before
    if (p0 != warp) {
#pragma unroll
      for (int i = 0; i < ATT_PRE; ++i) {
        const int p = p0 + ATT_NW * i;
        if (p < pos) {
          kk[i] = __ldcg(reinterpret_cast<const uint2*>(k_cache + base + (size_t)p * HD + lane * 4));
          vv[i] = __ldcg(reinterpret_cast<const uint2*>(v_cache + base + (size_t)p * HD + lane * 4));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < ATT_PRE; ++i) {
      const int p = p0 + ATT_NW * i;
      if (p < n) {
        float kf[4], vf[4];
        if (p == pos) {
          const float4 a = *reinterpret_cast<const float4*>(&s_kv[0][lane * 4]), c = *reinterpret_cast<const float4*>(&s_kv[1][lane * 4]);
          kf[0] = a.x; kf[1] = a.y; kf[2] = a.z; kf[3] = a.w; vf[0] = c.x; vf[1] = c.y; vf[2] = c.z; vf[3] = c.w;
        } else {
          kf[0] = bf16_lo(kk[i].x); kf[1] = bf16_hi(kk[i].x); kf[2] = bf16_lo(kk[i].y); kf[3] = bf16_hi(kk[i].y);
          vf[0] = bf16_lo(vv[i].x); vf[1] = bf16_hi(vv[i].x); vf[2] = bf16_lo(vv[i].y); vf[3] = bf16_hi(vv[i].y);
        }
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) { d0 = fmaf(q0[e], kf[e], d0); d1 = fmaf(q1[e], kf[e], d1); }
        d0 = warp_sum(d0) * scale;
        d1 = warp_sum(d1) * scale;
        const float nm0 = fmaxf(m0, d0), nm1 = fmaxf(m1, d1);
        const float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - nm0), c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - nm1);
        const float e0 = __expf(d0 - nm0), e1 = __expf(d1 - nm1);
        l0 = l0 * c0 + e0; l1 = l1 * c1 + e1;
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc0[e] = fmaf(e0, vf[e], acc0[e] * c0); acc1[e] = fmaf(e1, vf[e], acc1[e] * c1); }
        m0 = nm0; m1 = nm1;
      }
    }
  }
  if (lane == 0) { s_m[warp][0] = m0; s_m[warp][1] = m1; s_l[warp][0] = l0; s_l[warp][1] = l1; }
#pragma unroll
  for (int e = 0; e < 4; ++e) { s_acc[warp][0][lane * 4 + e] = acc0[e]; s_acc[warp][1][lane * 4 + e] = acc1[e]; }
  __syncthreads();
  for (int o = tid; o < 2 * HD; o += 32 * ATT_NW) {   // warps merged in warp order
    const int h = o >> 7, d = o & 127;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < ATT_NW; ++w) M = fmaxf(M, s_m[w][h]);
    float A = 0.f, Ls = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_NW; ++w) {
      const float f = (s_m[w][h] == -INFINITY) ? 0.f : __expf(s_m[w][h] - M);
      A = fmaf(s_acc[w][h][d], f, A);
      Ls = fmaf(s_l[w][h], f, Ls);
    }
    a_out[(size_t)b * QSZ + (2 * g + h) * HD + d] = __float2bfloat16_rn(A / Ls);
  }
  kt.flush(4);
}

// ---- decode attention with the context split over Z CTAs (deep contexts at B <= 32: one CTA per (stream, kv head) keeps only
//      16 KB of rows in flight, 1.1 TB/s of KV at B = 16 and 2000 positions).  CTA z takes positions [z C, (z + 1) C) of 0 .. pos,
//      every CTA finishes q itself (four small loads), the CTA of the last chunk also appends the new K / V row; the partial
//      (max, sum, accumulators) of a chunk goes to global memory and the LAST CTA of the item to arrive (one atomic ticket)
//      merges the Z partials in chunk order -- deterministic -- and writes the output.  grid = (B, 8, Z), block = 256.
//      part: f32[B][8][Z][2][130]; tickets: u32[B][8], zero between kernels (the merging CTA resets its ticket). --------------
constexpr int ATT_PART = 130, ATT_SPLIT_MIN = 256, ATT_SPLIT_MIN_WIDE = 1024;
template <int SPLITS, int Z>
__global__ void __launch_bounds__(256) kb_qkv_attention_split(const float* partial, int B, const int* positions, const __nv_bfloat16* q_norm,
                                                              const __nv_bfloat16* k_norm, const __nv_bfloat16* cos_t, const __nv_bfloat16* sin_t,
                                                              __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, __nv_bfloat16* a_out, float* part,
                                                              unsigned* tickets, int layer, int L, int max_seq, float scale) {
  __shared__ float s_q[2][HD];
  __shared__ float s_kv[2][HD];
  __shared__ float s_acc[ATT_NW][2][HD];
  __shared__ float s_m[ATT_NW][2], s_l[ATT_NW][2];
  __shared__ unsigned s_last;
  const int b = blockIdx.x, g = blockIdx.y, z = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int pos = positions[b];
  pos = pos < 0 ? 0 : (pos >= max_seq ? max_seq - 1 : pos);
  const int n = pos + 1;
  const int C = ((n + Z - 1) / Z + 7) & ~7;                  // chunk length, a multiple of the 8 warps
  const int p_lo = z * C, p_hi = p_lo + C < n ? p_lo + C : n;   // may be empty for the trailing chunks of a short context
  const bool owner = pos >= p_lo && pos < p_lo + C;          // the chunk that holds the new position appends the row
  const size_t base = (((size_t)b * L + layer) * NKVH + g) * max_seq * HD;
  uint2 kk[ATT_PRE], vv[ATT_PRE];
#pragma unroll
  for (int i = 0; i < ATT_PRE; ++i) {
    const int p = p_lo + warp + ATT_NW * i;
    if (p < p_hi && p < pos) {
      kk[i] = __ldcg(reinterpret_cast<const uint2*>(k_cache + base + (size_t)p * HD + lane * 4));
      vv[i] = __ldcg(reinterpret_cast<const uint2*>(v_cache + base + (size_t)p * HD + lane * 4));
    }
  }
  const __nv_bfloat16* wn = warp < 2 ? q_norm : k_norm;
  const uint2 wn_raw = warp < 3 ? *reinterpret_cast<const uint2*>(wn + lane * 4) : make_uint2(0u, 0u);
  const int dbase = (lane * 4) & 63;
  const uint2 cs_raw = *reinterpret_cast<const uint2*>(cos_t + (size_t)pos * HD + dbase), sn_raw = *reinterpret_cast<const uint2*>(sin_t + (size_t)pos * HD + dbase);
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  if (warp < 2 || (owner && warp < 4)) {
    const int row0 = warp < 2 ? (2 * g + warp) * HD : (warp == 2 ? QSZ + g * HD : QSZ + KVSZ + g * HD);
    float4 pp[SPLITS];
#pragma unroll
    for (int s = 0; s < SPLITS; ++s) pp[s] = __ldcg(reinterpret_cast<const float4*>(partial + ((size_t)s * B + b) * QKV_ROWS + row0 + lane * 4));
    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < SPLITS; ++s) { t[0] += pp[s].x; t[1] += pp[s].y; t[2] += pp[s].z; t[3] += pp[s].w; }
#pragma unroll
    for (int e = 0; e < 4; ++e) t[e] = bf16_round(t[e]);
    if (warp == 3) {
      *reinterpret_cast<float4*>(&s_kv[1][lane * 4]) = make_float4(t[0], t[1], t[2], t[3]);
      *reinterpret_cast<uint2*>(v_cache + base + (size_t)pos * HD + lane * 4) =
          make_uint2(bf16_bits(t[0]) | (bf16_bits(t[1]) << 16), bf16_bits(t[2]) | (bf16_bits(t[3]) << 16));
    } else {
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) ss = fmaf(t[e], t[e], ss);
      ss = warp_sum(ss);
      const float rms = sqrtf(ss * (1.0f / HD) + EPS);
      const float wf[4] = {bf16_lo(wn_raw.x), bf16_hi(wn_raw.x), bf16_lo(wn_raw.y), bf16_hi(wn_raw.y)};
      const float cf[4] = {bf16_lo(cs_raw.x), bf16_hi(cs_raw.x), bf16_lo(cs_raw.y), bf16_hi(cs_raw.y)};
      const float sf[4] = {bf16_lo(sn_raw.x), bf16_hi(sn_raw.x), bf16_lo(sn_raw.y), bf16_hi(sn_raw.y)};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float nv = bf16_round((t[e] / rms) * wf[e]);
        const float other = __shfl_xor_sync(0xffffffffu, nv, 16);
        const float x = bf16_round(nv * cf[e]), y = bf16_round(other * sf[e]);
        o[e] = bf16_round(lane < 16 ? x - y : x + y);
      }
      if (warp < 2) {
        *reinterpret_cast<float4*>(&s_q[warp][lane * 4]) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        *reinterpret_cast<float4*>(&s_kv[0][lane * 4]) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint2*>(k_cache + base + (size_t)pos * HD + lane * 4) =
            make_uint2(bf16_bits(o[0]) | (bf16_bits(o[1]) << 16), bf16_bits(o[2]) | (bf16_bits(o[3]) << 16));
      }
    }
  }
  __syncthreads();
  const float4 qa = *reinterpret_cast<const float4*>(&s_q[0][lane * 4]), qb = *reinterpret_cast<const float4*>(&s_q[1][lane * 4]);
  const float q0[4] = {qa.x, qa.y, qa.z, qa.w}, q1[4] = {qb.x, qb.y, qb.z, qb.w};
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f, acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
  for (int p0 = p_lo + warp; p0 < p_hi; p0 += ATT_NW * ATT_PRE) {
    if (p0 != p_lo + warp) {
#pragma unroll
      for (int i = 0; i < ATT_PRE; ++i) {
        const int p = p0 + ATT_NW * i;
        if (p < p_hi && p < pos) {
          kk[i] = __ldcg(reinterpret_cast<const uint2*>(k_cache + base + (size_t)p * HD + lane * 4));
          vv[i] = __ldcg(reinterpret_cast<const uint2*>(v_cache + base + (size_t)p * HD + lane * 4));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < ATT_PRE; ++i) {
      const int p = p0 + ATT_NW * i;
      if (p < p_hi) {
        float kf[4], vf[4];
        if (p == pos) {
          const float4 a = *reinterpret_cast<const float4*>(&s_kv[0][lane * 4]), c = *reinterpret_cast<const float4*>(&s_kv[1][lane * 4]);
          kf[0] = a.x; kf[1] = a.y; kf[2] = a.z; kf[3] = a.w; vf[0] = c.x; vf[1] = c.y; vf[2] = c.z; vf[3] = c.w;
        } else {
          kf[0] = bf16_lo(kk[i].x); kf[1] = bf16_hi(kk[i].x); kf[2] = bf16_lo(kk[i].y); kf[3] = bf16_hi(kk[i].y);
          vf[0] = bf16_lo(vv[i].x); vf[1] = bf16_hi(vv[i].x); vf[2] = bf16_lo(vv[i].y); vf[3] = bf16_hi(vv[i].y);
        }
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) { d0 = fmaf(q0[e], kf[e], d0); d1 = fmaf(q1[e], kf[e], d1); }
        d0 = warp_sum(d0) * scale;
        d1 = warp_sum(d1) * scale;
        const float nm0 = fmaxf(m0, d0), nm1 = fmaxf(m1, d1);
        const float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - nm0), c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - nm1);
        const float e0 = __expf(d0 - nm0), e1 = __expf(d1 - nm1);
        l0 = l0 * c0 + e0; l1 = l1 * c1 + e1;
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc0[e] = fmaf(e0, vf[e], acc0[e] * c0); acc1[e] = fmaf(e1, vf[e], acc1[e] * c1); }
        m0 = nm0; m1 = nm1;
      }
    }
  }
  if (lane == 0) { s_m[warp][0] = m0; s_m[warp][1] = m1; s_l[warp][0] = l0; s_l[warp][1] = l1; }
#pragma unroll
  for (int e = 0; e < 4; ++e) { s_acc[warp][0][lane * 4 + e] = acc0[e]; s_acc[warp][1][lane * 4 + e] = acc1[e]; }
  __syncthreads();
  const int h = tid >> 7, d = tid & 127;
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < ATT_NW; ++w) M = fmaxf(M, s_m[w][h]);
  float A = 0.f, Ls = 0.f;
#pragma unroll
  for (int w = 0; w < ATT_NW; ++w) {
    const float f = (s_m[w][h] == -INFINITY) ? 0.f : __expf(s_m[w][h] - M);
    A = fmaf(s_acc[w][h][d], f, A);
    Ls = fmaf(s_l[w][h], f, Ls);
  }
  // this chunk's partial -> global memory, then one ticket per CTA: the last one merges
  float* mine = part + ((((size_t)b * NKVH + g) * Z + z) * 2 + h) * ATT_PART;
  __stcg(mine + 2 + d, A);
  if (d == 0) { __stcg(mine, M); __stcg(mine + 1, Ls); }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned t = atomicAdd(&tickets[b * NKVH + g], 1u);
    s_last = (t == (unsigned)(Z - 1)) ? 1u : 0u;
    if (t == (unsigned)(Z - 1)) tickets[b * NKVH + g] = 0u;   // every CTA of the item has drawn: ready for the next kernel
  }
  __syncthreads();
  if (s_last == 0u) return;
  __threadfence();
  float pm[Z], pl[Z], pa[Z];
#pragma unroll
  for (int zz = 0; zz < Z; ++zz) {
    const float* src = part + ((((size_t)b * NKVH + g) * Z + zz) * 2 + h) * ATT_PART;
    pm[zz] = __ldcg(src); pl[zz] = __ldcg(src + 1); pa[zz] = __ldcg(src + 2 + d);
  }
  float Mx = -INFINITY;
#pragma unroll
  for (int zz = 0; zz < Z; ++zz) Mx = fmaxf(Mx, pm[zz]);
  float At = 0.f, Lt = 0.f;
#pragma unroll
  for (int zz = 0; zz < Z; ++zz) {
    const float f = (pm[zz] == -INFINITY) ? 0.f : __expf(pm[zz] - Mx);
    At = fmaf(pa[zz], f, At);
    Lt = fmaf(pl[zz], f, Lt);
  }
  a_out[(size_t)b * QSZ + (2 * g + h) * HD + d] = __float2bfloat16_rn(At / Lt);
}

// ---- one-pass prefill through the launch chain: lane i = position p0 + i of ONE utterance, all lanes share a B = 1 cache
//      [L][8][max_seq][128].  The QKV epilogue is split in two kernels because lane i attends to the rows the other lanes write:
//      (1) finish: split-K sum -> bf16 -> per-head norm + RoPE at position p0 + lane; q -> qpre (fp32, bf16-exact), k / v rows -> cache;
//      (2) causal attention of lane i over rows 0 .. p0 + i (the kernel boundary orders it after every lane's rows). --------------
// grid = (n, 8), block = 128
template <int SPLITS>
__global__ void __launch_bounds__(128) kb_qkv_finish_prefill(const float* partial, int B, int p0, const __nv_bfloat16* q_norm, const __nv_bfloat16* k_norm,
                                                             const __nv_bfloat16* cos_t, const __nv_bfloat16* sin_t, __nv_bfloat16* k_cache,
                                                             __nv_bfloat16* v_cache, float* qpre, int layer, int max_seq) {
  const int b = blockIdx.x, g = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pos = p0 + b;
  const size_t base = ((size_t)layer * NKVH + g) * max_seq * HD;
  const __nv_bfloat16* wn = warp < 2 ? q_norm : k_norm;
  const uint2 wn_raw = warp < 3 ? *reinterpret_cast<const uint2*>(wn + lane * 4) : make_uint2(0u, 0u);
  const int dbase = (lane * 4) & 63;
  const uint2 cs_raw = *reinterpret_cast<const uint2*>(cos_t + (size_t)pos * HD + dbase), sn_raw = *reinterpret_cast<const uint2*>(sin_t + (size_t)pos * HD + dbase);
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  // warp 0, 1: q heads 2g, 2g+1; warp 2: k head g; warp 3: v head g
  const int row0 = warp < 2 ? (2 * g + warp) * HD : (warp == 2 ? QSZ + g * HD : QSZ + KVSZ + g * HD);
  float4 pp[SPLITS];
#pragma unroll
  for (int s = 0; s < SPLITS; ++s) pp[s] = __ldcg(reinterpret_cast<const float4*>(partial + ((size_t)s * B + b) * QKV_ROWS + row0 + lane * 4));
  float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int s = 0; s < SPLITS; ++s) { t[0] += pp[s].x; t[1] += pp[s].y; t[2] += pp[s].z; t[3] += pp[s].w; }
#pragma unroll
  for (int e = 0; e < 4; ++e) t[e] = bf16_round(t[e]);
  if (warp == 3) {
    *reinterpret_cast<uint2*>(v_cache + base + (size_t)pos * HD + lane * 4) =
        make_uint2(bf16_bits(t[0]) | (bf16_bits(t[1]) << 16), bf16_bits(t[2]) | (bf16_bits(t[3]) << 16));
    return;
  }
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) ss = fmaf(t[e], t[e], ss);
  ss = warp_sum(ss);
  const float rms = sqrtf(ss * (1.0f / HD) + EPS);
  const float wf[4] = {bf16_lo(wn_raw.x), bf16_hi(wn_raw.x), bf16_lo(wn_raw.y), bf16_hi(wn_raw.y)};
  const float cf[4] = {bf16_lo(cs_raw.x), bf16_hi(cs_raw.x), bf16_lo(cs_raw.y), bf16_hi(cs_raw.y)};
  const float sf[4] = {bf16_lo(sn_raw.x), bf16_hi(sn_raw.x), bf16_lo(sn_raw.y), bf16_hi(sn_raw.y)};
  float o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float nv = bf16_round((t[e] / rms) * wf[e]);
    const float other = __shfl_xor_sync(0xffffffffu, nv, 16);
    const float x = bf16_round(nv * cf[e]), y = bf16_round(other * sf[e]);
    o[e] = bf16_round(lane < 16 ? x - y : x + y);
  }
  if (warp < 2)
    *reinterpret_cast<float4*>(qpre + ((size_t)b * 16 + 2 * g + warp) * HD + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
  else
    *reinterpret_cast<uint2*>(k_cache + base + (size_t)pos * HD + lane * 4) =
        make_uint2(bf16_bits(o[0]) | (bf16_bits(o[1]) << 16), bf16_bits(o[2]) | (bf16_bits(o[3]) << 16));
}
// grid = (n, 8), block = 256
__global__ void __launch_bounds__(256) kb_attention_prefill(const float* qpre, int p0, const __nv_bfloat16* k_cache, const __nv_bfloat16* v_cache,
                                                            __nv_bfloat16* a_out, int layer, int max_seq, float scale) {
  __shared__ float s_acc[ATT_NW][2][HD];
  __shared__ float s_m[ATT_NW][2], s_l[ATT_NW][2];
  const int b = blockIdx.x, g = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = p0 + b + 1;   // causal: rows 0 .. p0 + b
  const size_t base = ((size_t)layer * NKVH + g) * max_seq * HD;
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  const float4 qa = __ldcg(reinterpret_cast<const float4*>(qpre + ((size_t)b * 16 + 2 * g) * HD + lane * 4));
  const float4 qb = __ldcg(reinterpret_cast<const float4*>(qpre + ((size_t)b * 16 + 2 * g + 1) * HD + lane * 4));
  const float q0[4] = {qa.x, qa.y, qa.z, qa.w}, q1[4] = {qb.x, qb.y, qb.z, qb.w};
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f, acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
  for (int p00 = warp; p00 < n; p00 += ATT_NW * ATT_PRE) {
    uint2 kk[ATT_PRE], vv[ATT_PRE];
#pragma unroll
    for (int i = 0; i < ATT_PRE; ++i) {
      const int p = p00 + ATT_NW * i;
      if (p < n) {
        kk[i] = __ldcg(reinterpret_cast<const uint2*>(k_cache + base + (size_t)p * HD + lane * 4));
        vv[i] = __ldcg(reinterpret_cast<const uint2*>(v_cache + base + (size_t)p * HD + lane * 4));
      }
    }
#pragma unroll
    for (int i = 0; i < ATT_PRE; ++i) {
      const int p = p00 + ATT_NW * i;
      if (p < n) {
        const float kf[4] = {bf16_lo(kk[i].x), bf16_hi(kk[i].x), bf16_lo(kk[i].y), bf16_hi(kk[i].y)};
        const float vf[4] = {bf16_lo(vv[i].x), bf16_hi(vv[i].x), bf16_lo(vv[i].y), bf16_hi(vv[i].y)};
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) { d0 = fmaf(q0[e], kf[e], d0); d1 = fmaf(q1[e], kf[e], d1); }
        d0 = warp_sum(d0) * scale;
        d1 = warp_sum(d1) * scale;
        const float nm0 = fmaxf(m0, d0), nm1 = fmaxf(m1, d1);
        const float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - nm0), c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - nm1);
        const float e0 = __expf(d0 - nm0), e1 = __expf(d1 - nm1);
        l0 = l0 * c0 + e0; l1 = l1 * c1 + e1;
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc0[e] = fmaf(e0, vf[e], acc0[e] * c0); acc1[e] = fmaf(e1, vf[e], acc1[e] * c1); }
        m0 = nm0; m1 = nm1;
      }
    }
  }
  if (lane == 0) { s_m[warp][0] = m0; s_m[warp][1] = m1; s_l[warp][0] = l0; s_l[warp][1] = l1; }
#pragma unroll
  for (int e = 0; e < 4; ++e) { s_acc[warp][0][lane * 4 + e] = acc0[e]; s_acc[warp][1][lane * 4 + e] = acc1[e]; }
  __syncthreads();
  const int h = tid >> 7, d = tid & 127;
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < ATT_NW; ++w) M = fmaxf(M, s_m[w][h]);
  float A = 0.f, Ls = 0.f;
#pragma unroll
  for (int w = 0; w < ATT_NW; ++w) {
    const float f = (s_m[w][h] == -INFINITY) ? 0.f : __expf(s_m[w][h] - M);
    A = fmaf(s_acc[w][h][d], f, A);
    Ls = fmaf(s_l[w][h], f, Ls);
  }
  a_out[(size_t)b * QSZ + (2 * g + h) * HD + d] = __float2bfloat16_rn(A / Ls);
}
__global__ void kb_lane_positions(int* pos, int p0, int B) { if (threadIdx.x < B) pos[threadIdx.x] = p0 + threadIdx.x; }

// ---- gate/up epilogue: m = r( r(silu(r(g))) * r(u) ) -------------------------------------------------------------------
// grid = (B, 3), block = 256; partial: [splits][B][6144] (gate rows 0..3071, up rows 3072..6143)
template <int SPLITS>
__global__ void __launch_bounds__(256) kb_gu_epilogue(const float* partial, int B, __nv_bfloat16* m_out) {
  qmkb::KTrace kt;
  kt.mark(0);
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  kt.mark(1);
  const int b = blockIdx.x, j = blockIdx.y * 1024 + threadIdx.x * 4;
  __shared__ float4 s_pg[SPLITS][256], s_pu[SPLITS][256];
#pragma unroll
  for (int s = 0; s < SPLITS; ++s) {
    const float* p = partial + ((size_t)s * B + b) * GU_ROWS;
    cp_async16(&s_pg[s][threadIdx.x], p + j);
    cp_async16(&s_pu[s][threadIdx.x], p + INTER + j);
  }
  cp_async_wait_all();
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f), u = g;
#pragma unroll
  for (int s = 0; s < SPLITS; ++s) {
    const float4 pg = s_pg[s][threadIdx.x], pu = s_pu[s][threadIdx.x];
    g.x += pg.x; g.y += pg.y; g.z += pg.z; g.w += pg.w;
    u.x += pu.x; u.y += pu.y; u.z += pu.z; u.w += pu.w;
  }
  const float gv[4] = {g.x, g.y, g.z, g.w}, uv[4] = {u.x, u.y, u.z, u.w};
  uint32_t o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float gg = bf16_round(gv[e]);
    const float sg = bf16_round(__fdividef(gg, 1.0f + __expf(-gg)));
    o[e] = bf16_bits(sg * bf16_round(uv[e]));
  }
  *reinterpret_cast<uint2*>(m_out + (size_t)b * INTER + j) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
  kt.flush(5);
}

// ---- LM head epilogue: bf16 logits; argmax (lowest index on ties) or temperature / top-k / multinomial (the B = 1 engine's
//      sampler, csrc/qmk_sample.cuh; stream b draws from the generator of a B = 1 engine seeded seed + b * 0x632BE59BD9B4E019);
//      advances the stream's position ------------------------------------------------------------------------------------
// grid = B, block = 256; partial: [splits][B][rows]; codes_out (optional): int64, element b * codes_stride + codes_col
struct HeadSelect {
  int do_sample, top_k, group;
  float temperature;
  unsigned long long seed, counter;
  const unsigned long long* counter_ptr;
  long long* codes_out;
  int codes_stride, codes_col;
};
__global__ void kb_head_epilogue(const float* partial, int splits, int B, int rows, int* tokens_out, int* positions, HeadSelect sel) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  __shared__ float s_log[2048];
  __shared__ unsigned s_hist[512];
  __shared__ float s_red[64];
  extern __shared__ __align__(16) float s_stage[];   // [splits][rows]: the stream's partial logits (dynamic: splits * rows * 4 bytes)
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool sample = sel.do_sample && rows <= 2048 && rows % 256 == 0;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int s = 0; s < splits; ++s)
    for (int c = tid * 4; c < rows; c += 1024) cp_async16(s_stage + (size_t)s * rows + c, partial + ((size_t)s * B + b) * rows + c);
  cp_async_wait_all();
  for (int c = tid * 4; c < rows; c += 1024) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 pv = *reinterpret_cast<const float4*>(s_stage + (size_t)s * rows + c);
      acc.x += pv.x; acc.y += pv.y; acc.z += pv.z; acc.w += pv.w;
    }
    const float v4[4] = {bf16_round(acc.x), bf16_round(acc.y), bf16_round(acc.z), bf16_round(acc.w)};
    if (sample) *reinterpret_cast<float4*>(s_log + c) = make_float4(v4[0], v4[1], v4[2], v4[3]);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (v4[e] > best) { best = v4[e]; best_i = c + e; }   // ascending indices per thread: the lowest index of a tie is kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if (lane == 0) { s_v[warp] = best; s_i[warp] = best_i; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (s_v[w] > best || (s_v[w] == best && s_i[w] < best_i)) { best = s_v[w]; best_i = s_i[w]; }
  if (best_i == 0x7fffffff) best_i = 0;
  int chosen = best_i;
  if (sample)
    chosen = qmk2::sample_token2(s_log, s_hist, s_red, tid, warp, lane, rows, sel.top_k, sel.temperature,
                                 sel.seed + (unsigned long long)b * 0x632BE59BD9B4E019ull,
                                 sel.counter + (sel.counter_ptr != nullptr ? *sel.counter_ptr : 0ull), sel.group, best, best_i);
  if (tid == 0) {
    tokens_out[b] = chosen;
    if (sel.codes_out != nullptr) sel.codes_out[(size_t)b * sel.codes_stride + sel.codes_col] = (long long)chosen;
    positions[b] += 1;
  }
}

// ---- frame loop glue for B streams (upstream tts_engine.py:319-333 per stream): e[b] = talker_embed[codes[b][0]] +
//      sum_g group_table[g][codes[b][g + 1]] + extra[b]   (bf16 adds in the upstream order); grid = B, block = 256 -------------
struct GroupTables { const __nv_bfloat16* t[15]; };
__global__ void kb_embed_sum(const long long* codes, const __nv_bfloat16* talker_embed, int rows0, GroupTables gt, int rows,
                             const __nv_bfloat16* extra, int extra_stride, __nv_bfloat16* out) {
  const int b = blockIdx.x, t = threadIdx.x;
  int code[16];
#pragma unroll
  for (int g = 0; g < 16; ++g) {
    const int v = (int)codes[(size_t)b * 16 + g], hi = (g == 0 ? rows0 : rows) - 1;
    code[g] = v < 0 ? 0 : (v > hi ? hi : v);
  }
  uint2 row[16];
  row[0] = *reinterpret_cast<const uint2*>(talker_embed + (size_t)code[0] * H + t * 4);
#pragma unroll
  for (int g = 0; g < 15; ++g) row[g + 1] = *reinterpret_cast<const uint2*>(gt.t[g] + (size_t)code[g + 1] * H + t * 4);
  const uint2 vx = *reinterpret_cast<const uint2*>(extra + (size_t)b * extra_stride + t * 4);
  float e4[4] = {bf16_lo(row[0].x), bf16_hi(row[0].x), bf16_lo(row[0].y), bf16_hi(row[0].y)};
#pragma unroll
  for (int g = 1; g < 16; ++g) {
    e4[0] = bf16_round(e4[0] + bf16_lo(row[g].x)); e4[1] = bf16_round(e4[1] + bf16_hi(row[g].x));
    e4[2] = bf16_round(e4[2] + bf16_lo(row[g].y)); e4[3] = bf16_round(e4[3] + bf16_hi(row[g].y));
  }
  const uint32_t o0 = bf16_bits(e4[0] + bf16_lo(vx.x)) | (bf16_bits(e4[1] + bf16_hi(vx.x)) << 16);
  const uint32_t o1 = bf16_bits(e4[2] + bf16_lo(vx.y)) | (bf16_bits(e4[3] + bf16_hi(vx.y)) << 16);
  *reinterpret_cast<uint2*>(out + (size_t)b * H + t * 4) = make_uint2(o0, o1);
}

__global__ void kb_counter_add(unsigned long long* counter, unsigned long long inc) { *counter += inc; }

// concatenate row blocks of two / three [rows, K] matrices into one (one-time, at model creation)
__global__ void kb_copy_rows(const uint4* src, uint4* dst, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// [rows][K] (upstream layout) -> k-block-major [K / 64][rows_total][64] at row offset row_off: one 128-row x 64-k tile becomes
// 16 KB of contiguous memory (what the persistent kernel's TMA requests).  One uint4 = 8 bf16.
__global__ void kb_repack_kmajor(const uint4* src, uint4* dst, int rows, int K, int row_off, int rows_total) {
  const size_t n = (size_t)rows * (K / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (K / 8)), ch = (int)(i % (K / 8));
    dst[((size_t)(ch >> 3) * rows_total + row_off + r) * 8 + (ch & 7)] = src[i];
  }
}

thread_local std::string g_err;
int fail(int code, const char* msg) {
  g_err = msg;
  return code;
}

}  // namespace

constexpr int MAX_HEAD_ROWS_B = 8192;   // LM-head rows the head epilogue can stage in shared memory (4 splits x rows x 4 bytes = 128 KB)
constexpr size_t PARTIAL_ELEMS = (size_t)2 * 1024 * 1024;   // fp32 split-K partials: max splits x B x rows = 4 x 64 x 6144
constexpr int MAX_LANES = 64;
static size_t qmkb_hidden_offset() { return PARTIAL_ELEMS; }   // prefill: per-lane hidden states live behind the partials

struct BatchedDeviceGuard {
  int prev = -1;
  explicit BatchedDeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~BatchedDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct qmk_batched {
  int device = 0, L = 0, B = 0, max_seq = 0, head_rows = 0, residual_fp32 = 1;
  __nv_bfloat16 *w_qkv = nullptr, *w_gu = nullptr;           // [L][4096][1024], [L][6144][1024] (concatenated copies)
  // persistent kernel: k-block-major copies [L][K / 64][rows][64] of qkv, o, gate/up, down and [K / 64][rows][64] of the head
  __nv_bfloat16 *p_qkv = nullptr, *p_o = nullptr, *p_gu = nullptr, *p_down = nullptr, *p_head = nullptr;
  CUtensorMap pmap_w[5], pmap_x[3];
  std::vector<const void*> w_o, w_down, ln_in, ln_post, qn, kn;
  const void *final_norm = nullptr, *lm_head = nullptr, *embed = nullptr, *cos_t = nullptr, *sin_t = nullptr;
  std::vector<CUtensorMap> map_qkv, map_o, map_gu, map_down;
  CUtensorMap map_head, map_x1024, map_x2048, map_x3072;
  std::vector<CUtensorMap> extra_head_maps;                  // heads registered with qmk_batched_add_head (index 1, 2, ...)
  std::vector<int> extra_head_rows;
  float *res = nullptr, *partial = nullptr;
  __nv_bfloat16 *xn = nullptr, *abuf = nullptr, *mbuf = nullptr;
  // persistent step kernel (qmk_bstep.cuh)
  int persistent = 0, decode_persistent = 0, grid = 0, N = 0;   // persistent: kernel usable (prefill); decode_persistent: also used for decode steps
  const __nv_bfloat16** d_ptrs = nullptr; // [4][L]: ln_in, ln_post, qn, kn
  float* qbuf = nullptr;
  unsigned* d_bar = nullptr;
  unsigned bar_count = 0;
  int* d_status = nullptr;
  int* d_pos0 = nullptr;                  // prefill: first position
  int prefill_persistent = 0;            // QMK_PREFILL_PERSISTENT=1: the prefill as one persistent launch (csrc/qmk_bstep.cuh) instead of the chain
  float* att_part = nullptr;             // split attention: f32[B][8][4][2][130] chunk partials
  unsigned* att_tickets = nullptr;       // split attention: u32[B][8], zero between kernels
  float* qpre = nullptr;                 // chain prefill: q of every lane, f32[B][16][128]
  int* pos_scratch = nullptr;            // chain prefill: int[B] lane positions
  int* tok_scratch = nullptr;            // chain prefill: int[B] argmax tokens
  int splits_od = 16;                    // K slices of the O / down projections (8 measured: 741 / 892 us per step instead of 660 / 866)
  long long* d_trace = nullptr;           // QMK_BATCHED_TRACE=1: barrier stamps of CTA 0 (debug)
  int* d_tok_scratch = nullptr;           // prefill: per-lane argmax (only the last lane's is reported)
};

extern "C" const char* qmk_batched_last_error(void) { return g_err.c_str(); }

extern "C" int qmk_batched_create(int device, const LDGLayerWeights* layers_host, int num_layers, const void* final_norm_weight,
                                  const void* lm_head_weight, int lm_head_rows, const void* embed_weight,
                                  const void* cos_table, const void* sin_table, int residual_fp32, int batch,
                                  int max_seq_len, qmk_batched** out) {
  using namespace qmkb;
  if (!layers_host || !final_norm_weight || !lm_head_weight || !embed_weight || !cos_table || !sin_table || !out)
    return fail(QMK_ERR_ARG, "qmk_batched_create: null argument");
  if (batch < 16 || batch > MAX_N || batch % 16) return fail(QMK_ERR_ARG, "qmk_batched_create: batch must be 16, 32, 48 or 64");
  if (lm_head_rows % BM || lm_head_rows <= 0) return fail(QMK_ERR_ARG, "qmk_batched_create: lm_head_rows must be a multiple of 128");
  if (num_layers < 1 || max_seq_len < 1) return fail(QMK_ERR_ARG, "qmk_batched_create: bad num_layers / max_seq_len");
  if (lm_head_rows > MAX_HEAD_ROWS_B) return fail(QMK_ERR_ARG, "qmk_batched_create: lm_head_rows above 8192");
  if ((size_t)6 * batch * lm_head_rows > PARTIAL_ELEMS) return fail(QMK_ERR_ARG, "qmk_batched_create: lm_head_rows too large for the split-K partial buffer");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(QMK_ERR_ARG, "qmk_batched_create: bad device");
  BatchedDeviceGuard guard(device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) return fail(QMK_ERR_UNSUPPORTED, "qmk_batched_create: tcgen05 needs an sm_100 device");
  qmk_batched* h = new qmk_batched();
  h->device = device; h->L = num_layers; h->B = batch; h->max_seq = max_seq_len; h->head_rows = lm_head_rows;
  if (const char* env = getenv("QMK_PREFILL_PERSISTENT")) h->prefill_persistent = atoi(env) != 0;
  h->residual_fp32 = residual_fp32 ? 1 : 0;
  h->final_norm = final_norm_weight; h->lm_head = lm_head_weight; h->embed = embed_weight; h->cos_t = cos_table; h->sin_t = sin_table;
  const size_t L = num_layers;
  bool ok = cudaMalloc(&h->w_qkv, L * QKV_ROWS * H * 2) == cudaSuccess && cudaMalloc(&h->w_gu, L * GU_ROWS * H * 2) == cudaSuccess &&
            cudaMalloc(&h->res, (size_t)batch * H * 4) == cudaSuccess && cudaMalloc(&h->partial, (PARTIAL_ELEMS + (size_t)MAX_LANES * H) * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&h->xn, (size_t)batch * H * 2) == cudaSuccess &&
            cudaMalloc(&h->abuf, (size_t)batch * QSZ * 2) == cudaSuccess && cudaMalloc(&h->mbuf, (size_t)batch * INTER * 2) == cudaSuccess &&
            cudaMemset(h->xn, 0, (size_t)batch * H * 2) == cudaSuccess && cudaMemset(h->abuf, 0, (size_t)batch * QSZ * 2) == cudaSuccess &&
            cudaMemset(h->mbuf, 0, (size_t)batch * INTER * 2) == cudaSuccess &&   // rows of unused lanes feed the tensor cores too: keep them finite
            cudaMalloc(&h->qpre, (size_t)batch * QSZ * 4) == cudaSuccess && cudaMalloc(&h->pos_scratch, (size_t)batch * sizeof(int)) == cudaSuccess &&
            cudaMalloc(&h->tok_scratch, (size_t)batch * sizeof(int)) == cudaSuccess &&
            cudaMalloc(&h->att_part, (size_t)batch * NKVH * 4 * 2 * ATT_PART * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&h->att_tickets, (size_t)batch * NKVH * sizeof(unsigned)) == cudaSuccess &&
            cudaMemset(h->att_tickets, 0, (size_t)batch * NKVH * sizeof(unsigned)) == cudaSuccess;
  if (!ok) { qmk_batched_destroy(h); return fail(QMK_ERR_CUDA, "qmk_batched_create: allocation failed"); }
  for (int l = 0; l < num_layers; ++l) {
    const LDGLayerWeights& w = layers_host[l];
    auto cp = [&](const void* src, __nv_bfloat16* dst, size_t rows) {
      kb_copy_rows<<<296, 256>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), rows * H * 2 / 16);
    };
    cp(w.q_proj_weight, h->w_qkv + ((size_t)l * QKV_ROWS) * H, QSZ);
    cp(w.k_proj_weight, h->w_qkv + ((size_t)l * QKV_ROWS + QSZ) * H, KVSZ);
    cp(w.v_proj_weight, h->w_qkv + ((size_t)l * QKV_ROWS + QSZ + KVSZ) * H, KVSZ);
    cp(w.gate_proj_weight, h->w_gu + ((size_t)l * GU_ROWS) * H, INTER);
    cp(w.up_proj_weight, h->w_gu + ((size_t)l * GU_ROWS + INTER) * H, INTER);
    h->w_o.push_back(w.o_proj_weight); h->w_down.push_back(w.down_proj_weight);
    h->ln_in.push_back(w.input_layernorm_weight); h->ln_post.push_back(w.post_attn_layernorm_weight);
    h->qn.push_back(w.q_norm_weight); h->kn.push_back(w.k_norm_weight);
    CUtensorMap m;
    int rc = make_tensor_map(&m, h->w_qkv + (size_t)l * QKV_ROWS * H, QKV_ROWS, H, BM); h->map_qkv.push_back(m);
    rc |= make_tensor_map(&m, w.o_proj_weight, H, QSZ, BM); h->map_o.push_back(m);
    rc |= make_tensor_map(&m, h->w_gu + (size_t)l * GU_ROWS * H, GU_ROWS, H, BM); h->map_gu.push_back(m);
    rc |= make_tensor_map(&m, w.down_proj_weight, H, INTER, BM); h->map_down.push_back(m);
    if (rc) { qmk_batched_destroy(h); return fail(QMK_ERR_CUDA, "qmk_batched_create: cuTensorMapEncodeTiled failed"); }
  }
  int rc = make_tensor_map(&h->map_head, lm_head_weight, lm_head_rows, H, BM);
  rc |= make_tensor_map(&h->map_x1024, h->xn, batch, H, batch);
  rc |= make_tensor_map(&h->map_x2048, h->abuf, batch, QSZ, batch);
  rc |= make_tensor_map(&h->map_x3072, h->mbuf, batch, INTER, batch);
  if (rc) { qmk_batched_destroy(h); return fail(QMK_ERR_CUDA, "qmk_batched_create: cuTensorMapEncodeTiled failed"); }
  // ---- persistent step kernel: one cooperative launch per step (default when every projection's ~144 items fit the SMs) ----
  h->N = batch;
  h->grid = prop.multiProcessorCount;
  {
    // Decode steps run as the chain of per-projection launches by default: measured on B200 (round 2) the persistent kernel
    // needs 1.4 ms (B = 16) / 2.1 ms (B = 64) per step against the chain's 1.0 / 1.1 ms -- nine grid barriers per layer at ~2 us
    // each (fence.proxy.async + release/acquire round trip) and one CTA per SM walking its items serially cost more than the
    // chain's kernel boundaries, which overlap through programmatic dependent launch and several CTAs per SM.
    // QMK_BATCHED_PERSISTENT=1 selects it for decode; the one-pass prefill always uses it.
    int want = 1;
    if (const char* env = getenv("QMK_BATCHED_PERSISTENT")) h->decode_persistent = atoi(env) ? 1 : 0;
    int coop = 0, occ = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    cudaError_t pe = cudaFuncSetAttribute(qmk_bstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM);
    if (pe == cudaSuccess) pe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qmk_bstep_kernel, NT_ALL, PS_SMEM);
    if (pe != cudaSuccess) cudaGetLastError();
    h->persistent = (want && coop && pe == cudaSuccess && occ >= 1 && h->grid >= 144 && lm_head_rows / BM * 6 <= h->grid) ? 1 : 0;
    if (!h->persistent) h->decode_persistent = 0;
  }
  if (h->persistent) {
    std::vector<const void*> ptrs;
    for (auto* v : {&h->ln_in, &h->ln_post, &h->qn, &h->kn}) ptrs.insert(ptrs.end(), v->begin(), v->end());
    bool okp = cudaMalloc(&h->p_qkv, L * QKV_ROWS * H * 2) == cudaSuccess && cudaMalloc(&h->p_o, L * H * QSZ * 2) == cudaSuccess &&
               cudaMalloc(&h->p_gu, L * GU_ROWS * H * 2) == cudaSuccess && cudaMalloc(&h->p_down, L * H * INTER * 2) == cudaSuccess &&
               cudaMalloc(&h->p_head, (size_t)lm_head_rows * H * 2) == cudaSuccess &&
               cudaMalloc(&h->d_ptrs, ptrs.size() * sizeof(void*)) == cudaSuccess &&
               cudaMemcpy(h->d_ptrs, ptrs.data(), ptrs.size() * sizeof(void*), cudaMemcpyHostToDevice) == cudaSuccess &&
               cudaMalloc(&h->qbuf, (size_t)batch * QSZ * 4) == cudaSuccess && cudaMalloc(&h->d_bar, sizeof(unsigned)) == cudaSuccess &&
               cudaMemset(h->d_bar, 0, sizeof(unsigned)) == cudaSuccess && cudaMalloc(&h->d_status, sizeof(int)) == cudaSuccess &&
               cudaMemset(h->d_status, 0, sizeof(int)) == cudaSuccess && cudaMalloc(&h->d_pos0, sizeof(int)) == cudaSuccess &&
               cudaMalloc(&h->d_tok_scratch, (size_t)batch * sizeof(int)) == cudaSuccess;
    if (okp) {
      auto pack = [&](const void* src, __nv_bfloat16* dst, int rows, int K, int row_off, int rows_total) {
        kb_repack_kmajor<<<592, 256>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), rows, K, row_off, rows_total);
      };
      for (int l = 0; l < num_layers; ++l) {
        const LDGLayerWeights& w = layers_host[l];
        pack(w.q_proj_weight, h->p_qkv + (size_t)l * QKV_ROWS * H, QSZ, H, 0, QKV_ROWS);
        pack(w.k_proj_weight, h->p_qkv + (size_t)l * QKV_ROWS * H, KVSZ, H, QSZ, QKV_ROWS);
        pack(w.v_proj_weight, h->p_qkv + (size_t)l * QKV_ROWS * H, KVSZ, H, QSZ + KVSZ, QKV_ROWS);
        pack(w.o_proj_weight, h->p_o + (size_t)l * H * QSZ, H, QSZ, 0, H);
        pack(w.gate_proj_weight, h->p_gu + (size_t)l * GU_ROWS * H, INTER, H, 0, GU_ROWS);
        pack(w.up_proj_weight, h->p_gu + (size_t)l * GU_ROWS * H, INTER, H, INTER, GU_ROWS);
        pack(w.down_proj_weight, h->p_down + (size_t)l * H * INTER, H, INTER, 0, H);
      }
      pack(lm_head_weight, h->p_head, lm_head_rows, H, 0, lm_head_rows);
      // 2-D views [L * KB * rows, 64]: tile (layer, k-block, row tile) = rows (l * KB + kb) * rows + 128 t .. + 127, all 64 columns
      int prc = make_tensor_map(&h->pmap_w[0], h->p_qkv, L * (H / BK) * QKV_ROWS, BK, BM);
      prc |= make_tensor_map(&h->pmap_w[1], h->p_o, L * (QSZ / BK) * H, BK, BM);
      prc |= make_tensor_map(&h->pmap_w[2], h->p_gu, L * (H / BK) * GU_ROWS, BK, BM);
      prc |= make_tensor_map(&h->pmap_w[3], h->p_down, L * (INTER / BK) * H, BK, BM);
      prc |= make_tensor_map(&h->pmap_w[4], h->p_head, (uint64_t)(H / BK) * lm_head_rows, BK, BM);
      h->pmap_x[0] = h->map_x1024; h->pmap_x[1] = h->map_x2048; h->pmap_x[2] = h->map_x3072;
      okp = prc == 0 && cudaDeviceSynchronize() == cudaSuccess;
    }
    if (!okp) { qmk_batched_destroy(h); return fail(QMK_ERR_CUDA, "qmk_batched_create: allocation failed (persistent kernel)"); }
  }
  if (cudaFuncSetAttribute(qmk_bgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(kb_head_epilogue, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * MAX_HEAD_ROWS_B * (int)sizeof(float)) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    qmk_batched_destroy(h);
    return fail(QMK_ERR_CUDA, "qmk_batched_create: kernel setup failed");
  }
  *out = h;
  return QMK_OK;
}

extern "C" void qmk_batched_destroy(qmk_batched* h) {
  if (!h) return;
  BatchedDeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  cudaFree(h->w_qkv); cudaFree(h->w_gu); cudaFree(h->res); cudaFree(h->partial); cudaFree(h->qpre); cudaFree(h->pos_scratch); cudaFree(h->tok_scratch); cudaFree(h->att_part); cudaFree(h->att_tickets);
  cudaFree(h->xn); cudaFree(h->abuf); cudaFree(h->mbuf);
  cudaFree(h->p_qkv); cudaFree(h->p_o); cudaFree(h->p_gu); cudaFree(h->p_down); cudaFree(h->p_head); cudaFree(h->d_ptrs); cudaFree(h->qbuf); cudaFree(h->d_bar); cudaFree(h->d_status); cudaFree(h->d_pos0);
  cudaFree(h->d_tok_scratch); cudaFree(h->d_trace);
  delete h;
}

// Every kernel of the step chain is launched with programmatic stream serialization (PDL).
static thread_local cudaError_t g_launch_err = cudaSuccess;   // first failed launch of the current step
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = e;
}
template <typename... KArgs, typename... Args>
static void launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, unsigned cluster_x, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cluster_x; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = e;
}
// residual + RMSNorm after an O / down projection: a cluster of 4 CTAs per stream (2 beyond 32 streams: the grid stays <= 128 CTAs)
static void resid_norm(qmk_batched* h, cudaStream_t st, const void* w_norm, float* hidden_out, int* advance_positions) {
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(w_norm);
  if (h->splits_od == 8)
    launch_pdl_cluster(kb_resid_norm<8, 2>, dim3(h->B * 2), dim3(128), 2, st, (const float*)h->partial, h->B, h->res, h->residual_fp32, w, h->xn, hidden_out, advance_positions);
  else if (h->B <= 32)
    launch_pdl_cluster(kb_resid_norm<16, 4>, dim3(h->B * 4), dim3(64), 4, st, (const float*)h->partial, h->B, h->res, h->residual_fp32, w, h->xn, hidden_out, advance_positions);
  else
    launch_pdl_cluster(kb_resid_norm<16, 2>, dim3(h->B * 2), dim3(128), 2, st, (const float*)h->partial, h->B, h->res, h->residual_fp32, w, h->xn, hidden_out, advance_positions);
}
static void gemm(qmk_batched* h, const CUtensorMap& mw, const CUtensorMap& mx, int M, int K, int splits, cudaStream_t st) {
  qmkb::BgemmArgs a{h->partial, M, h->B, K, splits};
  launch_pdl(qmkb::qmk_bgemm_kernel, dim3(M / qmkb::BM, splits), dim3(128), (size_t)qmkb::smem_bytes_for(K / splits / qmkb::BK), st, mw, mx, a);
}

// One cooperative launch of the persistent step kernel for `lanes` lanes (decode: lanes = B utterances; prefill: lanes =
// consecutive positions of one utterance whose first position is *positions).
static int launch_persistent(qmk_batched* h, int lanes, int prefill, const int32_t* token_ids, const void* embeds, int32_t* positions,
                             void* k_cache, void* v_cache, float* hidden_out, int32_t* tokens_out, cudaStream_t st) {
  qmkb::BStepParams p;
  memset(&p, 0, sizeof(p));
  memcpy(p.map_w, h->pmap_w, sizeof(p.map_w));
  memcpy(p.map_x, h->pmap_x, sizeof(p.map_x));
  p.L = h->L; p.B = lanes; p.N = h->N;
  p.head_rows = h->head_rows; p.vocab = h->head_rows;
  p.residual_fp32 = h->residual_fp32;
  p.prefill = prefill;
  p.max_seq = h->max_seq;
  p.attn_scale = 0.08838834764831845f;
  p.ln_in = reinterpret_cast<const __nv_bfloat16* const*>(h->d_ptrs);
  p.ln_post = p.ln_in + h->L; p.qn = p.ln_in + 2 * h->L; p.kn = p.ln_in + 3 * h->L;
  p.final_norm = reinterpret_cast<const __nv_bfloat16*>(h->final_norm);
  p.embed = reinterpret_cast<const __nv_bfloat16*>(h->embed);
  p.cos_t = reinterpret_cast<const __nv_bfloat16*>(h->cos_t);
  p.sin_t = reinterpret_cast<const __nv_bfloat16*>(h->sin_t);
  p.token_ids = token_ids;
  p.embeds = reinterpret_cast<const __nv_bfloat16*>(embeds);
  p.positions = positions;
  p.k_cache = reinterpret_cast<__nv_bfloat16*>(k_cache);
  p.v_cache = reinterpret_cast<__nv_bfloat16*>(v_cache);
  p.hidden_out = hidden_out;
  p.tokens_out = tokens_out;
  p.res = h->res; p.partial = h->partial; p.qbuf = h->qbuf; p.xn = h->xn; p.abuf = h->abuf; p.mbuf = h->mbuf;
  p.bar = h->d_bar;
  p.bar_base = h->bar_count;
  p.status = h->d_status;
  p.timeout_cycles = 4000000000LL;
  if (const char* env = getenv("QMK_TIMEOUT_CYCLES")) p.timeout_cycles = atoll(env);
  if (getenv("QMK_BATCHED_TRACE")) {
    if (!h->d_trace) cudaMalloc(&h->d_trace, 1024 * sizeof(long long));
    p.trace = h->d_trace;
  }
  const unsigned n_barriers = 1u + (unsigned)h->L * (8u + (prefill ? 1u : 0u)) + 1u;
  h->bar_count += n_barriers * (unsigned)h->grid;
  void* args[] = {&p};
  const cudaError_t e = cudaLaunchCooperativeKernel((const void*)qmkb::qmk_bstep_kernel, dim3(h->grid), dim3(qmkb::NT_ALL), args, (size_t)qmkb::PS_SMEM, st);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(QMK_ERR_CUDA, cudaGetErrorString(e)); }
  return QMK_OK;
}

// One decode step for all B streams (the chain of per-projection launches).
static int chain_step(qmk_batched* h, const qmk_batched_step_args* a, cudaStream_t st) {
  const int B = h->B, L = h->L;
  const float scale = 0.08838834764831845f;
  __nv_bfloat16* kc = reinterpret_cast<__nv_bfloat16*>(a->k_cache);
  __nv_bfloat16* vc = reinterpret_cast<__nv_bfloat16*>(a->v_cache);
  const __nv_bfloat16* cos_t = reinterpret_cast<const __nv_bfloat16*>(h->cos_t);
  const __nv_bfloat16* sin_t = reinterpret_cast<const __nv_bfloat16*>(h->sin_t);
  const __nv_bfloat16* table = reinterpret_cast<const __nv_bfloat16*>(a->token_table ? a->token_table : h->embed);
  const int table_rows = a->token_table ? a->table_rows : h->head_rows;
  int head_rows = 0;
  const CUtensorMap* head_map = nullptr;
  if (a->head == 0) { head_rows = h->head_rows; head_map = &h->map_head; }
  else if (a->head > 0) { head_rows = h->extra_head_rows[a->head - 1]; head_map = &h->extra_head_maps[a->head - 1]; }
  launch_pdl(kb_input, dim3(B), dim3(256), 0, st, (const int*)a->token_ids, table, table_rows,
             reinterpret_cast<const __nv_bfloat16*>(a->embeds_bf16), (const float*)a->embeds_f32, h->res,
             reinterpret_cast<const __nv_bfloat16*>(h->ln_in[0]), h->xn, B);
  for (int l = 0; l < L; ++l) {
    gemm(h, h->map_qkv[l], h->map_x1024, QKV_ROWS, H, 4, st);                       // 32 tiles x 4 K-slices
    {
      const __nv_bfloat16* qn = reinterpret_cast<const __nv_bfloat16*>(h->qn[l]);
      const __nv_bfloat16* kn = reinterpret_cast<const __nv_bfloat16*>(h->kn[l]);
      const float* pp = h->partial;
      const int* pos = (const int*)a->positions;
      // Deep contexts: the context of an item is split over several CTAs (the caller's hint; positions live on the device).
      // Measured, ms per step at p = 500 / 1000 / 2000 with 1 / 2 / 4 / 8 CTAs per item: B = 16: 1.38 2.06 3.40 / 1.19 1.60 2.40 /
      // 1.10 1.40 2.02 / 1.17 1.40 1.91; B = 64: 2.07 3.18 5.35 / 2.07 3.07 5.01 / 2.21 3.19 5.08 / 2.46 3.41 5.27.
      const int zsplit = (B == 16 && a->depth_hint >= ATT_SPLIT_MIN) ? 4
                         : ((B == 32 && a->depth_hint >= ATT_SPLIT_MIN) || (B >= 48 && a->depth_hint >= ATT_SPLIT_MIN_WIDE)) ? 2 : 1;
      if (zsplit == 4)
        launch_pdl(kb_qkv_attention_split<4, 4>, dim3(B, NKVH, 4), dim3(256), 0, st, pp, B, pos, qn, kn, cos_t, sin_t, kc, vc, h->abuf, h->att_part,
                   h->att_tickets, l, L, h->max_seq, scale);
      else if (zsplit == 2)
        launch_pdl(kb_qkv_attention_split<4, 2>, dim3(B, NKVH, 2), dim3(256), 0, st, pp, B, pos, qn, kn, cos_t, sin_t, kc, vc, h->abuf, h->att_part,
                   h->att_tickets, l, L, h->max_seq, scale);
      else
        launch_pdl(kb_qkv_attention<4>, dim3(B, NKVH), dim3(32 * ATT_NW), 0, st, pp, B, pos, qn, kn, cos_t, sin_t, kc, vc, h->abuf, l, L, h->max_seq, scale);
    }
    gemm(h, h->map_o[l], h->map_x2048, H, QSZ, h->splits_od, st);                    // 8 tiles x 16 K-slices
    resid_norm(h, st, h->ln_post[l], nullptr, nullptr);
    gemm(h, h->map_gu[l], h->map_x1024, GU_ROWS, H, 4, st);                          // 48 tiles x 4 K-slices
    launch_pdl(kb_gu_epilogue<4>, dim3(B, 3), dim3(256), 0, st, (const float*)h->partial, B, h->mbuf);
    gemm(h, h->map_down[l], h->map_x3072, H, INTER, h->splits_od, st);               // 8 tiles x 16 K-slices
    const bool last = (l == L - 1);
    resid_norm(h, st, last ? h->final_norm : h->ln_in[l + 1], last ? a->hidden_out : (float*)nullptr,
               (last && head_map == nullptr) ? (int*)a->positions : (int*)nullptr);
  }
  if (head_map != nullptr) {
    gemm(h, *head_map, h->map_x1024, head_rows, H, 4, st);
    HeadSelect sel;
    sel.do_sample = (a->do_sample && a->temperature > 0.f) ? 1 : 0;
    sel.top_k = a->top_k; sel.group = a->group; sel.temperature = sel.do_sample ? a->temperature : 1.0f;
    sel.seed = a->seed; sel.counter = a->counter; sel.counter_ptr = reinterpret_cast<const unsigned long long*>(a->counter_ptr);
    sel.codes_out = reinterpret_cast<long long*>(a->codes_out); sel.codes_stride = a->codes_stride; sel.codes_col = a->codes_col;
    launch_pdl(kb_head_epilogue, dim3(B), dim3(256), (size_t)4 * head_rows * sizeof(float), st, (const float*)h->partial, 4, B, head_rows, (int*)a->tokens_out,
               (int*)a->positions, sel);
  }
  cudaError_t e = g_launch_err != cudaSuccess ? g_launch_err : cudaGetLastError();
  if (e != cudaSuccess) { cudaGetLastError(); return fail(QMK_ERR_CUDA, cudaGetErrorString(e)); }
  return QMK_OK;
}

// Register a further LM head (bf16 [rows, 1024], rows % 128 == 0): returns its index (1, 2, ...; 0 is the create-time head).
extern "C" int qmk_batched_add_head(qmk_batched* h, const void* lm_head_weight, int rows) {
  if (!h || !lm_head_weight) return fail(QMK_ERR_ARG, "qmk_batched_add_head: null argument");
  if (rows <= 0 || rows % qmkb::BM || rows > MAX_HEAD_ROWS_B || (size_t)6 * h->B * rows > PARTIAL_ELEMS) return fail(QMK_ERR_ARG, "qmk_batched_add_head: bad row count");
  CUtensorMap m;
  if (qmkb::make_tensor_map(&m, lm_head_weight, rows, H, qmkb::BM)) return fail(QMK_ERR_CUDA, "qmk_batched_add_head: cuTensorMapEncodeTiled failed");
  h->extra_head_maps.push_back(m);
  h->extra_head_rows.push_back(rows);
  return (int)h->extra_head_maps.size();
}

// General step (the code predictor's steps need more than qmk_batched_step offers: an input table per step, fp32 inputs, a head
// index or none, sampling, int64 code outputs).  Always the chain of launches.
extern "C" int qmk_batched_step_ex(qmk_batched* h, const qmk_batched_step_args* a, void* stream) {
  if (!h || !a || !a->positions || !a->k_cache || !a->v_cache) return fail(QMK_ERR_ARG, "qmk_batched_step_ex: null argument");
  if (!a->token_ids && !a->embeds_bf16 && !a->embeds_f32) return fail(QMK_ERR_ARG, "qmk_batched_step_ex: no input given");
  if (a->head > (int)h->extra_head_maps.size()) return fail(QMK_ERR_ARG, "qmk_batched_step_ex: head index not registered");
  if (a->head >= 0 && !a->tokens_out) return fail(QMK_ERR_ARG, "qmk_batched_step_ex: tokens_out is null");
  if (a->token_table && a->table_rows < 1) return fail(QMK_ERR_ARG, "qmk_batched_step_ex: table_rows must be positive");
  BatchedDeviceGuard guard(h->device);
  g_launch_err = cudaSuccess;
  return chain_step(h, a, (cudaStream_t)stream);
}

// 16-way embedding sum of the frame loop for B streams: out[b] = talker_embed[codes[b][0]] + sum_g tables[g][codes[b][g+1]] +
// extra[b * extra_stride ..] (extra_stride = 0: one vector for all streams).  group_tables: HOST array of 15 device pointers.
extern "C" int qmk_batched_embed_sum(int batch, const int64_t* codes, const void* talker_embed, int talker_rows,
                                     const void* const* group_tables, int group_rows, const void* extra_bf16, int extra_stride,
                                     void* out_bf16, void* stream) {
  if (batch < 1 || !codes || !talker_embed || !group_tables || !extra_bf16 || !out_bf16 || talker_rows < 1 || group_rows < 1)
    return fail(QMK_ERR_ARG, "qmk_batched_embed_sum: bad argument");
  GroupTables gt;
  for (int g = 0; g < 15; ++g) {
    if (!group_tables[g]) return fail(QMK_ERR_ARG, "qmk_batched_embed_sum: null group table");
    gt.t[g] = reinterpret_cast<const __nv_bfloat16*>(group_tables[g]);
  }
  kb_embed_sum<<<batch, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(codes), reinterpret_cast<const __nv_bfloat16*>(talker_embed),
                                                        talker_rows, gt, group_rows, reinterpret_cast<const __nv_bfloat16*>(extra_bf16),
                                                        extra_stride, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(QMK_ERR_CUDA, cudaGetErrorString(e));
  return QMK_OK;
}

extern "C" int qmk_batched_counter_add(uint64_t* counter, uint64_t inc, void* stream) {
  if (!counter) return fail(QMK_ERR_ARG, "qmk_batched_counter_add: null counter");
  kb_counter_add<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(counter), (unsigned long long)inc);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(QMK_ERR_CUDA, cudaGetErrorString(e));
  return QMK_OK;
}

// One decode step for all B streams.  token_ids (int32[B], device; entry < 0 or null pointer -> the stream's row of
// `embeds` bf16[B][1024] is the input, the upstream sentinel path), positions (int32[B], device, advanced by one),
// k_cache / v_cache: [B][L][8][max_seq][128] bf16.  Outputs: tokens_out int32[B], hidden_out f32[B][1024].
extern "C" int qmk_batched_step(qmk_batched* h, const int32_t* token_ids, const void* embeds, int32_t* positions,
                                void* k_cache, void* v_cache, float* hidden_out, int32_t* tokens_out, void* stream) {
  if (!h || !positions || !k_cache || !v_cache || !tokens_out) return fail(QMK_ERR_ARG, "qmk_batched_step: null argument");
  if (!token_ids && !embeds) return fail(QMK_ERR_ARG, "qmk_batched_step: neither token ids nor embeddings given");
  cudaStream_t st = (cudaStream_t)stream;
  BatchedDeviceGuard guard(h->device);
  g_launch_err = cudaSuccess;
  if (h->decode_persistent)
    return launch_persistent(h, h->B, 0, token_ids, embeds, positions, k_cache, v_cache, hidden_out, tokens_out, st);
  qmk_batched_step_args a;
  memset(&a, 0, sizeof(a));
  a.token_ids = token_ids; a.embeds_bf16 = embeds; a.positions = positions; a.k_cache = k_cache; a.v_cache = v_cache;
  a.hidden_out = hidden_out; a.tokens_out = tokens_out; a.head = 0; a.group = -1;
  return chain_step(h, &a, st);
}

// Prefill of ONE utterance as a single batched pass (upstream feeds the 8 prefill embeddings
// through 8 sequential decode steps, tts_engine.py:281-282): lane i is position position0 + i, all lanes share the caller's
// B = 1 cache [L][8][max_seq][128] and attend causally.  Writes the same KV rows as n sequential steps and returns the LAST
// position's post-norm hidden state (f32[1024]) and argmax token; n <= batch.  Needs the persistent step kernel.
// The prefill through the launch chain (default): the decode chain with lane = position, the QKV epilogue split into "finish"
// (all lanes' K / V rows into the shared cache) and causal attention; 9 launches per layer.
static int chain_prefill(qmk_batched* h, const void* embeds, int n, int position0, void* k_cache, void* v_cache, float* hid_all, cudaStream_t st) {
  const int B = h->B, L = h->L;
  const float scale = 0.08838834764831845f;
  __nv_bfloat16* kc = reinterpret_cast<__nv_bfloat16*>(k_cache);
  __nv_bfloat16* vc = reinterpret_cast<__nv_bfloat16*>(v_cache);
  const __nv_bfloat16* cos_t = reinterpret_cast<const __nv_bfloat16*>(h->cos_t);
  const __nv_bfloat16* sin_t = reinterpret_cast<const __nv_bfloat16*>(h->sin_t);
  g_launch_err = cudaSuccess;
  kb_lane_positions<<<1, 64, 0, st>>>(h->pos_scratch, position0, B);
  launch_pdl(kb_input, dim3(B), dim3(256), 0, st, (const int*)nullptr, reinterpret_cast<const __nv_bfloat16*>(h->embed), h->head_rows,
             reinterpret_cast<const __nv_bfloat16*>(embeds), (const float*)nullptr, h->res, reinterpret_cast<const __nv_bfloat16*>(h->ln_in[0]), h->xn, n);
  for (int l = 0; l < L; ++l) {
    gemm(h, h->map_qkv[l], h->map_x1024, QKV_ROWS, H, 4, st);
    launch_pdl(kb_qkv_finish_prefill<4>, dim3(n, NKVH), dim3(128), 0, st, (const float*)h->partial, B, position0,
               reinterpret_cast<const __nv_bfloat16*>(h->qn[l]), reinterpret_cast<const __nv_bfloat16*>(h->kn[l]), cos_t, sin_t, kc, vc, h->qpre, l, h->max_seq);
    launch_pdl(kb_attention_prefill, dim3(n, NKVH), dim3(256), 0, st, (const float*)h->qpre, position0, (const __nv_bfloat16*)kc, (const __nv_bfloat16*)vc,
               h->abuf, l, h->max_seq, scale);
    gemm(h, h->map_o[l], h->map_x2048, H, QSZ, h->splits_od, st);
    resid_norm(h, st, h->ln_post[l], nullptr, nullptr);
    gemm(h, h->map_gu[l], h->map_x1024, GU_ROWS, H, 4, st);
    launch_pdl(kb_gu_epilogue<4>, dim3(B, 3), dim3(256), 0, st, (const float*)h->partial, B, h->mbuf);
    gemm(h, h->map_down[l], h->map_x3072, H, INTER, h->splits_od, st);
    const bool last = (l == L - 1);
    resid_norm(h, st, last ? h->final_norm : h->ln_in[l + 1], last ? hid_all : (float*)nullptr, nullptr);
  }
  gemm(h, h->map_head, h->map_x1024, h->head_rows, H, 4, st);
  HeadSelect sel;
  memset(&sel, 0, sizeof(sel));
  sel.group = -1; sel.temperature = 1.0f;
  launch_pdl(kb_head_epilogue, dim3(B), dim3(256), (size_t)4 * h->head_rows * sizeof(float), st, (const float*)h->partial, 4, B, h->head_rows,
             h->tok_scratch, h->pos_scratch, sel);
  cudaError_t e = g_launch_err != cudaSuccess ? g_launch_err : cudaGetLastError();
  if (e != cudaSuccess) { cudaGetLastError(); return fail(QMK_ERR_CUDA, cudaGetErrorString(e)); }
  return QMK_OK;
}

extern "C" int qmk_batched_prefill(qmk_batched* h, const void* embeds, int n, int position0, void* k_cache, void* v_cache,
                                   float* hidden_out_last, int32_t* token_out_last, void* stream) {
  if (!h || !embeds || !k_cache || !v_cache) return fail(QMK_ERR_ARG, "qmk_batched_prefill: null argument");
  if (n < 1 || n > h->B) return fail(QMK_ERR_ARG, "qmk_batched_prefill: n must be in [1, batch]");
  if (position0 < 0 || position0 + n > h->max_seq) return fail(QMK_ERR_ARG, "qmk_batched_prefill: positions exceed max_seq_len");
  cudaStream_t st = (cudaStream_t)stream;
  BatchedDeviceGuard guard(h->device);
  if (!h->prefill_persistent) {
    float* hid = reinterpret_cast<float*>(h->partial) + qmkb_hidden_offset();   // scratch behind the partials: [B][1024]
    const int rc = chain_prefill(h, embeds, n, position0, k_cache, v_cache, hid, st);
    if (rc != QMK_OK) return rc;
    cudaError_t e = cudaSuccess;
    if (hidden_out_last) e = cudaMemcpyAsync(hidden_out_last, hid + (size_t)(n - 1) * H, H * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && token_out_last) e = cudaMemcpyAsync(token_out_last, h->tok_scratch + (n - 1), sizeof(int), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail(QMK_ERR_CUDA, cudaGetErrorString(e));
    return QMK_OK;
  }
  if (!h->persistent) return fail(QMK_ERR_UNSUPPORTED, "QMK_PREFILL_PERSISTENT=1 needs the persistent step kernel");
  if (cudaMemcpyAsync(h->d_pos0, &position0, sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(QMK_ERR_CUDA, "qmk_batched_prefill: copy failed");
  float* hid_all = reinterpret_cast<float*>(h->partial) + qmkb_hidden_offset();   // scratch behind the partials: [n][1024]
  int rc = launch_persistent(h, n, 1, nullptr, embeds, h->d_pos0, k_cache, v_cache, hid_all, h->d_tok_scratch, st);
  if (rc != QMK_OK) return rc;
  cudaError_t e = cudaSuccess;
  if (hidden_out_last) e = cudaMemcpyAsync(hidden_out_last, hid_all + (size_t)(n - 1) * H, H * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && token_out_last) e = cudaMemcpyAsync(token_out_last, h->d_tok_scratch + (n - 1), sizeof(int), cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return fail(QMK_ERR_CUDA, cudaGetErrorString(e));
  return QMK_OK;
}

// bit 0: the persistent step kernel (qmk_bstep.cuh) is available (one-pass prefill); bit 1: decode steps use it too.
extern "C" int qmk_batched_is_persistent(const qmk_batched* h) { return h ? (h->persistent ? 1 : 0) + (h->decode_persistent ? 2 : 0) : 0; }
// Synchronise `stream`; QMK_ERR_KERNEL if a wait inside the persistent kernel timed out since the last call.
extern "C" int qmk_batched_sync_status(qmk_batched* h, void* stream) {
  if (!h) return fail(QMK_ERR_ARG, "qmk_batched_sync_status: null handle");
  BatchedDeviceGuard guard(h->device);
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return fail(QMK_ERR_CUDA, "qmk_batched_sync_status: synchronize failed");
  if (!h->d_status) return QMK_OK;
  int st = 0;
  cudaMemcpy(&st, h->d_status, sizeof(int), cudaMemcpyDeviceToHost);
  if (st != 0) {
    cudaMemset(h->d_status, 0, sizeof(int));
    unsigned cur = 0;
    cudaMemcpy(&cur, h->d_bar, sizeof(unsigned), cudaMemcpyDeviceToHost);
    h->bar_count = cur;   // an aborted launch leaves the barrier counter short: resynchronise the host's view
    return fail(QMK_ERR_KERNEL, "batched step: device watchdog fired");
  }
  return QMK_OK;
}

// Debug: barrier stamps (clock64 of CTA 0: enter / leave of every grid barrier) of the latest persistent step launched with
// QMK_BATCHED_TRACE set; returns the number of stamps copied.
extern "C" int qmk_batched_trace_read(qmk_batched* h, void* stream, long long* host_out, int max_elems) {
  if (!h || !host_out || !h->d_trace) return 0;
  BatchedDeviceGuard guard(h->device);
  cudaStreamSynchronize((cudaStream_t)stream);
  const int n = std::min(max_elems, 1024);
  cudaMemcpy(host_out, h->d_trace, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost);
  return n;
}

// Debug: timeline of the launch chain.  enable != 0 arms the stamps (the first and the last CTA of every chain kernel record
// %globaltimer at entry, after the grid dependency, at exit; process-wide, one device); a call with host_out copies the records
// {tag = kernel id * 16 + event * 2 + (last CTA), nanoseconds} gathered since the previous call and returns their number.
extern "C" int qmk_batched_chain_trace(int enable, void* stream, unsigned long long* host_out, int max_records) {
  static unsigned long long* d_buf = nullptr;
  constexpr size_t WORDS = 1 + 2 * 16000;
  if (enable && !d_buf) {
    if (cudaMalloc(&d_buf, WORDS * 8) != cudaSuccess) return fail(QMK_ERR_CUDA, "qmk_batched_chain_trace: allocation failed");
    cudaMemset(d_buf, 0, WORDS * 8);
    cudaMemcpyToSymbol(qmkb::g_ktrace, &d_buf, sizeof(d_buf));
  }
  if (!d_buf) return 0;
  cudaStreamSynchronize((cudaStream_t)stream);
  int n = 0;
  if (host_out && max_records > 0) {
    unsigned long long cnt = 0;
    cudaMemcpy(&cnt, d_buf, 8, cudaMemcpyDeviceToHost);
    n = (int)std::min<unsigned long long>(std::min<unsigned long long>(cnt, 16000), (unsigned long long)max_records);
    cudaMemcpy(host_out, d_buf + 1, (size_t)n * 16, cudaMemcpyDeviceToHost);
  }
  cudaMemset(d_buf, 0, 8);
  if (!enable) {
    unsigned long long* null_ptr = nullptr;
    cudaMemcpyToSymbol(qmkb::g_ktrace, &null_ptr, sizeof(null_ptr));
  }
  return n;
}

// ---- text side of the prefill: TextProjection.embed_text_ids on the same tcgen05 GEMM ----
#include "qmk_text.cuh"
