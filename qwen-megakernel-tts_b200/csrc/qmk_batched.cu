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
#include <string>
#include <vector>

#include "qmk_b200.h"
#include "qmk_bgemm.cuh"

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
__global__ void kb_input(const int* token_ids, const __nv_bfloat16* embed_table, int vocab, const __nv_bfloat16* embeds, float* res,
                         const __nv_bfloat16* w_in, __nv_bfloat16* xn) {
  __shared__ float s_red[8];
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  const int b = blockIdx.x, t = threadIdx.x;
  int tok = token_ids ? token_ids[b] : -1;
  if (tok >= vocab) tok = vocab - 1;             // device-side ids are not trusted: clamp instead of reading out of bounds
  if (tok < 0 && embeds == nullptr) tok = 0;     // sentinel without an embedding buffer
  const __nv_bfloat16* src = tok >= 0 ? embed_table + (size_t)tok * H : embeds + (size_t)b * H;
  const uint2 v = *reinterpret_cast<const uint2*>(src + t * 4);
  const float x[4] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y)};
  *reinterpret_cast<float4*>(res + (size_t)b * H + t * 4) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<uint2*>(xn + (size_t)b * H + t * 4) = rmsnorm4(x, w_in, s_red);
}

// ---- O / down epilogue: split-K sum -> bf16 -> residual -> next RMSNorm ----------------------------------------------
// grid = B, block = 256.  partial: [splits][B][1024].  hidden_out (optional, final norm only): f32[B][1024]
__global__ void kb_resid_norm(const float* partial, int splits, int B, float* res, int residual_fp32,
                              const __nv_bfloat16* w_norm, __nv_bfloat16* xn, float* hidden_out) {
  __shared__ float s_red[8];
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  const int b = blockIdx.x, t = threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 p = *reinterpret_cast<const float4*>(partial + ((size_t)s * B + b) * H + t * 4);
    acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
  }
  float4 r = *reinterpret_cast<const float4*>(res + (size_t)b * H + t * 4);
  const float o[4] = {bf16_round(acc.x), bf16_round(acc.y), bf16_round(acc.z), bf16_round(acc.w)};
  float x[4] = {r.x + o[0], r.y + o[1], r.z + o[2], r.w + o[3]};
  if (!residual_fp32) {
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = bf16_round(x[e]);
  }
  *reinterpret_cast<float4*>(res + (size_t)b * H + t * 4) = make_float4(x[0], x[1], x[2], x[3]);
  const uint2 n = rmsnorm4(x, w_norm, s_red);
  *reinterpret_cast<uint2*>(xn + (size_t)b * H + t * 4) = n;
  if (hidden_out)
    *reinterpret_cast<float4*>(hidden_out + (size_t)b * H + t * 4) = make_float4(bf16_lo(n.x), bf16_hi(n.x), bf16_lo(n.y), bf16_hi(n.y));
}

// ---- QKV epilogue + decode attention, one CTA per (stream, kv head) --------------------------------------------------
// Warps 0-3 first finish the projection for this group: split-K sum -> bf16 -> per-head RMSNorm + rotate-half RoPE in
// bf16 steps (q heads 2g, 2g+1 and k) -> q into shared memory, k / v appended to the cache row positions[b].  Then all
// 8 warps stride over positions 0 .. pos (lane owns 4 dims), fp32 online softmax, fixed-order cross-warp merge.
// grid = (B, 8), block = 256.  partial: [splits][B][4096] (q rows 0..2047, k 2048..3071, v 3072..4095).
__global__ void kb_qkv_attention(const float* partial, int splits, int B, const int* positions, const __nv_bfloat16* q_norm,
                                 const __nv_bfloat16* k_norm, const __nv_bfloat16* cos_t, const __nv_bfloat16* sin_t,
                                 __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, __nv_bfloat16* a_out, int layer, int L,
                                 int max_seq, float scale) {
  __shared__ float s_q[2][HD];
  __shared__ float s_acc[8][2][HD];
  __shared__ float s_m[8][2], s_l[8][2];
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  const int b = blockIdx.x, g = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int pos = positions[b];
  pos = pos < 0 ? 0 : (pos >= max_seq ? max_seq - 1 : pos);   // a stream that ran past its cache keeps rewriting the last row
  const size_t base = (((size_t)b * L + layer) * NKVH + g) * max_seq * HD;
  if (warp < 4) {
    // warp 0, 1: q heads 2g, 2g+1; warp 2: k head g; warp 3: v head g
    const int row0 = warp < 2 ? (2 * g + warp) * HD : (warp == 2 ? QSZ + g * HD : QSZ + KVSZ + g * HD);
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < splits; ++s) {
      const float4 p = *reinterpret_cast<const float4*>(partial + ((size_t)s * B + b) * QKV_ROWS + row0 + lane * 4);
      t[0] += p.x; t[1] += p.y; t[2] += p.z; t[3] += p.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) t[e] = bf16_round(t[e]);
    if (warp == 3) {
      *reinterpret_cast<uint2*>(v_cache + base + (size_t)pos * HD + lane * 4) =
          make_uint2(bf16_bits(t[0]) | (bf16_bits(t[1]) << 16), bf16_bits(t[2]) | (bf16_bits(t[3]) << 16));
    } else {
      const __nv_bfloat16* wn = warp < 2 ? q_norm : k_norm;
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) ss = fmaf(t[e], t[e], ss);
      ss = warp_sum(ss);
      const float rms = sqrtf(ss * (1.0f / HD) + EPS);
      const int dbase = (lane * 4) & 63;
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float n = bf16_round((t[e] / rms) * __bfloat162float(wn[lane * 4 + e]));
        const float other = __shfl_xor_sync(0xffffffffu, n, 16);
        const float cs = __bfloat162float(cos_t[(size_t)pos * HD + dbase + e]), sn = __bfloat162float(sin_t[(size_t)pos * HD + dbase + e]);
        const float x = bf16_round(n * cs), y = bf16_round(other * sn);
        o[e] = bf16_round(lane < 16 ? x - y : x + y);
      }
      if (warp < 2) {
        *reinterpret_cast<float4*>(&s_q[warp][lane * 4]) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        *reinterpret_cast<uint2*>(k_cache + base + (size_t)pos * HD + lane * 4) =
            make_uint2(bf16_bits(o[0]) | (bf16_bits(o[1]) << 16), bf16_bits(o[2]) | (bf16_bits(o[3]) << 16));
      }
    }
  }
  __syncthreads();   // q in shared memory; the new K / V row (global, written by this CTA) is visible to the whole CTA
  const int n = pos + 1;
  const float4 qa = *reinterpret_cast<const float4*>(&s_q[0][lane * 4]), qb = *reinterpret_cast<const float4*>(&s_q[1][lane * 4]);
  const float q0[4] = {qa.x, qa.y, qa.z, qa.w}, q1[4] = {qb.x, qb.y, qb.z, qb.w};
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f, acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
  for (int p = warp; p < n; p += 8) {
    const uint2 kk = *reinterpret_cast<const uint2*>(k_cache + base + (size_t)p * HD + lane * 4);
    const uint2 vv = *reinterpret_cast<const uint2*>(v_cache + base + (size_t)p * HD + lane * 4);
    const float kf[4] = {bf16_lo(kk.x), bf16_hi(kk.x), bf16_lo(kk.y), bf16_hi(kk.y)};
    const float vf[4] = {bf16_lo(vv.x), bf16_hi(vv.x), bf16_lo(vv.y), bf16_hi(vv.y)};
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
  if (lane == 0) { s_m[warp][0] = m0; s_m[warp][1] = m1; s_l[warp][0] = l0; s_l[warp][1] = l1; }
#pragma unroll
  for (int e = 0; e < 4; ++e) { s_acc[warp][0][lane * 4 + e] = acc0[e]; s_acc[warp][1][lane * 4 + e] = acc1[e]; }
  __syncthreads();
  const int h = threadIdx.x >> 7, d = threadIdx.x & 127;
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
  a_out[(size_t)b * QSZ + (2 * g + h) * HD + d] = __float2bfloat16_rn(A / Ls);
}

// ---- gate/up epilogue: m = r( r(silu(r(g))) * r(u) ) -------------------------------------------------------------------
// grid = (B, 3), block = 256; partial: [splits][B][6144] (gate rows 0..3071, up rows 3072..6143)
__global__ void kb_gu_epilogue(const float* partial, int splits, int B, __nv_bfloat16* m_out) {
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  const int b = blockIdx.x, j = blockIdx.y * 1024 + threadIdx.x * 4;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f), u = g;
  for (int s = 0; s < splits; ++s) {
    const float* p = partial + ((size_t)s * B + b) * GU_ROWS;
    const float4 pg = *reinterpret_cast<const float4*>(p + j), pu = *reinterpret_cast<const float4*>(p + INTER + j);
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
}

// ---- LM head epilogue: bf16 logits, argmax with lowest index on ties; advances the stream's position ----------------
// grid = B, block = 256; partial: [splits][B][rows]
__global__ void kb_head_epilogue(const float* partial, int splits, int B, int rows, int* tokens_out, int* positions) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  const int b = blockIdx.x;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int r = threadIdx.x; r < rows; r += 256) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[((size_t)s * B + b) * rows + r];
    const float v = bf16_round(acc);
    if (v > best) { best = v; best_i = r; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = best_i; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < best_i)) { best = s_v[w]; best_i = s_i[w]; }
    tokens_out[b] = best_i;
    positions[b] += 1;
  }
}

// concatenate row blocks of two / three [rows, K] matrices into one (one-time, at model creation)
__global__ void kb_copy_rows(const uint4* src, uint4* dst, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

thread_local std::string g_err;
int fail(int code, const char* msg) {
  g_err = msg;
  return code;
}

}  // namespace

constexpr size_t PARTIAL_ELEMS = (size_t)2 * 1024 * 1024;   // fp32 split-K partials: max splits x B x rows = 4 x 64 x 6144

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
  std::vector<const void*> w_o, w_down, ln_in, ln_post, qn, kn;
  const void *final_norm = nullptr, *lm_head = nullptr, *embed = nullptr, *cos_t = nullptr, *sin_t = nullptr;
  std::vector<CUtensorMap> map_qkv, map_o, map_gu, map_down;
  CUtensorMap map_head, map_x1024, map_x2048, map_x3072;
  float *res = nullptr, *partial = nullptr;
  __nv_bfloat16 *xn = nullptr, *abuf = nullptr, *mbuf = nullptr;
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
  if ((size_t)4 * batch * lm_head_rows > PARTIAL_ELEMS) return fail(QMK_ERR_ARG, "qmk_batched_create: lm_head_rows too large for the split-K partial buffer");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(QMK_ERR_ARG, "qmk_batched_create: bad device");
  BatchedDeviceGuard guard(device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) return fail(QMK_ERR_UNSUPPORTED, "qmk_batched_create: tcgen05 needs an sm_100 device");
  qmk_batched* h = new qmk_batched();
  h->device = device; h->L = num_layers; h->B = batch; h->max_seq = max_seq_len; h->head_rows = lm_head_rows;
  h->residual_fp32 = residual_fp32 ? 1 : 0;
  h->final_norm = final_norm_weight; h->lm_head = lm_head_weight; h->embed = embed_weight; h->cos_t = cos_table; h->sin_t = sin_table;
  const size_t L = num_layers;
  bool ok = cudaMalloc(&h->w_qkv, L * QKV_ROWS * H * 2) == cudaSuccess && cudaMalloc(&h->w_gu, L * GU_ROWS * H * 2) == cudaSuccess &&
            cudaMalloc(&h->res, (size_t)batch * H * 4) == cudaSuccess && cudaMalloc(&h->partial, PARTIAL_ELEMS * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&h->xn, (size_t)batch * H * 2) == cudaSuccess &&
            cudaMalloc(&h->abuf, (size_t)batch * QSZ * 2) == cudaSuccess && cudaMalloc(&h->mbuf, (size_t)batch * INTER * 2) == cudaSuccess;
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
  if (cudaFuncSetAttribute(qmk_bgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
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
  cudaFree(h->w_qkv); cudaFree(h->w_gu); cudaFree(h->res); cudaFree(h->partial);
  cudaFree(h->xn); cudaFree(h->abuf); cudaFree(h->mbuf);
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
static void gemm(qmk_batched* h, const CUtensorMap& mw, const CUtensorMap& mx, int M, int K, int splits, cudaStream_t st) {
  qmkb::BgemmArgs a{h->partial, M, h->B, K, splits};
  launch_pdl(qmkb::qmk_bgemm_kernel, dim3(M / qmkb::BM, splits), dim3(128), (size_t)qmkb::SMEM_BYTES, st, mw, mx, a);
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
  const int B = h->B, L = h->L;
  const float scale = 0.08838834764831845f;
  __nv_bfloat16* kc = reinterpret_cast<__nv_bfloat16*>(k_cache);
  __nv_bfloat16* vc = reinterpret_cast<__nv_bfloat16*>(v_cache);
  const __nv_bfloat16* cos_t = reinterpret_cast<const __nv_bfloat16*>(h->cos_t);
  const __nv_bfloat16* sin_t = reinterpret_cast<const __nv_bfloat16*>(h->sin_t);
  launch_pdl(kb_input, dim3(B), dim3(256), 0, st, (const int*)token_ids, reinterpret_cast<const __nv_bfloat16*>(h->embed), h->head_rows,
             reinterpret_cast<const __nv_bfloat16*>(embeds), h->res, reinterpret_cast<const __nv_bfloat16*>(h->ln_in[0]), h->xn);
  for (int l = 0; l < L; ++l) {
    gemm(h, h->map_qkv[l], h->map_x1024, QKV_ROWS, H, 4, st);                       // 32 tiles x 4 K-slices
    launch_pdl(kb_qkv_attention, dim3(B, NKVH), dim3(256), 0, st, (const float*)h->partial, 4, B, (const int*)positions,
               reinterpret_cast<const __nv_bfloat16*>(h->qn[l]), reinterpret_cast<const __nv_bfloat16*>(h->kn[l]), cos_t, sin_t,
               kc, vc, h->abuf, l, L, h->max_seq, scale);
    gemm(h, h->map_o[l], h->map_x2048, H, QSZ, 16, st);                              // 8 tiles x 16 K-slices
    launch_pdl(kb_resid_norm, dim3(B), dim3(256), 0, st, (const float*)h->partial, 16, B, h->res, h->residual_fp32,
               reinterpret_cast<const __nv_bfloat16*>(h->ln_post[l]), h->xn, (float*)nullptr);
    gemm(h, h->map_gu[l], h->map_x1024, GU_ROWS, H, 4, st);                          // 48 tiles x 4 K-slices
    launch_pdl(kb_gu_epilogue, dim3(B, 3), dim3(256), 0, st, (const float*)h->partial, 4, B, h->mbuf);
    gemm(h, h->map_down[l], h->map_x3072, H, INTER, 16, st);                         // 8 tiles x 16 K-slices
    const bool last = (l == L - 1);
    launch_pdl(kb_resid_norm, dim3(B), dim3(256), 0, st, (const float*)h->partial, 16, B, h->res, h->residual_fp32,
               reinterpret_cast<const __nv_bfloat16*>(last ? h->final_norm : h->ln_in[l + 1]), h->xn,
               last ? hidden_out : (float*)nullptr);
  }
  gemm(h, h->map_head, h->map_x1024, h->head_rows, H, 4, st);
  launch_pdl(kb_head_epilogue, dim3(B), dim3(256), 0, st, (const float*)h->partial, 4, B, h->head_rows, (int*)tokens_out, (int*)positions);
  cudaError_t e = g_launch_err != cudaSuccess ? g_launch_err : cudaGetLastError();
  if (e != cudaSuccess) { cudaGetLastError(); return fail(QMK_ERR_CUDA, cudaGetErrorString(e)); }
  return QMK_OK;
}
