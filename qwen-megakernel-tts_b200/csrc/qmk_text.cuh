// qmk_text.cuh — the text side of the prefill: upstream TextProjection.embed_text_ids
// (model_tts.py:348-374), once per utterance on its content tokens (tts_engine.py:262-263):
//
//     x  = text_embedding[ids]                       bf16 [T][2048]
//     y  = bf16( silu( bf16( x fc1_w^T + fc1_b ) ) )  bf16 [T][2048]
//     out = bf16( y fc2_w^T + fc2_b )                 bf16 [T][1024]
//
// Upstream issues four PyTorch operators (embedding, linear, silu, linear).  Here a pass of up to 512 tokens is one chain of
// five kernels under programmatic dependent launch: row gather -> fc1 on tcgen05 (qmk_bgemm_kernel: UMMA M = 128 weight rows,
// N = tokens, split-K 8 so that 128 CTAs pull the 8.4 MB of fc1) -> split-K sum + bias + SiLU -> fc2 on tcgen05 (split-K 16,
// 128 CTAs, 4.2 MB) -> split-K sum + bias.  More than 64 tokens are blocks of 64 along gridDim.z of the same launches (every
// block multiplies the same weight tiles, which come from L2 after the first).  The weights are read in place in the upstream
// [out, in] layout through TMA tensor maps.  Rounding points are upstream's: the products accumulate in fp32, the bias is added
// in fp32, one rounding to bf16 per operator (F.linear, F.silu on bf16).
//
// Included at the end of qmk_batched.cu (same translation unit: the tcgen05 GEMM, the PDL launcher and the error helpers).
#pragma once

namespace {

constexpr int TXT_H = 2048;                 // text hidden size (fc1: 2048 -> 2048, fc2: 2048 -> 1024)
constexpr int TXT_LANES = 64;               // tokens per block = the UMMA N limit
constexpr int TXT_BLOCKS = 8;               // blocks per pass (512 tokens)
constexpr int TXT_SPLITS1 = 8, TXT_SPLITS2 = 16;

// grid = lanes (multiple of 16), block = 256: lane n < n_valid copies table row ids[n] (16 bytes per thread), spare lanes are
// zero (they feed the tensor cores too).  Ids come from device memory and are clamped instead of trusted.
__global__ void __launch_bounds__(256) kt_gather(const long long* ids, int n_valid, const uint4* table, int vocab, uint4* x0) {
  qmkb::pdl_wait();
  qmkb::pdl_launch_dependents();
  const int n = blockIdx.x;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (n < n_valid) {
    long long id = ids[n];
    id = id < 0 ? 0 : (id >= vocab ? (long long)vocab - 1 : id);
    v = table[(size_t)id * (TXT_H / 8) + threadIdx.x];
  }
  x0[(size_t)n * (TXT_H / 8) + threadIdx.x] = v;
}

// partial: [blocks][splits][N][M]; token n = lane n % N of block n / N
__device__ __forceinline__ float4 kt_split_sum(const float* partial, int splits, int N, int M, int n, int r) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* base = partial + (((size_t)(n / N) * splits) * N + n % N) * M + r;
#pragma unroll 8
  for (int s = 0; s < splits; ++s) {
    const float4 p = __ldcg(reinterpret_cast<const float4*>(base + (size_t)s * N * M));
    acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
  }
  return acc;
}
__device__ __forceinline__ float kt_silu_bf16(float pre) {
  const float y = bf16_round(pre);                 // F.linear's bf16 output
  return bf16_round(y / (1.0f + expf(-y)));        // F.silu on bf16: fp32 arithmetic, one rounding
}
// grid = (lanes of all blocks, 2), block = 256; thread = 4 consecutive fc1 rows of one token.  partial: [blocks][SPLITS1][N][2048].
__global__ void __launch_bounds__(256) kt_fc1_epilogue(const float* partial, int N, const __nv_bfloat16* bias, __nv_bfloat16* x1) {
  qmkb::pdl_launch_dependents();
  const int n = blockIdx.x, r = blockIdx.y * 1024 + threadIdx.x * 4;
  const uint2 bv = *reinterpret_cast<const uint2*>(bias + r);
  qmkb::pdl_wait();
  const float4 a = kt_split_sum(partial, TXT_SPLITS1, N, TXT_H, n, r);
  const __nv_bfloat162 y01 = __floats2bfloat162_rn(kt_silu_bf16(a.x + bf16_lo(bv.x)), kt_silu_bf16(a.y + bf16_hi(bv.x)));
  const __nv_bfloat162 y23 = __floats2bfloat162_rn(kt_silu_bf16(a.z + bf16_lo(bv.y)), kt_silu_bf16(a.w + bf16_hi(bv.y)));
  *reinterpret_cast<uint2*>(x1 + (size_t)n * TXT_H + r) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&y01), *reinterpret_cast<const uint32_t*>(&y23));
}
// grid = valid tokens, block = 256; thread = 4 consecutive fc2 rows.  partial: [blocks][SPLITS2][N][1024]; out: bf16[tokens][1024].
__global__ void __launch_bounds__(256) kt_fc2_epilogue(const float* partial, int N, const __nv_bfloat16* bias, __nv_bfloat16* out) {
  qmkb::pdl_launch_dependents();
  const int n = blockIdx.x, r = threadIdx.x * 4;
  const uint2 bv = *reinterpret_cast<const uint2*>(bias + r);
  qmkb::pdl_wait();
  const float4 a = kt_split_sum(partial, TXT_SPLITS2, N, H, n, r);
  const __nv_bfloat162 y01 = __floats2bfloat162_rn(a.x + bf16_lo(bv.x), a.y + bf16_hi(bv.x));
  const __nv_bfloat162 y23 = __floats2bfloat162_rn(a.z + bf16_lo(bv.y), a.w + bf16_hi(bv.y));
  *reinterpret_cast<uint2*>(out + (size_t)n * H + r) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&y01), *reinterpret_cast<const uint32_t*>(&y23));
}

}  // namespace

struct qmk_text_proj {
  int device = 0, vocab = 0;
  const void *table = nullptr, *fc1_b = nullptr, *fc2_b = nullptr;
  CUtensorMap map_w1, map_w2;
  CUtensorMap map_x0[TXT_LANES / 16], map_x1[TXT_LANES / 16];   // activations as UMMA N = 16 / 32 / 48 / 64 rows
  __nv_bfloat16 *x0 = nullptr, *x1 = nullptr;                   // [512][2048] gathered rows / SiLU outputs
  float* partial = nullptr;                                      // split-K partials: 8 blocks x max(8 x 64 x 2048, 16 x 64 x 1024) floats
  std::mutex mu;                                                 // host-side enqueue of one call at a time
  cudaStream_t last_stream = nullptr;                            // the staging buffers are shared: calls are ordered across streams
  bool has_last_stream = false;
  cudaEvent_t handover = nullptr;
};

// A call on another stream than the handle's previous call first waits for that stream (the calls share x0 / x1 / partial).
// Calls recorded into a CUDA graph are left alone: whoever replays the graph orders it.
static void text_order_after_previous_stream(qmk_text_proj* h, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return; }
  if (cs != cudaStreamCaptureStatusNone) return;
  if (h->has_last_stream && h->last_stream != st) {
    cudaError_t err = h->handover ? cudaSuccess : cudaEventCreateWithFlags(&h->handover, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventRecord(h->handover, h->last_stream);
    if (err == cudaSuccess) err = cudaStreamWaitEvent(st, h->handover, 0);
    if (err != cudaSuccess) {   // the previous stream no longer exists: its work is ordered by a device-wide wait
      cudaGetLastError();
      cudaDeviceSynchronize();
    }
  }
  h->last_stream = st;
  h->has_last_stream = true;
}

extern "C" void qmk_text_proj_destroy(qmk_text_proj* h) {
  if (!h) return;
  BatchedDeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  cudaFree(h->x0); cudaFree(h->x1); cudaFree(h->partial);
  if (h->handover) cudaEventDestroy(h->handover);
  delete h;
}

extern "C" int qmk_text_proj_create(int device, const void* text_embedding, int vocab_rows, const void* fc1_weight,
                                    const void* fc1_bias, const void* fc2_weight, const void* fc2_bias, qmk_text_proj** out) {
  using namespace qmkb;
  if (!text_embedding || !fc1_weight || !fc1_bias || !fc2_weight || !fc2_bias || !out)
    return fail(QMK_ERR_ARG, "qmk_text_proj_create: null argument");
  *out = nullptr;
  if (vocab_rows < 1) return fail(QMK_ERR_ARG, "qmk_text_proj_create: vocab_rows must be positive");
  for (const void* p : {text_embedding, fc1_weight, fc1_bias, fc2_weight, fc2_bias})
    if (reinterpret_cast<uintptr_t>(p) & 15) return fail(QMK_ERR_ARG, "qmk_text_proj_create: tensors must be 16-byte aligned");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(QMK_ERR_ARG, "qmk_text_proj_create: bad device");
  BatchedDeviceGuard guard(device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) return fail(QMK_ERR_UNSUPPORTED, "qmk_text_proj_create: tcgen05 needs an sm_100 device");
  qmk_text_proj* h = new qmk_text_proj();
  h->device = device; h->vocab = vocab_rows; h->table = text_embedding; h->fc1_b = fc1_bias; h->fc2_b = fc2_bias;
  const size_t act = (size_t)TXT_BLOCKS * TXT_LANES * TXT_H * 2;
  const size_t part = (size_t)TXT_BLOCKS * TXT_LANES * std::max(TXT_SPLITS1 * TXT_H, TXT_SPLITS2 * H) * sizeof(float);
  bool ok = cudaMalloc(&h->x0, act) == cudaSuccess && cudaMalloc(&h->x1, act) == cudaSuccess && cudaMalloc(&h->partial, part) == cudaSuccess &&
            cudaMemset(h->x0, 0, act) == cudaSuccess && cudaMemset(h->x1, 0, act) == cudaSuccess;
  if (!ok) { qmk_text_proj_destroy(h); return fail(QMK_ERR_CUDA, "qmk_text_proj_create: allocation failed"); }
  int rc = make_tensor_map(&h->map_w1, fc1_weight, TXT_H, TXT_H, BM);
  rc |= make_tensor_map(&h->map_w2, fc2_weight, H, TXT_H, BM);
  for (int i = 0; i < TXT_LANES / 16; ++i) {
    rc |= make_tensor_map(&h->map_x0[i], h->x0, TXT_BLOCKS * TXT_LANES, TXT_H, 16 * (i + 1));
    rc |= make_tensor_map(&h->map_x1[i], h->x1, TXT_BLOCKS * TXT_LANES, TXT_H, 16 * (i + 1));
  }
  if (rc) { qmk_text_proj_destroy(h); return fail(QMK_ERR_CUDA, "qmk_text_proj_create: cuTensorMapEncodeTiled failed"); }
  if (cudaFuncSetAttribute(qmk_bgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    qmk_text_proj_destroy(h);
    return fail(QMK_ERR_CUDA, "qmk_text_proj_create: kernel setup failed");
  }
  *out = h;
  return QMK_OK;
}

// ids: int64[n_ids] in DEVICE memory (clamped to the table); out: bf16[n_ids][1024] in device memory.  Asynchronous on `stream`;
// only enqueues kernels (no allocation, no synchronisation), so it may be captured in a CUDA graph.  Calls on one handle share
// its staging buffers: a call on another stream than the previous one waits for it (event); captured calls are not ordered here.
extern "C" int qmk_text_proj_embed(qmk_text_proj* h, const int64_t* ids, int n_ids, void* out_bf16, void* stream) {
  using namespace qmkb;
  if (!h || n_ids < 0 || (n_ids > 0 && (!ids || !out_bf16))) return fail(QMK_ERR_ARG, "qmk_text_proj_embed: bad argument");
  if (n_ids == 0) return QMK_OK;
  BatchedDeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(h->mu);
  text_order_after_previous_stream(h, st);
  g_launch_err = cudaSuccess;
  for (int off = 0; off < n_ids; off += TXT_BLOCKS * TXT_LANES) {
    // one pass: nb blocks of N lanes (a single block is trimmed to the next multiple of 16 tokens)
    const int nv = std::min(TXT_BLOCKS * TXT_LANES, n_ids - off), nb = (nv + TXT_LANES - 1) / TXT_LANES;
    const int N = nb > 1 ? TXT_LANES : (nv + 15) & ~15, mi = N / 16 - 1;
    launch_pdl(kt_gather, dim3(nb * N), dim3(256), 0, st, reinterpret_cast<const long long*>(ids) + off, nv,
               reinterpret_cast<const uint4*>(h->table), h->vocab, reinterpret_cast<uint4*>(h->x0));
    BgemmArgs a1{h->partial, TXT_H, N, TXT_H, TXT_SPLITS1};
    launch_pdl(qmk_bgemm_kernel, dim3(TXT_H / BM, TXT_SPLITS1, nb), dim3(128), (size_t)smem_bytes_for(TXT_H / TXT_SPLITS1 / BK), st,
               h->map_w1, h->map_x0[mi], a1);
    launch_pdl(kt_fc1_epilogue, dim3(nb * N, 2), dim3(256), 0, st, (const float*)h->partial, N,
               reinterpret_cast<const __nv_bfloat16*>(h->fc1_b), h->x1);
    BgemmArgs a2{h->partial, H, N, TXT_H, TXT_SPLITS2};
    launch_pdl(qmk_bgemm_kernel, dim3(H / BM, TXT_SPLITS2, nb), dim3(128), (size_t)smem_bytes_for(TXT_H / TXT_SPLITS2 / BK), st,
               h->map_w2, h->map_x1[mi], a2);
    launch_pdl(kt_fc2_epilogue, dim3(nv), dim3(256), 0, st, (const float*)h->partial, N,
               reinterpret_cast<const __nv_bfloat16*>(h->fc2_b), reinterpret_cast<__nv_bfloat16*>(out_bf16) + (size_t)off * H);
  }
  if (g_launch_err != cudaSuccess) return fail(QMK_ERR_CUDA, cudaGetErrorString(g_launch_err));
  return QMK_OK;
}
