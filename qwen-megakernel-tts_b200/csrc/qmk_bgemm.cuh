// qmk_bgemm.cuh — batched-decode projection on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   P[split][n][row] (fp32) = sum_{k in K-slice `split`} W[row][k] * X[n][k]        W: [M, K] bf16 row-major (the
//   upstream [out, in] layout, read in place through a TMA tensor map), X: [N, K] bf16 (one row per stream).
//
// One CTA = one 128-row weight tile x one K slice: M = 128 is the UMMA M, N = number of streams (16 .. 64) is the
// UMMA N, so D = W_tile * X^T lives in N TMEM columns.  K is split across CTAs (split-K) because a decode GEMM has
// only M/128 = 8 .. 48 row tiles and all 148 SMs must pull weights to saturate HBM; the partials are summed by the
// fused epilogue kernels (qmk_batched.cu), which also apply the bf16 rounding / norm / RoPE / SwiGLU of the step.
// gridDim.z > 1 (text projection, qmk_text.cuh): z-th block of N rows of X against the same weights, partials behind each other.
// Warp 0 lane 0: TMA producer (all k-blocks of the slice are issued up front: <= 6 x 24 KB); warp 1 lane 0: MMA
// issuer (tcgen05.mma, cta_group::1, kind::f16, bf16 x bf16 -> fp32); warps 0-3: epilogue (tcgen05.ld 32x32b).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmkb {

constexpr int BM = 128;        // rows per CTA tile (UMMA M)
constexpr int BK = 64;         // k per block: 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int MAX_KB = 6;      // k-blocks per CTA (K slice <= 384)
constexpr int MAX_N = 64;
constexpr int TMEM_COLS = 64;  // power of two >= 32 and >= N

constexpr int A_TILE_BYTES = BM * BK * 2;       // 16 KB
constexpr int B_TILE_BYTES = MAX_N * BK * 2;    // 8 KB (N x 128 B used)
constexpr int SMEM_BYTES = MAX_KB * (A_TILE_BYTES + B_TILE_BYTES) + 1024 /*alignment slack*/ + 256;
// Shared memory a launch needs for K slices of `nkb` k-blocks.  The decode projections use slices of <= 4 k-blocks (97.25 KB):
// two CTAs fit on an SM, so a 192-CTA grid is ONE wave and the next projection's CTAs can become resident (and pull their
// weights) while this one's are still at work.
__host__ __device__ constexpr int smem_bytes_for(int nkb) { return nkb * (A_TILE_BYTES + B_TILE_BYTES) + 1024 + 256; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Programmatic dependent launch (PDL): a kernel of the step chain lets its successor start early and waits for its
// predecessor's results only where it needs them, so launch latency, barrier / TMEM set-up and -- for the GEMM -- the
// weight TMA loads overlap with the tail of the previous kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Debug timeline of the launch chain (qmk_batched_chain_trace): the first and the last CTA of every kernel read %globaltimer at
// entry (0), after the grid dependency resolved (1), at an intermediate point (2) and at exit (3); the four values stay in
// registers and are stored together at exit, so the stamps do not delay the work they time.  g_ktrace: [0] = next free
// record, records of 2 words {tag = kernel id * 16 + event * 2 + (last CTA), ns} follow.
__device__ unsigned long long* g_ktrace = nullptr;
struct KTrace {
  unsigned long long t[4];
  bool on, first;
  __device__ __forceinline__ KTrace() {
    first = (blockIdx.x | blockIdx.y) == 0;
    on = g_ktrace != nullptr && threadIdx.x == 0 && (first || (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1));
    t[0] = t[1] = t[2] = t[3] = 0;
  }
  __device__ __forceinline__ void mark(int e) {
    if (on) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t[e]));
  }
  __device__ __forceinline__ void flush(int kernel_id) {
    if (!on) return;
    mark(3);
    unsigned long long* buf = g_ktrace;
    const unsigned long long i = atomicAdd(buf, 4ull);
    if (i + 4 <= 16000) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { buf[1 + 2 * (i + e)] = (unsigned long long)(kernel_id * 16 + e * 2 + (first ? 0 : 1)); buf[2 + 2 * (i + e)] = t[e]; }
    }
  }
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major operand, SWIZZLE_128B, rows of 128 B stored
// contiguously, 8-row groups 1024 B apart.  bits 0-13 start >> 4 | 16-29 LBO >> 4 | 32-45 SBO >> 4 | 46-47 version = 1
// | 61-63 layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major layouts)
  d |= (uint64_t)(1024 >> 4) << 32;             // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D = f32 (bits 4-5 = 1), A = B = bf16 (bits 7-9,
// 10-12 = 1), both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__device__ __forceinline__ uint32_t make_instr_desc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

struct BgemmArgs {
  float* partial;   // [gridDim.z][splits][N][M]
  int M, N, K;      // N multiple of 16, <= 64; K multiple of 64 * splits
  int splits;
};

// grid = (M / 128, splits, row blocks of X), block = 128
__global__ void __launch_bounds__(128, 2) qmk_bgemm_kernel(const __grid_constant__ CUtensorMap map_w,
                                                           const __grid_constant__ CUtensorMap map_x, BgemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B tiles: 1 KB aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int kc = a.K / a.splits;          // K slice of this CTA
  const int k0 = blockIdx.y * kc;
  const int nkb = kc / BK;
  uint8_t* sA = smem;
  uint8_t* sB = smem + nkb * A_TILE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + nkb * (A_TILE_BYTES + B_TILE_BYTES));   // launch: smem_bytes_for(nkb)
  uint64_t* done = full + MAX_KB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_KB; ++i) mbar_init(&full[i], 1);
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // one warp allocates the accumulator columns and later frees them
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  pdl_launch_dependents();
  KTrace kt;
  kt.mark(0);

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: every k-block of the slice, up front.  The weight tiles do not depend on the previous
    //       kernel: they are in flight before the grid dependency is resolved; the activations follow it. =====
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_expect_tx(&full[kb], (uint32_t)(A_TILE_BYTES + a.N * BK * 2));
      tma_load_2d(sA + kb * A_TILE_BYTES, &map_w, k0 + kb * BK, m0, &full[kb]);
    }
    pdl_wait();
    for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sB + kb * B_TILE_BYTES, &map_x, k0 + kb * BK, (int)blockIdx.z * a.N, &full[kb]);
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_instr_desc(a.N);
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&full[kb], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t adesc = make_smem_desc(smem_u32(sA + kb * A_TILE_BYTES));
      const uint64_t bdesc = make_smem_desc(smem_u32(sB + kb * B_TILE_BYTES));
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k)   // +32 B per UMMA_K inside the 128 B swizzle row: +2 in 16-byte units
        umma_f16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
    }
    umma_commit(done);   // implies tcgen05.fence::before_thread_sync
  }
  __syncwarp();

  // ===== epilogue: TMEM lane = row of the tile, column = stream; warp w owns lanes 32 w .. 32 w + 31 =====
  pdl_wait();   // the previous epilogue kernel has finished reading the partial buffer this kernel overwrites
  kt.mark(1);
  mbar_wait(done, 0);
  kt.mark(2);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = m0 + warp * 32 + lane;
  float* out = a.partial + ((size_t)blockIdx.z * a.splits + blockIdx.y) * a.N * a.M;
  for (int n0 = 0; n0 < a.N; n0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (row < a.M) out[(size_t)(n0 + j) * a.M + row] = __uint_as_float(v[j]);   // consecutive lanes -> consecutive rows
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  kt.flush(1);
}

// ---- host: tensor maps through the driver entry point (no libcuda link dependency) -----------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], SWIZZLE_128B.  Returns 0 on success.
inline int make_tensor_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return -1;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace qmkb
