// Exchange micro-benchmarks, part 2: ping-pong access-type variants, flag+bulk design, cluster/DSMEM push design.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ u64 ll_ld(const u64* p) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void ll_st(u64* p, unsigned payload, unsigned epoch) {
  u64 v = ((u64)epoch << 32) | payload; asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int ST, int LD>
__global__ void k_pingpong(u64* a, u64* b, int iters, int peer, long long* out) {
  if (threadIdx.x != 0) return;
  auto st = [](u64* p, u64 v) {
    if (ST == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (ST == 1) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (ST == 2) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (ST == 3) asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (ST == 4) { u64 o; asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(o) : "l"(p), "l"(v) : "memory"); }
    if (ST == 5) asm volatile("st.global.wt.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  };
  auto ld = [](const u64* p) {
    u64 v;
    if (LD == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 1) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 2) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 3) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
  };
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) { st(a, (u64)i); while (ld(b) != (u64)i) {} }
    out[0] = clock64() - t0;
  } else if (blockIdx.x == peer) {
    for (int i = 1; i <= iters; ++i) { while (ld(a) != (u64)i) {} st(b, (u64)i); }
  }
}

// pipelined polling: 4 loads in flight
__global__ void k_pingpong_pipe(u64* a, u64* b, int iters, int peer, long long* out) {
  if (threadIdx.x >= 4) return;
  // lanes 0..3 poll in a staggered fashion; any lane seeing the value ends the wait (ballot)
  auto wait = [&](const u64* p, u64 want) {
    for (;;) { u64 v = ll_ld(p); if (__any_sync(0xf, v == want)) return; }
  };
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) { if (threadIdx.x == 0) ll_st(a, 0, i); wait(b, (u64)i << 32); }
    if (threadIdx.x == 0) out[0] = clock64() - t0;
  } else if (blockIdx.x == peer) {
    for (int i = 1; i <= iters; ++i) { wait(a, (u64)i << 32); if (threadIdx.x == 0) ll_st(b, 0, i); }
  }
}

__global__ void k_gtimer(long long* out) {
  u64 t0, t1; int n = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); ++n; } while (t1 == t0);
  u64 t2 = t1; int m = 0;
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2)); ++m; } while (t2 == t1);
  out[0] = (long long)(t2 - t1); out[1] = m;
}

// Design P: plain data stores + release-increment of one counter per round; consumer: one thread polls, all bulk-load.
__global__ void k_flagbulk(float* data, unsigned* counters, int words, int iters, long long* out, float* sink) {
  extern __shared__ float sf[];
  const int G = gridDim.x, cta = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  const int w0 = (int)((long long)cta * words / G), w1 = (int)((long long)(cta + 1) * words / G);
  float acc = 0;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    float* d = data + (size_t)(it & 1) * words;
    unsigned* ctr = counters + (it & 1) * 32;
    if (tid < w1 - w0) {
      asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(d + w0 + tid), "f"((float)(it + tid)) : "memory");
    }
    __syncthreads();
    if (tid == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      unsigned want = (unsigned)G * ((it + 1) / 2), v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < want);
    }
    __syncthreads();
    for (int i = tid * 4; i < words; i += T * 4) {
      float4 v; asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(d + i) : "memory");
      *reinterpret_cast<float4*>(sf + i) = v;
    }
    __syncthreads();
    acc += sf[(tid * 7 + it) % words];
    __syncthreads();
  }
  if (tid == 0) out[cta] = clock64() - t0;
  sink[cta * T + tid] = acc;
}

// Design H: global LL publish; each CTA of a cluster fetches 1/C of the vector and pushes it into every
// cluster member's shared memory with st.async (complete_tx on the member's mbarrier).
__global__ void k_cluster(u64* buf, int words, int iters, long long* out, float* sink) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* bar = reinterpret_cast<u64*>(smem_raw);          // 1 mbarrier
  unsigned* recv = reinterpret_cast<unsigned*>(smem_raw + 16);
  cg::cluster_group cl = cg::this_cluster();
  const int C = cl.num_blocks(), rank = cl.block_rank();
  const int G = gridDim.x, cta = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  const int w0 = (int)((long long)cta * words / G), w1 = (int)((long long)(cta + 1) * words / G);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cl.sync();
  const int per = words / C;  // words fetched by this CTA
  float acc = 0;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    u64* b = buf + (size_t)(it & 1) * words;
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(words * 4) : "memory");
    if (tid < w1 - w0) ll_st(b + w0 + tid, (unsigned)(it + tid), (unsigned)it);
    for (int i = tid; i < per; i += T) {
      const int idx = rank * per + i;
      u64 w = ll_ld(b + idx);
      while ((unsigned)(w >> 32) != (unsigned)it) w = ll_ld(b + idx);
      const unsigned local_addr = smem_u32(recv + idx), local_bar = smem_u32(bar);
      for (int c = 0; c < C; ++c) {
        unsigned ra, rb;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(c));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(local_bar), "r"(c));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"((unsigned)w), "r"(rb) : "memory");
      }
    }
    // wait for the full vector
    {
      unsigned ok = 0; const unsigned parity = (it - 1) & 1;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
      }
    }
    acc += __uint_as_float(recv[(tid * 7 + it) % words]);
    __syncthreads();
  }
  if (tid == 0) out[cta] = clock64() - t0;
  sink[cta * T + tid] = acc;
  cl.sync();
}

static double max_out(long long* out, int G) {
  std::vector<long long> h(G);
  CK(cudaMemcpy(h.data(), out, G * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  return (double)mx;
}

template <int ST, int LD>
static void pp(u64* ab, int G, long long* out, const char* name) {
  for (int peer : {1, 2}) {
    CK(cudaMemset(ab, 0, 4096));
    int iters = 2000; u64* a = ab; u64* b = ab + 64;
    void* args[] = {&a, &b, &iters, &peer, &out};
    CK(cudaLaunchCooperativeKernel((void*)k_pingpong<ST, LD>, dim3(G), dim3(32), args, 0, 0));
    CK(cudaDeviceSynchronize());
    long long r; CK(cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost));
    printf("pingpong %-28s peer=%d: one-way %.1f cycles\n", name, peer, (double)r / iters / 2);
  }
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int G = prop.multiProcessorCount;
  long long* out; CK(cudaMalloc(&out, 1024 * sizeof(long long)));
  float* sink; CK(cudaMalloc(&sink, 1024 * 1024 * 4));
  u64* ab; CK(cudaMalloc(&ab, 4096));
  k_gtimer<<<1, 1>>>(out); CK(cudaDeviceSynchronize());
  { long long r[2]; CK(cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost)); printf("globaltimer tick = %lld ns (%lld reads per tick)\n", r[0], r[1]); }
  pp<0, 0>(ab, G, out, "st.relaxed / ld.relaxed");
  pp<1, 1>(ab, G, out, "st.volatile / ld.volatile");
  pp<2, 2>(ab, G, out, "st.release / ld.acquire");
  pp<3, 0>(ab, G, out, "red.max / ld.relaxed");
  pp<4, 0>(ab, G, out, "atom.exch / ld.relaxed");
  pp<5, 3>(ab, G, out, "st.wt / ld.cv");
  pp<0, 3>(ab, G, out, "st.relaxed / ld.cv");
  {
    for (int peer : {1, 2}) {
      CK(cudaMemset(ab, 0, 4096));
      int iters = 2000; u64* a = ab; u64* b = ab + 64;
      void* args[] = {&a, &b, &iters, &peer, &out};
      CK(cudaLaunchCooperativeKernel((void*)k_pingpong_pipe, dim3(G), dim3(32), args, 0, 0));
      CK(cudaDeviceSynchronize());
      long long r; CK(cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost));
      printf("pingpong 4-lane polling              peer=%d: one-way %.1f cycles\n", peer, (double)r / iters / 2);
    }
  }
  // Design P
  {
    float* data; CK(cudaMalloc(&data, 2 * 4096 * 4)); unsigned* ctr; CK(cudaMalloc(&ctr, 256));
    for (int words : {1024, 3072}) {
      CK(cudaMemset(ctr, 0, 256));
      int iters = 2000;
      void* args[] = {&data, &ctr, &words, &iters, &out, &sink};
      CK(cudaLaunchCooperativeKernel((void*)k_flagbulk, dim3(G), dim3(448), args, words * 4, 0));
      CK(cudaDeviceSynchronize());
      printf("flag+bulk words=%d: %.1f cycles/round\n", words, max_out(out, G) / iters);
    }
  }
  // Design H
  {
    u64* buf; CK(cudaMalloc(&buf, 2 * 4096 * 8));
    for (int C : {1, 2, 4, 8}) {
      for (int words : {1024, 3072}) {
        size_t smem = 16 + words * 4;
        size_t smem_big = 200 * 1024;  // occupancy as in the real kernel
        CK(cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
        CK(cudaFuncSetAttribute(k_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(448); cfg.dynamicSmemBytes = smem_big; cfg.stream = 0;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeCooperative; attr[1].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cfg.gridDim = dim3(C);
        int nclusters = 0;
        CK(cudaOccupancyMaxActiveClusters(&nclusters, (void*)k_cluster, &cfg));
        int Gc = nclusters * C; if (Gc > G) Gc = G / C * C;
        cfg.gridDim = dim3(Gc);
        (void)smem;
        if (words % C) continue;
        CK(cudaMemset(buf, 0, 2 * 4096 * 8));
        int iters = 2000;
        cfg.numAttrs = 2;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_cluster, buf, words, iters, out, sink);
        if (e != cudaSuccess) { printf("cluster C=%d launch (coop) failed: %s; retry non-coop\n", C, cudaGetErrorString(e)); cudaGetLastError(); cfg.numAttrs = 1; CK(cudaLaunchKernelEx(&cfg, k_cluster, buf, words, iters, out, sink)); }
        CK(cudaDeviceSynchronize());
        printf("cluster push C=%d (max clusters %d -> %d CTAs) words=%d: %.1f cycles/round\n", C, nclusters, Gc, words, max_out(out, Gc) / iters);
      }
    }
  }
  return 0;
}
