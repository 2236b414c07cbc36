// Exchange micro-benchmark 7 (v2 engine primitives), 8 clusters x 16 CTAs (one cluster per kv head):
//   A. "red-gather": K-split partial sums meet in L2.  Every CTA adds 64 fixed-point partials with red.add.u64
//      (value in the low 56 bits, arrival count in the high 8 bits -> integer adds are order-independent, no zeroing,
//      no epoch: use n of a word is complete when its count reaches 8 n), then every CTA gathers all 1024 words.
//   B. group-local LL4 exchange through L2: every CTA publishes 32 LL words {epoch, payload}; the 16 CTAs of a group
//      gather the group's 512 words.  (Clusters of 16 were the first choice -- DSMEM push, 215 cycles -- but
//      cudaOccupancyMaxActiveClusters reports only 7 co-resident 16-CTA clusters and 15 8-CTA clusters on this B200,
//      one GPC being smaller, so 8 groups x 16 cannot be cluster-resident.)
//   C. a layer-shaped sequence: A, B, A, B with `work` cycles of compute between them.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void red64(u64* p, u64 v) { asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void ld2x64(const u64* p, u64& a, u64& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster32(unsigned addr, unsigned v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void spin(int cyc) { if (cyc > 0) { long long t = clock64(); while (clock64() - t < cyc) {} } }

struct P { unsigned* gb; int gmap; u64* acc; int iters, mode, delay, cdelay, work, jitter; long long* out; unsigned* sink; };

constexpr int CS = 16, NT = 256;

__device__ __forceinline__ void red_gather(const P& p, u64* acc, int j, unsigned use, float* s_vec, unsigned& chk) {
  const int tid = threadIdx.x;
  // contribute: this CTA's 64 rows (rows 64 j .. 64 j + 63), one add per row
  if (tid < 64) red64(acc + 64 * j + tid, (1ull << 56) + (u64)(tid + 1));
  spin(p.delay);
  const unsigned want = (8u * use) & 0xffu;
  // gather: thread owns words 4 tid .. 4 tid + 3
  u64 w[4];
  for (unsigned sp = 0;; ++sp) {
    if (sp > (1u << 14) || *(volatile unsigned*)p.sink == 0xdeadu) { p.sink[0] = 0xdeadu; break; }
    ld2x64(acc + 4 * tid, w[0], w[1]);
    ld2x64(acc + 4 * tid + 2, w[2], w[3]);
    bool ok = true;
#pragma unroll
    for (int e = 0; e < 4; ++e) ok &= ((unsigned)((w[e] + (1ull << 55)) >> 56) & 0xffu) == want;
    if (ok) break;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { s_vec[4 * tid + e] = (float)(long long)(w[e] << 8); chk += (unsigned)w[e]; }
  __syncthreads();
}

// variant: every CTA adds a partial to ALL 1024 words (128 adds per word): the down projection K-split per CTA
__device__ __forceinline__ void red_gather_all(const P& p, u64* acc, unsigned use, float* s_vec, unsigned& chk) {
  const int tid = threadIdx.x;
  const int rot = (blockIdx.x * 8) & 1023;   // stagger the start word per CTA
#pragma unroll
  for (int e = 0; e < 4; ++e) red64(acc + ((tid + 256 * e + rot) & 1023), (1ull << 56) + (u64)(tid + 1));
  spin(p.delay);
  const unsigned want = (128u * use) & 0xffu;
  u64 w[4];
  for (unsigned sp = 0;; ++sp) {
    if (sp > (1u << 14) || *(volatile unsigned*)p.sink == 0xdeadu) { p.sink[0] = 0xdeadu; break; }
    ld2x64(acc + 4 * tid, w[0], w[1]);
    ld2x64(acc + 4 * tid + 2, w[2], w[3]);
    bool ok = true;
#pragma unroll
    for (int e = 0; e < 4; ++e) ok &= ((unsigned)((w[e] + (1ull << 55)) >> 56) & 0xffu) == want;
    if (ok) break;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { s_vec[4 * tid + e] = (float)(long long)(w[e] << 8); chk += (unsigned)w[e]; }
  __syncthreads();
}
__device__ __forceinline__ void st32(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 ld128(const unsigned* p) {
  uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void group_xchg(const P& p, unsigned* gbuf, int j, unsigned epoch, unsigned& chk) {
  const int tid = threadIdx.x;
  if (tid < 32) st32(gbuf + 32 * j + tid, (epoch << 16) | (unsigned)tid);
  spin(p.cdelay);
  if (tid < 128) {
    uint4 w;
    for (unsigned sp = 0;; ++sp) {
      if (sp > (1u << 14) || *(volatile unsigned*)p.sink == 0xdeadu) { p.sink[0] = 0xdeadu; p.sink[1] = 0xdeadu; break; }
      w = ld128(gbuf + 4 * tid);
      if ((w.x >> 16) == epoch && (w.y >> 16) == epoch && (w.z >> 16) == epoch && (w.w >> 16) == epoch) break;
    }
    chk += w.x + w.y + w.z + w.w;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT, 1) k_v2(P p) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* s_vec = reinterpret_cast<float*>(smem);
  const int cta = blockIdx.x;
  const int g = p.gmap ? (cta & 7) : (cta >> 4), j = p.gmap ? (cta >> 3) : (cta & 15);
  const int tid = threadIdx.x;
  unsigned chk = 0, rng = blockIdx.x * 2654435761u + 12345u;
  unsigned use_a = 0, use_b = 0, ep = 0;
  long long t0 = clock64();
  for (int it = 1; it <= p.iters; ++it) {
    if (*(volatile unsigned*)p.sink == 0xdeadu) break;
    rng = rng * 1664525u + 1013904223u;
    const int w = p.work + (p.jitter ? (int)((rng >> 8) % (unsigned)p.jitter) : 0);
    if (p.mode == 0) {   // two accumulators alternate: a word may only be re-used after a full exchange on the other one
      spin(w);
      if (it & 1) red_gather(p, p.acc, j, ++use_a, s_vec, chk); else red_gather(p, p.acc + 1024, j, ++use_b, s_vec, chk);
    } else if (p.mode == 3) {
      spin(w);
      if (it & 1) red_gather_all(p, p.acc, ++use_a, s_vec, chk); else red_gather_all(p, p.acc + 1024, ++use_b, s_vec, chk);
    } else if (p.mode == 4) {   // layer shape with the CTA-local down: A(all-words), group, B(64 rows)
      spin(w);
      red_gather_all(p, p.acc, ++use_a, s_vec, chk);
      spin(w);
      ++ep; group_xchg(p, p.gb + (ep & 1) * 8192 + g * 512, j, ep & 0xffffu, chk);
      spin(w);
      red_gather(p, p.acc + 1024, j, ++use_b, s_vec, chk);
      spin(w);
    } else if (p.mode == 1) {
      spin(w);
      ++ep;
      group_xchg(p, p.gb + (ep & 1) * 8192 + g * 512, j, ep & 0xffffu, chk);
    } else {
      spin(w);
      red_gather(p, p.acc, j, ++use_a, s_vec, chk);
      spin(w);
      ++ep; group_xchg(p, p.gb + (ep & 1) * 8192 + g * 512, j, ep & 0xffffu, chk);
      spin(w);
      red_gather(p, p.acc + 1024, j, ++use_b, s_vec, chk);
      spin(w);
      ++ep; group_xchg(p, p.gb + (ep & 1) * 8192 + g * 512, j, ep & 0xffffu, chk);
    }
  }
  if (tid == 0) p.out[blockIdx.x] = clock64() - t0;
  p.sink[8 + blockIdx.x * NT + tid] = chk;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("SMs %d\n", prop.multiProcessorCount);
  P p; CK(cudaMalloc(&p.out, 1024 * 8)); CK(cudaMalloc(&p.sink, 1 << 22)); CK(cudaMalloc(&p.acc, 2048 * 8)); CK(cudaMalloc(&p.gb, 2 * 8192 * 4));
  const size_t smem = 200 * 1024;
  CK(cudaFuncSetAttribute(k_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  auto run = [&](int mode, int gmap, int delay, int cdelay, int work, int jitter) {
    p.iters = 1000; p.mode = mode; p.gmap = gmap; p.delay = delay; p.cdelay = cdelay; p.work = work; p.jitter = jitter;
    CK(cudaMemset(p.acc, 0, 2048 * 8)); CK(cudaMemset(p.sink, 0, 8)); CK(cudaMemset(p.gb, 0, 2 * 8192 * 4));
    void* args[] = {&p};
    CK(cudaLaunchCooperativeKernel((void*)k_v2, dim3(128), dim3(NT), args, smem, 0));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(128); CK(cudaMemcpy(h.data(), p.out, 128 * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    unsigned flags[2]; CK(cudaMemcpy(flags, p.sink, 8, cudaMemcpyDeviceToHost));
    if (flags[0] == 0xdeadu || flags[1] == 0xdeadu) printf("TIMEOUT flags %x %x  ", flags[0], flags[1]);
    const int per = (mode == 2 || mode == 4) ? 4 : 1;
    printf("mode=%d gmap=%d delay=%4d cdelay=%4d work=%4d jitter=%4d : %8.1f cyc/round  net of work %8.1f\n", mode, gmap, delay, cdelay, work, jitter,
           (double)mx / p.iters, (double)mx / p.iters - per * (work + jitter / 2.0));
  };
  for (int d : {0, 400, 800, 1200}) run(3, 0, d, 0, 0, 0);
  for (int d : {400, 800}) run(3, 0, d, 0, 1000, 300);
  for (int d : {400, 800}) run(4, 0, d, 300, 1000, 0);
  for (int d : {400, 600}) run(2, 0, d, 300, 1000, 0);
  return 0;
  for (int d : {0, 200, 400, 600, 800}) run(0, 0, d, 0, 0, 0);
  for (int d : {400, 600, 800}) run(0, 0, d, 0, 1000, 300);
  for (int gm : {0, 1}) for (int d : {0, 200, 400, 600}) run(1, gm, 0, d, 0, 0);
  for (int gm : {0, 1}) for (int d : {200, 400}) run(1, gm, 0, d, 1000, 300);
  for (int gm : {0, 1}) for (int d : {400, 600}) for (int cd : {200, 400}) run(2, gm, d, cd, 1000, 0);
  for (int d : {400, 600}) run(2, 0, d, 300, 1000, 300);
  return 0;
}
