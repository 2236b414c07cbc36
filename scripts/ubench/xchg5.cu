// Exchange micro-benchmark 5: counter doorbell + LL gather (replicas), with per-CTA jitter to emulate skew.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void ll_st(u64* p, unsigned payload, unsigned epoch) {
  u64 v = ((u64)epoch << 32) | payload; asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ ulonglong2 ll_ld2(const u64* p) {
  ulonglong2 v; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_u32(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

struct P { u64* buf; unsigned* ctr; int words, R, iters, mode, delay, backoff, jitter, work; long long* out; unsigned* sink; };
// mode 0: fixed delay after own publish, then gather.   mode 1: counter doorbell (thread 0 polls, backoff ns), then gather.
// mode 2: counter doorbell polled by thread 0 after an initial `delay`.

__global__ void __launch_bounds__(448, 1) k_x(P p) {
  extern __shared__ unsigned s[];
  const int G = gridDim.x, cta = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  const int words = p.words, R = p.R;
  const int w0 = (int)((long long)cta * words / G), w1 = (int)((long long)(cta + 1) * words / G);
  const int nmine = w1 - w0;
  unsigned acc = 0, rng = cta * 2654435761u + 12345u;
  long long t0 = clock64();
  for (int it = 1; it <= p.iters; ++it) {
    u64* buf = p.buf + (size_t)(it & 1) * 8 * 4096;
    const u64* rd = buf + (size_t)(cta % R) * words;
    // emulated phase work with per-CTA jitter
    rng = rng * 1664525u + 1013904223u;
    const int w = p.work + (p.jitter ? (int)((rng >> 8) % (unsigned)p.jitter) : 0);
    if (w > 0) { long long t = clock64(); while (clock64() - t < w) {} }
    for (int i = tid; i < nmine * R; i += T) {
      const int v = i / R, r = i % R;
      ll_st(buf + (size_t)r * words + w0 + v, (unsigned)(it + v), (unsigned)it);
    }
    if (p.mode >= 1) {
      __syncthreads();  // all publishing threads have issued their stores
      if (tid == 0) {
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p.ctr) : "memory");
        if (p.delay > 0) { long long t = clock64(); while (clock64() - t < p.delay) {} }
        const unsigned want = (unsigned)G * (unsigned)it;
        if (p.mode == 3) {
          unsigned v;
          do { asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 0;" : "=r"(v) : "l"(p.ctr) : "memory"); if (p.backoff > 0 && (int)(v - want) < 0) __nanosleep(p.backoff); } while ((int)(v - want) < 0);
        } else if (p.mode == 4) {
          unsigned v;
          do { asm volatile("atom.relaxed.gpu.global.or.b32 %0, [%1], 0;" : "=r"(v) : "l"(p.ctr) : "memory"); } while ((int)(v - want) < 0);
        } else
        while ((int)(ld_u32(p.ctr) - want) < 0) { if (p.backoff > 0) __nanosleep(p.backoff); }
      }
      __syncthreads();
    } else if (p.delay > 0) { long long t = clock64(); while (clock64() - t < p.delay) {} }
    for (int i = tid * 2; i < words; i += T * 2) {
      ulonglong2 wv = ll_ld2(rd + i);
      while ((unsigned)(wv.x >> 32) != (unsigned)it || (unsigned)(wv.y >> 32) != (unsigned)it) wv = ll_ld2(rd + i);
      s[i] = (unsigned)wv.x; s[i + 1] = (unsigned)wv.y;
    }
    __syncthreads();
    acc += s[(tid * 7 + it) % words];
    __syncthreads();
  }
  if (tid == 0) p.out[cta] = clock64() - t0;
  p.sink[cta * T + tid] = acc;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int G = prop.multiProcessorCount;
  P p; CK(cudaMalloc(&p.out, 1024 * 8)); CK(cudaMalloc(&p.sink, 1 << 22));
  size_t bytes = (size_t)2 * 8 * 4096 * 8; CK(cudaMalloc(&p.buf, bytes)); CK(cudaMalloc(&p.ctr, 256));
  CK(cudaFuncSetAttribute(k_x, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  auto run = [&](int words, int R, int mode, int delay, int backoff, int jitter, int work) {
    p.words = words; p.R = R; p.iters = 1500; p.mode = mode; p.delay = delay; p.backoff = backoff; p.jitter = jitter; p.work = work;
    CK(cudaMemset(p.buf, 0, bytes)); CK(cudaMemset(p.ctr, 0, 256));
    void* args[] = {&p};
    CK(cudaLaunchCooperativeKernel((void*)k_x, dim3(G), dim3(448), args, 200 * 1024, 0));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(G); CK(cudaMemcpy(h.data(), p.out, G * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    // expected work per round: work + E[max jitter] ~ work + jitter
    printf("words=%4d R=%d mode=%d delay=%4d backoff=%3d jitter=%4d work=%4d : %7.1f cyc/round  (-work-jitter %7.1f)\n", words, R, mode, delay, backoff,
           jitter, work, (double)mx / p.iters, (double)mx / p.iters - work - jitter);
  };
  for (int m : {3, 4}) { run(512, 1, m, 0, 0, 0, 0); run(512, 1, m, 0, 50, 0, 0); run(512, 1, m, 200, 0, 0, 0); }
  run(342, 1, 3, 0, 0, 0, 0); run(1024, 1, 3, 0, 0, 0, 0); run(1536, 1, 3, 0, 0, 0, 0); run(512, 2, 3, 0, 0, 0, 0);
  for (int j : {300, 1000}) { run(512, 1, 3, 0, 0, j, 1000); run(512, 1, 3, 0, 50, j, 1000); run(1536, 1, 3, 0, 0, j, 1000); run(512, 1, 0, 800, 0, j, 1000); }
  return 0;
}
