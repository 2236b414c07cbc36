// Exchange micro-benchmark 6: LL4 words {bf16 payload, u16 epoch}; fixed delay vs sentinel warp; jitter; u32 vs u64 ping-pong.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void st32(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld32(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ld128(const unsigned* p) {
  uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }

template <typename T>
__global__ void k_pp(T* a, T* b, int iters, int peer, long long* out) {
  if (threadIdx.x != 0) return;
  volatile T* va = a; volatile T* vb = b;
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) { *va = (T)i; while (*vb != (T)i) {} }
    out[0] = clock64() - t0;
  } else if (blockIdx.x == peer) {
    for (int i = 1; i <= iters; ++i) { while (*va != (T)i) {} *vb = (T)i; }
  }
}

struct P { unsigned* buf; int words, iters, mode, delay, backoff, jitter, work; long long* out; unsigned* sink; };
// mode 0: fixed delay, then gather (retry with backoff).  mode 1: sentinel warp (first poll after delay), bar, gather.

__device__ __forceinline__ bool ok4(uint4 w, unsigned e) {
  return (w.x >> 16) == e && (w.y >> 16) == e && (w.z >> 16) == e && (w.w >> 16) == e;
}

__global__ void __launch_bounds__(448, 1) k_x(P p) {
  extern __shared__ unsigned s[];
  const int G = gridDim.x, cta = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  const int words = p.words;
  const int w0 = (int)((long long)cta * words / G), w1 = (int)((long long)(cta + 1) * words / G);
  const int nmine = w1 - w0;
  unsigned acc = 0, rng = cta * 2654435761u + 12345u;
  long long t0 = clock64();
  for (int it = 1; it <= p.iters; ++it) {
    unsigned* buf = p.buf + (size_t)(it & 1) * 4096;
    const unsigned e = (unsigned)it & 0xffffu;
    rng = rng * 1664525u + 1013904223u;
    const int w = p.work + (p.jitter ? (int)((rng >> 8) % (unsigned)p.jitter) : 0);
    if (w > 0) { long long t = clock64(); while (clock64() - t < w) {} }
    if (tid < nmine) st32(buf + w0 + tid, (e << 16) | (unsigned)(tid & 0xffff));
    if (p.delay > 0) { long long t = clock64(); while (clock64() - t < p.delay) {} }
    if (p.mode == 1) {
      if (tid < 32) {
        const int i = (int)(((long long)tid * words) >> 5) + (words >> 6);
        while ((ld32(buf + i) >> 16) != e) { if (p.backoff > 0) __nanosleep(p.backoff); }
      }
      __syncthreads();
    }
    for (int i = tid * 4; i < words; i += T * 4) {
      uint4 wv = ld128(buf + i);
      while (!ok4(wv, e)) { if (p.backoff > 0) __nanosleep(p.backoff); wv = ld128(buf + i); }
      *reinterpret_cast<uint4*>(s + i) = wv;
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
  size_t bytes = (size_t)2 * 4096 * 4; CK(cudaMalloc(&p.buf, bytes));
  {
    void* ab; CK(cudaMalloc(&ab, 4096));
    for (int peer : {1, 2, 75, 140}) {
      int iters = 2000;
      CK(cudaMemset(ab, 0, 4096));
      unsigned* a = (unsigned*)ab; unsigned* b = a + 128;
      void* args[] = {&a, &b, &iters, &peer, &p.out};
      CK(cudaLaunchCooperativeKernel((void*)k_pp<unsigned>, dim3(G), dim3(32), args, 0, 0)); CK(cudaDeviceSynchronize());
      long long r; CK(cudaMemcpy(&r, p.out, 8, cudaMemcpyDeviceToHost));
      printf("pingpong u32 peer=%3d one-way %.1f   ", peer, (double)r / iters / 2);
      CK(cudaMemset(ab, 0, 4096));
      u64* a8 = (u64*)ab; u64* b8 = a8 + 64;
      void* args8[] = {&a8, &b8, &iters, &peer, &p.out};
      CK(cudaLaunchCooperativeKernel((void*)k_pp<u64>, dim3(G), dim3(32), args8, 0, 0)); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&r, p.out, 8, cudaMemcpyDeviceToHost));
      printf("u64 one-way %.1f\n", (double)r / iters / 2);
    }
  }
  CK(cudaFuncSetAttribute(k_x, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  auto run = [&](int words, int mode, int delay, int backoff, int jitter, int work) {
    p.words = words; p.iters = 1500; p.mode = mode; p.delay = delay; p.backoff = backoff; p.jitter = jitter; p.work = work;
    CK(cudaMemset(p.buf, 0, bytes));
    void* args[] = {&p};
    CK(cudaLaunchCooperativeKernel((void*)k_x, dim3(G), dim3(448), args, 200 * 1024, 0));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(G); CK(cudaMemcpy(h.data(), p.out, G * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    printf("LL4 words=%4d mode=%d delay=%4d backoff=%3d jitter=%4d work=%4d : %7.1f cyc/round  (-work-jitter %7.1f)\n", words, mode, delay, backoff,
           jitter, work, (double)mx / p.iters, (double)mx / p.iters - work - jitter);
  };
  for (int d : {0, 200, 300, 400, 500, 700}) run(1024, 0, d, 0, 0, 0);
  for (int d : {300, 400, 600}) run(3072, 0, d, 0, 0, 0);
  for (int d : {0, 200, 400}) run(1024, 1, d, 0, 0, 0);
  for (int d : {0, 200, 400}) run(1024, 1, d, 100, 0, 0);
  for (int j : {300, 1000}) {
    for (int d : {400, 800}) { run(1024, 0, d, 0, j, 1000); run(1024, 0, d, 100, j, 1000); }
    for (int d : {200, 400}) { run(1024, 1, d, 0, j, 1000); run(1024, 1, d, 100, j, 1000); run(1024, 1, d, 300, j, 1000); }
    run(3072, 0, 600, 100, j, 1000); run(3072, 1, 300, 100, j, 1000);
  }
  return 0;
}
