// Exchange micro-benchmark 3: LL8 words gathered with 16-byte loads; sweeps poll delay / backoff / replicas / #gatherers.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void ll_st(u64* p, unsigned payload, unsigned epoch) {
  u64 v = ((u64)epoch << 32) | payload; asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ll_ld(const u64* p) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ ulonglong2 ll_ld2(const u64* p) {
  ulonglong2 v; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v; }

struct P { u64* buf; int words, R, iters, delay, backoff, ngather, work; long long* out; unsigned* sink; };

__global__ void __launch_bounds__(448, 1) k_x(P p) {
  extern __shared__ unsigned s[];
  const int G = gridDim.x, cta = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  const int words = p.words, R = p.R;
  const int w0 = (int)((long long)cta * words / G), w1 = (int)((long long)(cta + 1) * words / G);
  const int nmine = w1 - w0;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 1; it <= p.iters; ++it) {
    u64* buf = p.buf + (size_t)(it & 1) * 16 * 4096;
    const u64* rd = buf + (size_t)(cta % R) * words;
    for (int i = tid; i < nmine * R; i += T) {
      const int v = i / R, r = i % R;
      ll_st(buf + (size_t)r * words + w0 + v, (unsigned)(it + v), (unsigned)it);
    }
    if (p.delay > 0) { long long t = clock64(); while (clock64() - t < p.delay) {} }
    if (cta < p.ngather) {
      for (int i = tid * 2; i < words; i += T * 2) {
        ulonglong2 w = ll_ld2(rd + i);
        while ((unsigned)(w.x >> 32) != (unsigned)it || (unsigned)(w.y >> 32) != (unsigned)it) {
          if (p.backoff > 0) __nanosleep(p.backoff);
          w = ll_ld2(rd + i);
        }
        s[i] = (unsigned)w.x; s[i + 1] = (unsigned)w.y;
      }
    } else if (tid < 32) {
      // light participant: waits for one word of each of 32 evenly spaced producers
      int i = (int)(((long long)tid * words) >> 5);
      while ((unsigned)(ll_ld(rd + i) >> 32) != (unsigned)it) { if (p.backoff > 0) __nanosleep(p.backoff); }
    }
    __syncthreads();
    acc += s[(tid * 7 + it) % words];
    if (p.work > 0) { long long t = clock64(); while (clock64() - t < p.work) {} }
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
  size_t bytes = (size_t)2 * 16 * 4096 * 8; CK(cudaMalloc(&p.buf, bytes));
  CK(cudaFuncSetAttribute(k_x, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  auto run = [&](int words, int R, int delay, int backoff, int ngather, int work) {
    p.words = words; p.R = R; p.iters = 2000; p.delay = delay; p.backoff = backoff; p.ngather = ngather; p.work = work;
    CK(cudaMemset(p.buf, 0, bytes));
    void* args[] = {&p};
    CK(cudaLaunchCooperativeKernel((void*)k_x, dim3(G), dim3(448), args, 200 * 1024, 0));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(G); CK(cudaMemcpy(h.data(), p.out, G * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    printf("words=%4d R=%2d delay=%4d backoff=%3d ngather=%3d work=%4d : %7.1f cycles/round (minus delay+work: %7.1f)\n", words, R, delay, backoff,
           ngather, work, (double)mx / p.iters, (double)mx / p.iters - delay - work);
  };
  for (int ng : {0, 8, 32, 74, 148}) run(512, 1, 0, 0, ng, 0);
  for (int ng : {8, 148}) run(1536, 1, 0, 0, ng, 0);
  for (int R : {2, 4, 8}) run(512, R, 0, 0, 148, 0);
  for (int d : {300, 600, 900, 1200}) run(512, 1, d, 0, 148, 0);
  for (int b : {20, 100, 300}) run(512, 1, 0, b, 148, 0);
  for (int d : {600, 900}) run(512, 4, d, 100, 148, 0);
  for (int w : {1000, 3000}) run(512, 1, 0, 0, 148, w);
  for (int w : {1000, 3000}) run(512, 1, 600, 100, 148, w);
  run(1536, 1, 600, 100, 148, 0);
  run(1536, 4, 600, 100, 148, 0);
  return 0;
}
