// Exchange micro-benchmark 4: instrumented rounds (where do the cycles go), publish flavours, private doorbells.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void ll_st(u64* p, unsigned payload, unsigned epoch, int kind) {
  u64 v = ((u64)epoch << 32) | payload;
  if (kind == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  else { u64 o; asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(o) : "l"(p), "l"(v) : "memory"); }
}
__device__ __forceinline__ u64 ll_ld(const u64* p) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ ulonglong2 ll_ld2(const u64* p) {
  ulonglong2 v; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ld4(const unsigned* p) {
  uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }

struct P { u64* buf; unsigned* bell; int words, iters, delay, kind, mode, work; long long* out; long long* prof; unsigned* sink; };
// mode 0: fixed delay then LL gather.  mode 1: private doorbells (bell[consumer][producer] = epoch), warp 0 polls them, then LL gather.
// mode 2: like 1 but the doorbell poll starts after `delay`.

__global__ void __launch_bounds__(448, 1) k_x(P p) {
  extern __shared__ unsigned s[];
  __shared__ unsigned s_retry; __shared__ long long s_t2;
  const int G = gridDim.x, cta = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  const int words = p.words;
  const int w0 = (int)((long long)cta * words / G), w1 = (int)((long long)(cta + 1) * words / G);
  const int nmine = w1 - w0;
  unsigned acc = 0;
  long long sum_pub = 0, sum_wait = 0, sum_first = 0, sum_all = 0, sum_sync = 0; unsigned retries = 0;
  if (tid == 0) { s_retry = 0; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 1; it <= p.iters; ++it) {
    u64* buf = p.buf + (size_t)(it & 1) * 4096;
    long long ta = clock64();
    if (tid < nmine) ll_st(buf + w0 + tid, (unsigned)(it + tid), (unsigned)it, p.kind);
    if (p.mode >= 1 && tid >= 64 && tid < 64 + G) {
      // doorbell: tell consumer (tid-64) that producer `cta` has published round it
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p.bell + (size_t)(tid - 64) * 160 + cta), "r"((unsigned)it) : "memory");
    }
    long long tb = clock64();
    if (p.delay > 0) { while (clock64() - tb < p.delay) {} }
    long long tc = clock64();
    if (p.mode >= 1) {
      if (tid < 32) {
        const unsigned* mybell = p.bell + (size_t)cta * 160;
        for (;;) {
          bool ok = true;
          for (int j = tid * 4; j < G; j += 128) {
            uint4 v = ld4(mybell + j);
            ok = ok && v.x == (unsigned)it && (j + 1 >= G || v.y == (unsigned)it) && (j + 2 >= G || v.z == (unsigned)it) && (j + 3 >= G || v.w == (unsigned)it);
          }
          if (__all_sync(0xffffffffu, ok)) break;
        }
      }
      __syncthreads();
    }
    long long td = clock64();
    long long tfirst = 0;
    for (int i = tid * 2; i < words; i += T * 2) {
      ulonglong2 w = ll_ld2(buf + i);
      if (tfirst == 0) tfirst = clock64();
      while ((unsigned)(w.x >> 32) != (unsigned)it || (unsigned)(w.y >> 32) != (unsigned)it) { w = ll_ld2(buf + i); ++retries; }
      s[i] = (unsigned)w.x; s[i + 1] = (unsigned)w.y;
    }
    long long te = clock64();
    __syncthreads();
    long long tf = clock64();
    acc += s[(tid * 7 + it) % words];
    if (p.work > 0) { long long t = clock64(); while (clock64() - t < p.work) {} }
    __syncthreads();
    if (tid == 0) { sum_pub += tb - ta; sum_wait += td - tc; sum_first += tfirst - td; sum_all += te - td; sum_sync += tf - te; }
  }
  if (retries) atomicAdd(&s_retry, retries);
  __syncthreads();
  if (tid == 0) {
    p.out[cta] = clock64() - t0;
    long long* pr = p.prof + cta * 8;
    pr[0] = sum_pub; pr[1] = sum_wait; pr[2] = sum_first; pr[3] = sum_all; pr[4] = sum_sync; pr[5] = s_retry;
  }
  p.sink[cta * T + tid] = acc;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int G = prop.multiProcessorCount;
  P p; CK(cudaMalloc(&p.out, 1024 * 8)); CK(cudaMalloc(&p.sink, 1 << 22)); CK(cudaMalloc(&p.prof, 1024 * 8 * 8));
  size_t bytes = (size_t)2 * 4096 * 8; CK(cudaMalloc(&p.buf, bytes)); CK(cudaMalloc(&p.bell, 160 * 160 * 4));
  CK(cudaFuncSetAttribute(k_x, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  auto run = [&](int words, int mode, int kind, int delay, int work) {
    p.words = words; p.iters = 2000; p.delay = delay; p.kind = kind; p.mode = mode; p.work = work;
    CK(cudaMemset(p.buf, 0, bytes)); CK(cudaMemset(p.bell, 0, 160 * 160 * 4));
    void* args[] = {&p};
    CK(cudaLaunchCooperativeKernel((void*)k_x, dim3(G), dim3(448), args, 200 * 1024, 0));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(G), pr(G * 8); CK(cudaMemcpy(h.data(), p.out, G * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pr.data(), p.prof, G * 64, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    double a[6] = {0};
    for (int c = 0; c < G; ++c) for (int k = 0; k < 6; ++k) a[k] += (double)pr[c * 8 + k] / G / p.iters;
    printf("words=%4d mode=%d kind=%d delay=%4d work=%4d : %7.1f cyc/round (-delay-work %7.1f) | pub %5.0f bellwait %6.0f first-ld %6.0f gather %6.0f sync %5.0f retries/round/cta %.2f\n",
           words, mode, kind, delay, work, (double)mx / p.iters, (double)mx / p.iters - delay - work, a[0], a[1], a[2], a[3], a[4], a[5]);
  };
  for (int d : {350, 400, 450, 500, 550, 600, 700}) run(512, 0, 0, d, 0);
  for (int d : {400, 500, 600}) run(512, 0, 1, d, 0);
  for (int d : {500, 600, 800}) run(1536, 0, 0, d, 0);
  run(512, 1, 0, 0, 0);
  run(512, 1, 1, 0, 0);
  for (int d : {300, 500}) run(512, 2, 0, d, 0);
  run(1536, 1, 0, 0, 0);
  run(512, 1, 0, 0, 2000);
  run(512, 0, 0, 600, 2000);
  return 0;
}
