// Micro-benchmarks for the inter-CTA exchange primitive (LL words through L2) on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg xchg.cu && ./xchg
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;
typedef unsigned long long u64;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void ll_st(u64* p, unsigned payload, unsigned epoch) {
  u64 v = ((u64)epoch << 32) | payload;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ll_ld(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ ulonglong2 ll_ld2(const u64* p) {
  ulonglong2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}

// 1) L2 latency: dependent chain of ld.relaxed.gpu over a buffer of n u64 indices
__global__ void k_chase(const u64* buf, int iters, long long* out) {
  u64 idx = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) idx = ll_ld(buf + idx);
  long long t1 = clock64();
  out[0] = t1 - t0;
  out[1] = (long long)idx;
}

// 2) ping-pong between CTA 0 and CTA `peer`
__global__ void k_pingpong(u64* a, u64* b, int iters, int peer, long long* out) {
  if (threadIdx.x != 0) return;
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
      ll_st(a, 0, i);
      while ((unsigned)(ll_ld(b) >> 32) != (unsigned)i) {}
    }
    out[0] = clock64() - t0;
  } else if (blockIdx.x == peer) {
    for (int i = 1; i <= iters; ++i) {
      while ((unsigned)(ll_ld(a) >> 32) != (unsigned)i) {}
      ll_st(b, 0, i);
    }
  }
}

// 3) all-to-all exchange rounds: every CTA publishes its share of `words` words into R replicas and
//    gathers all `words` words from replica (cta % R).
template <int MODE>
__global__ void k_xchg(u64* buf, int words, int R, int iters, int delay, long long* out, unsigned* sink) {
  extern __shared__ unsigned s[];
  const int G = gridDim.x, cta = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  const int w0 = (int)((long long)cta * words / G), w1 = (int)((long long)(cta + 1) * words / G);
  const int nmine = w1 - w0;
  unsigned acc = 0;
  long long t0 = clock64();
  u64* const buf0 = buf;
  for (int it = 1; it <= iters; ++it) {
    buf = buf0 + (size_t)(it & 1) * 32 * 4096;   // double-buffered: round it+2 reuses the words of round it
    const u64* rd = buf + (size_t)(cta % R) * words;
    // publish: (value v, replica r) pairs spread over the threads
    for (int i = tid; i < nmine * R; i += T) {
      const int v = i / R, r = i % R;
      ll_st(buf + (size_t)r * words + w0 + v, (unsigned)(it + v), (unsigned)it);
    }
    // gather
    if (MODE == 0) {
      for (int base = 0; base < words; base += T * 4) {
        u64 w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { int i = base + u * T + tid; if (i < words) w[u] = ll_ld(rd + i); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          int i = base + u * T + tid;
          if (i < words) {
            while ((unsigned)(w[u] >> 32) != (unsigned)it) w[u] = ll_ld(rd + i);
            s[i] = (unsigned)w[u];
          }
        }
      }
    } else if (MODE == 1) {  // 16-byte loads: two words per access
      for (int base = 0; base < words; base += T * 4) {
        ulonglong2 w[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) { int i = base + (u * T + tid) * 2; if (i < words) w[u] = ll_ld2(rd + i); }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          int i = base + (u * T + tid) * 2;
          if (i < words) {
            while ((unsigned)(w[u].x >> 32) != (unsigned)it || (unsigned)(w[u].y >> 32) != (unsigned)it) w[u] = ll_ld2(rd + i);
            s[i] = (unsigned)w[u].x; s[i + 1] = (unsigned)w[u].y;
          }
        }
      }
    } else {  // MODE 2: warp 0 probes 32 sample words, then everyone loads
      if (tid < 32) {
        int i = (int)(((long long)tid * words) >> 5) + (words >> 6);
        while ((unsigned)(ll_ld(rd + i) >> 32) != (unsigned)it) {}
      }
      __syncthreads();
      for (int base = 0; base < words; base += T * 4) {
        u64 w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { int i = base + u * T + tid; if (i < words) w[u] = ll_ld(rd + i); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          int i = base + u * T + tid;
          if (i < words) {
            while ((unsigned)(w[u] >> 32) != (unsigned)it) w[u] = ll_ld(rd + i);
            s[i] = (unsigned)w[u];
          }
        }
      }
    }
    __syncthreads();
    acc += s[(tid * 7 + it) % words];
    if (delay > 0) { long long t = clock64(); while (clock64() - t < delay) {} }
    __syncthreads();
  }
  if (tid == 0) out[cta] = clock64() - t0;
  sink[cta * T + tid] = acc;
}

// 4) cooperative grid.sync for reference
__global__ void k_gridsync(int iters, long long* out) {
  cg::grid_group g = cg::this_grid();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) g.sync();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

template <int MODE>
static void run_xchg(u64* buf, size_t buf_bytes, int G, int T, int words, int R, int iters, int delay, long long* out, unsigned* sink) {
  CK(cudaMemset(buf, 0, buf_bytes));
  void* args[] = {&buf, &words, &R, &iters, &delay, &out, &sink};
  size_t smem = (size_t)words * 4 + 16;
  CK(cudaFuncSetAttribute(k_xchg<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaLaunchCooperativeKernel((void*)k_xchg<MODE>, dim3(G), dim3(T), args, smem, 0));
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(G);
  CK(cudaMemcpy(h.data(), out, G * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  printf("xchg mode=%d words=%5d R=%2d delay=%4d : %8.1f cycles/round\n", MODE, words, R, delay, (double)mx / iters - delay);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int G = prop.multiProcessorCount;
  printf("%s, %d SMs, clock %d kHz\n", prop.name, G, prop.clockRate);
  long long* out; CK(cudaMalloc(&out, 1024 * sizeof(long long)));
  unsigned* sink; CK(cudaMalloc(&sink, 1024 * 1024 * 4));
  // chase
  for (int n : {1 << 10, 1 << 17, 1 << 21}) {  // 8 KB, 1 MB, 16 MB of u64
    std::vector<u64> h(n);
    // stride permutation: idx -> (idx + 4099*16+1) % n  (jump > 128 B lines)
    for (int i = 0; i < n; ++i) h[i] = ((u64)i * 1 + 65537) % n;
    u64* d; CK(cudaMalloc(&d, n * 8)); CK(cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice));
    k_chase<<<1, 1>>>(d, 2000, out); CK(cudaDeviceSynchronize());
    k_chase<<<1, 1>>>(d, 2000, out); CK(cudaDeviceSynchronize());
    long long r[2]; CK(cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost));
    printf("chase n=%8d (%6.1f KB): %7.1f cycles/load\n", n, n * 8 / 1024.0, (double)r[0] / 2000);
    cudaFree(d);
  }
  // ping-pong
  {
    u64* ab; CK(cudaMalloc(&ab, 4096)); 
    for (int peer : {1, 2, 37, 74, 100, 147}) {
      CK(cudaMemset(ab, 0, 4096));
      int iters = 2000;
      u64* a = ab; u64* b = ab + 64;
      void* args[] = {&a, &b, &iters, &peer, &out};
      CK(cudaLaunchCooperativeKernel((void*)k_pingpong, dim3(G), dim3(32), args, 0, 0));
      CK(cudaDeviceSynchronize());
      long long r; CK(cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost));
      printf("pingpong peer=%3d: %7.1f cycles/round-trip (one-way %.1f)\n", peer, (double)r / iters, (double)r / iters / 2);
    }
  }
  // grid.sync
  {
    int iters = 1000;
    void* args[] = {&iters, &out};
    CK(cudaLaunchCooperativeKernel((void*)k_gridsync, dim3(G), dim3(448), args, 0, 0));
    CK(cudaDeviceSynchronize());
    long long r; CK(cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost));
    printf("grid.sync (148 x 448): %7.1f cycles\n", (double)r / iters);
  }
  size_t buf_bytes = (size_t)2 * 32 * 4096 * 8;
  u64* buf; CK(cudaMalloc(&buf, buf_bytes));
  const int T = 448, iters = 2000;
  for (int words : {1024, 3072}) {
    for (int R : {1, 2, 4, 8, 16, 32}) {
      run_xchg<0>(buf, buf_bytes, G, T, words, R, iters, 0, out, sink);
    }
    for (int R : {1, 4, 16}) run_xchg<1>(buf, buf_bytes, G, T, words, R, iters, 0, out, sink);
    for (int R : {1, 4, 16}) run_xchg<2>(buf, buf_bytes, G, T, words, R, iters, 0, out, sink);
    for (int R : {1, 4, 16}) run_xchg<0>(buf, buf_bytes, G, T, words, R, iters, 2000, out, sink);
  }
  return 0;
}
