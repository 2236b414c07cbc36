// Throughput / latency of legacy mma.sync (HMMA.16816 bf16 -> f32), ldmatrix and packed fma.f32x2 on sm_100a.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1;} } while (0)

__device__ __forceinline__ void mma16816(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm4(unsigned (&r)[4], unsigned addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// mode 0: dependent chain of mma (latency); mode 1: 4 independent accumulators (throughput); mode 2: ldmatrix + mma (4 acc)
__global__ void k_mma(int iters, int mode, long long* out, float* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(sm)[i] = 0x3f803f80u;
  __syncthreads();
  unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
  float d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0}, d3[4] = {0, 0, 0, 0};
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 2048;
  long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) { mma16816(d0, a, b); mma16816(d0, a, b); mma16816(d0, a, b); mma16816(d0, a, b); }
  } else if (mode == 1) {
    for (int i = 0; i < iters; ++i) { mma16816(d0, a, b); mma16816(d1, a, b); mma16816(d2, a, b); mma16816(d3, a, b); }
  } else {
    for (int i = 0; i < iters; ++i) {
      unsigned a0[4], a1[4], a2[4], a3[4];
      ldsm4(a0, base); ldsm4(a1, base + 512); ldsm4(a2, base + 1024); ldsm4(a3, base + 1536);
      mma16816(d0, a0, b); mma16816(d1, a1, b); mma16816(d2, a2, b); mma16816(d3, a3, b);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = d0[0] + d1[1] + d2[2] + d3[3];
}

// packed fp32x2 FMA vs scalar FMA throughput (8 independent chains per thread)
__global__ void k_fma(int iters, int mode, long long* out, float* sink) {
  float acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  const float m = 1.0001f, c = 0.5f;
  long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = fmaf(acc[j], m, c);
    }
  } else {
    unsigned long long mm, cc;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    unsigned long long p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[j]) : "f"(acc[2 * j]), "f"(acc[2 * j + 1]));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(mm), "l"(cc));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * j]), "=f"(acc[2 * j + 1]) : "l"(p[j]));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  long long* out; CK(cudaMalloc(&out, 4096)); float* sink; CK(cudaMalloc(&sink, 1 << 22));
  CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int iters = 2000;
  for (int warps : {1, 4, 8, 14, 16}) {
    for (int mode : {0, 1, 2}) {
      k_mma<<<148, warps * 32, 65536>>>(iters, mode, out, sink); CK(cudaDeviceSynchronize());
      long long r; CK(cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost));
      printf("mma warps=%2d mode=%d: %.2f cycles per mma per warp; SM-wide %.2f cycles/mma\n", warps, mode, (double)r / (iters * 4), (double)r / (iters * 4) / warps);
    }
  }
  for (int warps : {4, 16}) {
    for (int mode : {0, 1}) {
      k_fma<<<148, warps * 32>>>(iters, mode, out, sink); CK(cudaDeviceSynchronize());
      long long r; CK(cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost));
      printf("fma warps=%2d %s: %.2f cycles per 16 fp32 FMAs per warp\n", warps, mode ? "f32x2 " : "scalar", (double)r / iters);
    }
  }
  return 0;
}
