// How long does one GEMV stage (B-fragment loads + 8 ldmatrix + 8 HMMA per warp, 8 warps) take in isolation,
// after an idle gap (like the exchange wait) and with / without straight-line padding code in between?
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1;} } while (0)
__device__ __forceinline__ void mma16816(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm4(unsigned (&r)[4], unsigned addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__global__ void __launch_bounds__(256, 1) k_stage(int iters, int gap, int nst, long long* out, float* sink) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 100 * 1024 / 4; i += 256) reinterpret_cast<unsigned*>(sm)[i] = 0x3c003c00u + i;
  __syncthreads();
  const int mi = lane >> 3; int arow = (lane & 7) + (mi & 1) * 8; if (arow >= 14) arow -= 8;
  const unsigned a_off = 2560 + arow * 2048; const int sw = arow & 7, kh = mi >> 1;
  const unsigned ring = (unsigned)__cvta_generic_to_shared(sm);
  const uint4* xs = reinterpret_cast<const uint4*>(sm + 96 * 1024 + warp * 256 + (lane & 3) * 64);
  float* s_part = reinterpret_cast<float*>(sm + 98 * 1024);
  long long t_stage = 0, t_red = 0;
  float accum = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (gap > 0) { long long t = clock64(); while (clock64() - t < gap) {} }
    __syncthreads();
    long long t0 = clock64();
    unsigned bfrag[8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const uint4 v = xs[i]; bfrag[2*i][0] = v.x; bfrag[2*i][1] = v.y; bfrag[2*i+1][0] = v.z; bfrag[2*i+1][1] = v.w; }
    float acc[3][4];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.f;
      if (s < nst) {
        float acc2[4] = {0, 0, 0, 0};
        const unsigned base = ring + s * 31232 + a_off;
#pragma unroll
        for (int jb = 0; jb < 8; jb += 4) {
          unsigned af[4][4];
#pragma unroll
          for (int j = 0; j < 4; ++j) ldsm4(af[j], base + ((unsigned)((((warp * 8 + jb + j) * 2 + kh) ^ sw)) << 4));
          mma16816(acc[s], af[0], bfrag[jb]); mma16816(acc2, af[1], bfrag[jb + 1]);
          mma16816(acc[s], af[2], bfrag[jb + 2]); mma16816(acc2, af[3], bfrag[jb + 3]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[s][e] += acc2[e];
      }
    }
    long long t1 = clock64();
#pragma unroll
    for (int s = 0; s < 3; ++s) if (s < nst && (lane & 3) == 0) { s_part[((s * 14 + (lane >> 2)) * 8 + warp)] = acc[s][0]; if ((lane >> 2) < 6) s_part[((s * 14 + (lane >> 2) + 8) * 8 + warp)] = acc[s][2]; }
    __syncthreads();
    if (tid < 28) { const float4* v = reinterpret_cast<const float4*>(s_part + tid * 8); float4 a = v[0], b = v[1]; accum += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)); }
    long long t2 = clock64();
    if (tid == 0) { t_stage += t1 - t0; t_red += t2 - t1; }
  }
  if (tid == 0) { out[0] = t_stage; out[1] = t_red; }
  sink[blockIdx.x * 256 + tid] = accum;
}
int main() {
  long long* out; CK(cudaMalloc(&out, 64)); float* sink; CK(cudaMalloc(&sink, 1 << 20));
  CK(cudaFuncSetAttribute(k_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int nst : {1, 2, 3}) for (int gap : {0, 2000}) {
    int iters = 500;
    k_stage<<<148, 256, 200 * 1024>>>(iters, gap, nst, out, sink); CK(cudaDeviceSynchronize());
    long long r[2]; CK(cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost));
    printf("nst=%d gap=%4d: stages %.0f cycles, partial+bar+sum %.0f cycles\n", nst, gap, (double)r[0] / iters, (double)r[1] / iters);
  }
  return 0;
}
