// Cluster feasibility probe for the group kernel (round 2): can the 8 kv-head groups live in thread-block clusters?
//   1. cudaOccupancyMaxActiveClusters for cluster sizes 1..16 at the decode kernel's footprint (256 threads, 211 KB shared
//      memory, 1 CTA/SM) and at a half-size footprint (2 CTAs/SM).
//   2. which SMs the CTAs of each cluster land on (reveals the GPC sizes of this chip).
//   3. DSMEM "push" all-gather inside a cluster: every CTA stores its 32 LL4 words {epoch, payload} into every peer's shared
//      memory, then polls its own copy; cycles per round for cluster sizes 8..16 -- the DSMEM counterpart of the group
//      exchange that costs 1650-2500 cycles through L2.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster32(unsigned addr, unsigned v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }

__global__ void __launch_bounds__(256, 1) probe_kernel(int* out, int rounds, long long* cyc) {
  extern __shared__ __align__(16) unsigned char smem[];
  volatile unsigned* buf = reinterpret_cast<volatile unsigned*>(smem);   // [16][32] words
  const unsigned rank = cluster_rank(), cs = cluster_size();
  const int tid = threadIdx.x;
  if (tid == 0) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    out[blockIdx.x * 3 + 0] = (int)cluster_id(); out[blockIdx.x * 3 + 1] = (int)rank; out[blockIdx.x * 3 + 2] = (int)smid;
  }
  for (int i = tid; i < 16 * 32; i += blockDim.x) buf[i] = 0;
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  const unsigned base = smem_u32(smem);
  const long long t0 = clock64();
  unsigned bad = 0;
  for (int r = 1; r <= rounds; ++r) {
    // push: thread (peer = tid / 32 < cs, word = tid % 32) stores this CTA's word into the peer's buffer slot `rank`
    for (int peer = tid >> 5; peer < (int)cs; peer += 8)
      st_cluster32(mapa(base + (rank * 32 + (tid & 31)) * 4, peer), ((unsigned)r << 16) | (unsigned)(tid & 31));
    // poll the local copy: thread i < cs * 8 checks 4 words
    if (tid < (int)cs * 8) {
      unsigned spins = 0;
      for (;;) {
        uint4 w; asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(base + tid * 16) : "memory");
        if ((w.x >> 16) >= (unsigned)r && (w.y >> 16) >= (unsigned)r && (w.z >> 16) >= (unsigned)r && (w.w >> 16) >= (unsigned)r) break;
        if (++spins > (1u << 20)) { bad = 1; break; }
      }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = bad ? -1 : (t1 - t0);
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("SMs %d\n", prop.multiProcessorCount);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const size_t smems[2] = {216320, 110000};
  int maxc[2][17] = {};
  for (int si = 0; si < 2; ++si) {
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smems[si]));
    printf("smem %zu: max active clusters by cluster size:", smems[si]);
    for (int cs = 1; cs <= 16; ++cs) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 8); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smems[si];
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
      if (e != cudaSuccess) { n = -1; cudaGetLastError(); }
      maxc[si][cs] = n;
      printf(" %d:%d(%d CTAs)", cs, n, n * cs);
    }
    printf("\n");
  }
  int* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 4096 * 3 * sizeof(int)));
  CK(cudaMalloc(&d_cyc, 4096 * sizeof(long long)));
  for (int si = 0; si < 2; ++si) {
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smems[si]));
    for (int cs : {2, 4, 6, 8, 10, 12, 14, 16}) {
      int ncl = maxc[si][cs];
      if (ncl <= 0) { printf("smem %zu cluster size %d: not launchable\n", smems[si], cs); continue; }
      if (si == 1 && ncl > 8) ncl = 8;   // 2 CTAs/SM footprint: launch exactly 8 clusters to see how they are packed
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * ncl); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smems[si];
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      const int rounds = 200;
      CK(cudaMemset(d_out, 0xff, 4096 * 3 * sizeof(int)));
      cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel, d_out, rounds, d_cyc);
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("smem %zu cluster size %d x %d: launch failed: %s\n", smems[si], cs, ncl, cudaGetErrorString(e)); cudaGetLastError(); continue; }
      std::vector<int> h(cs * ncl * 3); std::vector<long long> c(cs * ncl);
      CK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(int), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(c.data(), d_cyc, c.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0, mn = 1LL << 62; int nbad = 0;
      for (auto v : c) { if (v < 0) { ++nbad; continue; } mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
      printf("smem %zu cluster size %2d x %2d clusters: DSMEM all-gather %lld..%lld cycles/round (bad %d)\n", smems[si], cs, ncl, mn / rounds, mx / rounds, nbad);
      std::vector<int> used(prop.multiProcessorCount, 0);
      for (int k = 0; k < ncl; ++k) {
        printf("   cluster %2d: SMs", k);
        for (int b = 0; b < cs * ncl; ++b) if (h[b * 3] == k) { printf(" %d", h[b * 3 + 2]); used[h[b * 3 + 2]]++; }
        printf("\n");
      }
      int distinct = 0, dbl = 0;
      for (int s = 0; s < prop.multiProcessorCount; ++s) { distinct += used[s] > 0; dbl += used[s] > 1; }
      printf("   distinct SMs %d, SMs hosting 2 CTAs %d\n", distinct, dbl);
    }
  }
  return 0;
}
