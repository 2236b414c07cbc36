// Back-to-back launch cost of a 148 x 256 kernel with 208 KB dynamic smem and a ~3.5 KB parameter block.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1;} } while (0)
struct Big { int v[880]; };
struct Small { int v[8]; };
template <typename P> __global__ void __launch_bounds__(256, 1) k_empty(const __grid_constant__ P p, int* out) {
  extern __shared__ unsigned char sm[];
  if (threadIdx.x == 0 && p.v[0] == 12345) { sm[0] = 1; out[blockIdx.x] = sm[0]; }
}
template <typename P> __global__ void __launch_bounds__(256, 1) k_spin(const __grid_constant__ P p, int* out, int cycles) {
  extern __shared__ unsigned char sm[];
  long long t = clock64(); while (clock64() - t < cycles) {}
  if (threadIdx.x == 0 && p.v[0] == 12345) { sm[0] = 1; out[blockIdx.x] = sm[0]; }
}
template <typename P> int run(const char* name, int smem, bool coop, int spin) {
  int* out; CK(cudaMalloc(&out, 4096));
  P p{}; 
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  void* fn = spin ? (void*)k_spin<P> : (void*)k_empty<P>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  void* args[] = {&p, &out, &spin};
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(a));
    for (int i = 0; i < 200; ++i) {
      if (coop) CK(cudaLaunchCooperativeKernel(fn, dim3(148), dim3(256), args, smem, 0));
      else CK(cudaLaunchKernel(fn, dim3(148), dim3(256), args, smem, 0));
    }
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  }
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  printf("%-10s smem=%6d coop=%d spin=%6d cycles: %.2f us per launch\n", name, smem, (int)coop, spin, ms * 1000 / 200);
  return 0;
}
int main() {
  run<Small>("small", 0, false, 0); run<Small>("small", 208 * 1024, false, 0); run<Big>("big", 208 * 1024, false, 0);
  run<Big>("big", 208 * 1024, true, 0);
  run<Big>("big", 208 * 1024, false, 200000); run<Big>("big", 208 * 1024, true, 200000);
  run<Small>("small", 208 * 1024, false, 200000); run<Small>("small", 0, false, 200000);
  return 0;
}
