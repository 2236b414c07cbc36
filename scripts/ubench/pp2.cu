#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1;} } while (0)
template <int ST, int LD>
__global__ void k_pp(unsigned* a, unsigned* b, int iters, int peer, long long* out) {
  if (threadIdx.x != 0) return;
  auto st = [](unsigned* p, unsigned v) {
    if (ST == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 1) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
  };
  auto ld = [](unsigned* p) {
    unsigned v;
    if (LD == 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (LD == 1) asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 0;" : "=r"(v) : "l"(p) : "memory");
    if (LD == 2) asm volatile("atom.relaxed.gpu.global.or.b32 %0, [%1], 0;" : "=r"(v) : "l"(p) : "memory");
    if (LD == 3) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (LD == 4) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
  };
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) { st(a, (unsigned)i); while (ld(b) != (unsigned)i) {} }
    out[0] = clock64() - t0;
  } else if (blockIdx.x == peer) {
    for (int i = 1; i <= iters; ++i) { while (ld(a) != (unsigned)i) {} st(b, (unsigned)i); }
  }
}
template <int ST, int LD> int pp(unsigned* ab, long long* out, const char* name) {
  for (int peer : {1, 2, 75}) {
    CK(cudaMemset(ab, 0, 4096));
    int iters = 2000; unsigned* a = ab; unsigned* b = ab + 128;
    void* args[] = {&a, &b, &iters, &peer, &out};
    CK(cudaLaunchCooperativeKernel((void*)k_pp<ST, LD>, dim3(148), dim3(32), args, 0, 0));
    CK(cudaDeviceSynchronize());
    long long r; CK(cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost));
    printf("pingpong %-28s peer=%2d: one-way %.1f cycles\n", name, peer, (double)r / iters / 2);
  }
  return 0;
}
int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  long long* out; CK(cudaMalloc(&out, 64)); unsigned* ab; CK(cudaMalloc(&ab, 4096));
  pp<0, 0>(ab, out, "st / ld.relaxed.gpu");
  pp<0, 1>(ab, out, "st / atom.add 0");
  pp<1, 1>(ab, out, "red.add / atom.add 0");
  pp<0, 2>(ab, out, "st / atom.or 0");
  pp<0, 3>(ab, out, "st / ld.cv");
  pp<0, 4>(ab, out, "st / ld.relaxed.sys");
  return 0;
}
