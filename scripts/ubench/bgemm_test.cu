// Self-checking test of the tcgen05 split-K decode GEMM (qmk_bgemm.cuh) against a CPU reference.
#include "../../qwen-megakernel-tts_b200/csrc/qmk_bgemm.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); return 1;} } while (0)
using namespace qmkb;
static float bf(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t tobf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fff + ((u >> 16) & 1); return (uint16_t)(u >> 16); }
int run(int M, int N, int K, int splits) {
  std::vector<uint16_t> W((size_t)M * K), X((size_t)N * K);
  srand(1234 + M + N + K);
  for (auto& w : W) w = tobf((rand() % 2001 - 1000) / 1000.0f);
  for (auto& x : X) x = tobf((rand() % 2001 - 1000) / 500.0f);
  void *dW, *dX; float* dP;
  CK(cudaMalloc(&dW, W.size() * 2)); CK(cudaMalloc(&dX, X.size() * 2)); CK(cudaMalloc(&dP, (size_t)splits * N * M * 4));
  CK(cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dP, 0xff, (size_t)splits * N * M * 4));
  CUtensorMap mw, mx;
  if (make_tensor_map(&mw, dW, M, K, BM) || make_tensor_map(&mx, dX, N, K, N)) { printf("tensor map failed\n"); return 1; }
  CK(cudaFuncSetAttribute(qmk_bgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  BgemmArgs a{dP, M, N, K, splits};
  qmk_bgemm_kernel<<<dim3(M / BM, splits), 128, SMEM_BYTES>>>(mw, mx, a);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  std::vector<float> P((size_t)splits * N * M);
  CK(cudaMemcpy(P.data(), dP, P.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int n = 0; n < N; ++n) for (int m = 0; m < M; ++m) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)bf(W[(size_t)m * K + k]) * bf(X[(size_t)n * K + k]);
    double got = 0; for (int s = 0; s < splits; ++s) got += P[((size_t)s * N + n) * M + m];
    maxerr = fmax(maxerr, fabs(got - ref)); maxref = fmax(maxref, fabs(ref));
  }
  printf("M=%d N=%d K=%d splits=%d: max abs err %.5f (max |ref| %.2f) %s\n", M, N, K, splits, maxerr, maxref, maxerr < 1e-2 * maxref ? "OK" : "MISMATCH");
  cudaFree(dW); cudaFree(dX); cudaFree(dP);
  return maxerr < 1e-2 * maxref ? 0 : 1;
}
int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int bad = 0;
  bad += run(128, 16, 64, 1);
  bad += run(128, 16, 128, 1);
  bad += run(256, 32, 256, 2);
  bad += run(1024, 64, 1024, 4);
  bad += run(4096, 64, 1024, 4);
  bad += run(1024, 16, 3072, 8);
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
