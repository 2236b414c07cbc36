// Cost of a kernel boundary in a dependent chain (the batched launch chain's dominant term): N kernels, each reads the word its
// predecessor wrote and writes the next one.  Variants: plain stream order; programmatic dependent launch with the trigger at
// the start / at the end of the kernel; chains captured in a CUDA graph.  Grid / block / shared memory are the chain's.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_link(const unsigned* in, unsigned* out, int trigger_early, int work) {
  extern __shared__ unsigned char sm[];
  if (trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned v = __ldcg(in + threadIdx.x);
  if (work > 0) { long long t = clock64(); while (clock64() - t < work) {} }
  out[threadIdx.x] = v + 1;
  if (!trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

static void launch(cudaStream_t st, dim3 grid, int threads, size_t smem, bool pdl, const unsigned* in, unsigned* out, int early, int work) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, k_link, in, out, early, work));
}

int main() {
  unsigned* buf; CK(cudaMalloc(&buf, 2 * 4096)); CK(cudaMemset(buf, 0, 2 * 4096));
  CK(cudaFuncSetAttribute(k_link, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const int N = 400;
  struct Cfg { const char* name; int grid, threads; size_t smem; } cfgs[] = {
      {"128 x 128, 97 KB (GEMM)", 128, 128, 97 * 1024}, {"16 x 256 (epilogue B=16)", 16, 256, 0}, {"512 x 256 (attention B=64)", 512, 256, 0},
      {"alternating GEMM / epilogue", -1, 0, 0},
      {"512 x 128", 512, 128, 0}, {"512 x 64", 512, 64, 0}, {"256 x 256", 256, 256, 0}, {"256 x 512", 256, 512, 0}, {"128 x 1024", 128, 1024, 0},
      {"148 x 256", 148, 256, 0}, {"296 x 256", 296, 256, 0}, {"192 x 256", 192, 256, 0}, {"64 x 256", 64, 256, 0}, {"128 x 256", 128, 256, 0}};
  const bool quick = getenv("PDL_GAP_QUICK") != nullptr;
  for (auto& c : cfgs) {
    for (int work : {0, 2000}) {
      for (int mode = 0; mode < 3; ++mode) {   // 0 plain, 1 PDL trigger early, 2 PDL trigger late
        for (int graph = 0; graph < 2; ++graph) {
          if (quick && !(mode == 1 && graph == 1)) continue;
          auto issue = [&]() {
            for (int i = 0; i < N; ++i) {
              int grid = c.grid, threads = c.threads; size_t smem = c.smem;
              if (c.grid < 0) { if (i & 1) { grid = 16; threads = 256; smem = 0; } else { grid = 128; threads = 128; smem = 97 * 1024; } }
              launch(st, dim3(grid), threads, smem, mode != 0, buf + (i & 1) * 1024, buf + ((i + 1) & 1) * 1024, mode == 1, work);
            }
          };
          float ms = 0;
          if (graph) {
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal)); issue(); CK(cudaStreamEndCapture(st, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
            CK(cudaEventRecord(a, st)); CK(cudaGraphLaunch(ge, st)); CK(cudaEventRecord(b, st)); CK(cudaEventSynchronize(b));
            CK(cudaEventElapsedTime(&ms, a, b));
            cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
          } else {
            issue(); CK(cudaStreamSynchronize(st));
            CK(cudaEventRecord(a, st)); issue(); CK(cudaEventRecord(b, st)); CK(cudaEventSynchronize(b));
            CK(cudaEventElapsedTime(&ms, a, b));
          }
          printf("%-30s work=%4d %-18s %-6s: %.2f us per kernel\n", c.name, work, mode == 0 ? "stream order" : (mode == 1 ? "PDL trigger early" : "PDL trigger late"),
                 graph ? "graph" : "stream", ms * 1000 / N);
        }
      }
    }
  }
  return 0;
}
