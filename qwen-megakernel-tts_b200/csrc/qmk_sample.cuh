// qmk_sample.cuh — device-side token selection shared by the B = 1 group kernel and the batched code-predictor step:
// temperature / top-k (ties kept) / softmax / inverse-CDF draw from a counter-based generator (upstream model_tts.py:756-762
// evaluated on the device).  256 threads, barrier id 1 (`bar.sync 1, 256`).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace qmk2 {

constexpr int SMP_NCT = 256, SMP_NCW = 8;
__device__ __forceinline__ void smp_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SMP_NCT) : "memory"); }

// ---- token sampling (temperature / top-k with ties / multinomial), group-kernel version -------------------------
// Same selection as qmk::sample_token (two-pass radix select of the k-th largest 16-bit key, softmax over the kept
// logits, inverse-CDF draw in index order), restructured for latency: keys live in registers, both histograms are cleared
// up front, and EVERY warp scans a histogram itself instead of waiting for warp 0 to publish the result -- 5 CTA barriers
// instead of 10.  hist: unsigned[2][256].
__device__ __forceinline__ void scan256_from_top(const unsigned* h, int lane, unsigned k, unsigned& bin, unsigned& rank_in_bin) {
  // lane l owns bins 255 - 8 l .. 248 - 8 l
  const uint4 a = *reinterpret_cast<const uint4*>(h + 248 - 8 * lane), b = *reinterpret_cast<const uint4*>(h + 252 - 8 * lane);
  const unsigned cnt[8] = {b.w, b.z, b.y, b.x, a.w, a.z, a.y, a.x};   // descending bin order
  unsigned tot = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) tot += cnt[e];
  unsigned incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned before = incl - tot;
  const bool mine = before < k && incl >= k;
  unsigned my_bin = 0, my_rank = 0, run = before;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (run < k && run + cnt[e] >= k) { my_bin = (unsigned)(255 - 8 * lane - e); my_rank = k - run; }
    run += cnt[e];
  }
  const unsigned who = __ballot_sync(0xffffffffu, mine);
  const int src = who ? __ffs(who) - 1 : 0;
  bin = __shfl_sync(0xffffffffu, my_bin, src);
  rank_in_bin = __shfl_sync(0xffffffffu, my_rank, src);
}
static __device__ __noinline__ int sample_token2(const float* s_log, unsigned* hist, float* s_red, int tid, int warp, int lane, int hrows,
                                          int top_k, float temperature, unsigned long long seed, unsigned long long counter,
                                          int group, float best, int best_i) {
  unsigned* s_sel = reinterpret_cast<unsigned*>(s_red) + 40;
  const int per = hrows / SMP_NCT;   // <= 8 contiguous elements per thread (CDF in index order)
  const int i0 = tid * per;
  float v[8];
  unsigned key[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    v[e] = e < per ? s_log[i0 + e] : 0.f;
    const unsigned bits = __float_as_uint(v[e]) >> 16;
    key[e] = (bits & 0x8000u) ? (~bits & 0xffffu) : (bits | 0x8000u);   // order-preserving 16-bit key of the bf16 logit
  }
  unsigned thr_key = 0;
  if (top_k > 0 && top_k < hrows) {
    hist[tid] = 0;
    hist[256 + tid] = 0;
    smp_bar();
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (e < per) atomicAdd(&hist[key[e] >> 8], 1u);
    smp_bar();
    unsigned hi_bin, rank2, lo_bin, unused;
    scan256_from_top(hist, lane, (unsigned)top_k, hi_bin, rank2);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (e < per && (key[e] >> 8) == hi_bin) atomicAdd(&hist[256 + (key[e] & 0xffu)], 1u);
    smp_bar();
    scan256_from_top(hist + 256, lane, rank2, lo_bin, unused);
    thr_key = (hi_bin << 8) | lo_bin;
  }
  const float inv_t = 1.0f / temperature;
  const float zmax = best * inv_t;
  float pe[8];
  float local = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    pe[e] = (e < per && key[e] >= thr_key) ? __expf(v[e] * inv_t - zmax) : 0.f;
    local += pe[e];
  }
  float incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_red[16 + warp] = incl;
  if (tid == 0) s_sel[2] = (unsigned)best_i;   // fallback if rounding leaves the target beyond the last element
  smp_bar();
  float wbase = 0.f, total = 0.f;
#pragma unroll
  for (int w = 0; w < SMP_NCW; ++w) {
    const float t = s_red[16 + w];
    if (w < warp) wbase += t;
    total += t;
  }
  unsigned long long zr = seed + 0x9E3779B97F4A7C15ull * (counter * 16ull + (unsigned long long)(group + 1));
  zr = (zr ^ (zr >> 30)) * 0xBF58476D1CE4E5B9ull;
  zr = (zr ^ (zr >> 27)) * 0x94D049BB133111EBull;
  zr ^= zr >> 31;
  const float target = (float)(zr >> 40) * (1.0f / 16777216.0f) * total;
  const float lo = wbase + incl - local;
  if (target >= lo && target < lo + local) {
    float run = lo;
    int pick = -1;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (e < per && pe[e] > 0.f) {
        if (pick < 0 && target < run + pe[e]) pick = i0 + e;
        run += pe[e];
      }
    }
    if (pick >= 0) s_sel[2] = (unsigned)pick;
  }
  smp_bar();
  return (int)s_sel[2];
}

}  // namespace qmk2
