// Device pieces shared by the standalone (d) kernels (smc.cu) and the fused per-step kernels (sweep.cu).
#pragma once
#include "common.cuh"

namespace vcsmc {

constexpr int kMaxRoots = 256;  // forest positions are stored as uint8

// Pair proposal of vcsmc.py:298-305 for ONE particle by ONE warp.
// z = -log(-log u) is increasing in u, so ranking u reproduces tf.nn.top_k(z) / top_k(-z) without a float32
// libm dependence; ties go to the lower index in BOTH rankings, exactly like tf.nn.top_k (including the
// reference's duplicate-on-tie quirk).  su: this warp's shared row holding u[0..n).
//   rd(i) = #{j : u_j > u_i or (u_j == u_i and j < i)}   (descending rank)  -> coal = rd 0, 1
//   ra(i) = #{j : u_j < u_i or (u_j == u_i and j < i)}   (ascending rank)   -> rem[ra] = i for ra < n-2
template <typename F>
__device__ __forceinline__ void rank_pairs_warp(const float* su, int n, int lane, int& c0, int& c1, F&& emit_rem) {
  int my0 = -1, my1 = -1;
  for (int i = lane; i < n; i += 32) {
    const float ui = su[i];
    int rd = 0, ra = 0;
    for (int j = 0; j < n; ++j) {
      const float uj = su[j];
      const bool tie_lo = (uj == ui) && (j < i);
      rd += (uj > ui) || tie_lo;
      ra += (uj < ui) || tie_lo;
    }
    if (rd == 0) my0 = i;
    if (rd == 1) my1 = i;
    if (ra < n - 2) emit_rem(ra, i);
  }
  // exactly one lane holds each of rank 0 / rank 1
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    my0 = max(my0, __shfl_xor_sync(0xffffffffu, my0, o));
    my1 = max(my1, __shfl_xor_sync(0xffffffffu, my1, o));
  }
  c0 = my0;
  c1 = my1;
}

// first i in [0,K) with cdf[i] > t (std::upper_bound), clamped to K-1
__device__ __forceinline__ int upper_bound_cdf(const double* __restrict__ cdf, int64_t K, double t) {
  int64_t lo = 0, hi = K;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (cdf[mid] > t) hi = mid;
    else lo = mid + 1;
  }
  return (int)(lo < K ? lo : K - 1);
}

}  // namespace vcsmc
