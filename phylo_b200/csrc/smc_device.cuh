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


// ---------------------------------------------------------------------------------------------
// log-sum-exp, ESS and the categorical CDF of K log-weights in tiles of 2048 elements (resample, vcsmc.py:284-285; also
// compute_log_ZSMC's reduce_logsumexp, vcsmc.py:276).  Three stages, each needing the previous one's partials of ALL tiles;
// `vb` is the tile a CTA of exactly 256 threads works on.  Every reduction has a fixed order that depends on K only, so
// whoever runs the stages (four launches, one CTA, the lazy forward's cooperative event kernel; 1 or 8 GPUs) gets the
// same bits, hence the same ancestors.
//   stats[0] = logsumexp(lw), stats[1] = total = cdf[K-1], stats[2] = ESS, stats[3] = max(lw)
// ---------------------------------------------------------------------------------------------
constexpr int kCdfTile = 2048;  // elements per tile: 256 threads x 8

__device__ __forceinline__ double block_reduce_array(const double* p, int n, bool is_max, double* sm) {
  // every thread returns the same value: strided partials in thread order, then the block tree (fixed order)
  double v = is_max ? -INFINITY : 0.0;
  for (int i = threadIdx.x; i < n; i += 256) v = is_max ? fmax(v, p[i]) : v + p[i];
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = sm[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) t = is_max ? fmax(t, sm[i]) : t + sm[i];
  __syncthreads();
  return t;
}

__device__ __forceinline__ void cdf_stage_max(int vb, const double* lw, int64_t K, double* pmax, double* sm) {
  const int64_t b = (int64_t)vb * kCdfTile + threadIdx.x;
  double m = -INFINITY;
#pragma unroll
  for (int q = 0; q < 8; ++q) m = fmax(m, b + q * 256 < K ? lw[b + q * 256] : -INFINITY);
  m = warp_max(m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = sm[0];
    for (int i = 1; i < 8; ++i) t = fmax(t, sm[i]);
    pmax[vb] = t;
  }
}

// Stage 2: unnormalised weights w = exp(lw - M), M = the maximum of all K log-weights (exact in any order: the tile
// maxima of stage 1 reduced, or an atomic max), and their tile sums.  tf.random.categorical draws from exp(logit - max
// logit) with logit = lw - logsumexp(lw) (vcsmc.py:284-285), i.e. from the same numbers up to the rounding of two
// subtractions instead of one; logsumexp = M + log(sum w) needs no pass of its own.
// live (optional): 1 where the weight is not zero in double precision -- the particles that can be drawn by the next
// resampling or carry a gradient
__device__ __forceinline__ void cdf_stage_weights(int vb, const double* lw, int64_t K, double M, double* w_out, double* pw,
                                                  double* pq, int32_t* live, double* sm) {
  const int64_t b = (int64_t)vb * kCdfTile + (int64_t)threadIdx.x * 8;   // 8 CONSECUTIVE elements per thread
  double s = 0.0, q2 = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (b + q < K) {
      const double w = exp(lw[b + q] - M);
      w_out[b + q] = w;
      if (live) live[b + q] = w != 0.0;
      s += w;
      q2 = fma(w, w, q2);
    }
  }
  const double ts = block_sum<256>(s, sm);
  const double tq = block_sum<256>(q2, sm);
  if (threadIdx.x == 0) {
    pw[vb] = ts;
    pq[vb] = tq;
  }
}

__device__ __forceinline__ void cdf_stage_scan(int vb, int64_t K, int nb, double M, const double* pw, const double* pq,
                                               double* cdf, double* stats, double* sm, double* wsum) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const double offset = block_reduce_array(pw, vb, false, sm);   // sum of the tiles before this one
  const int64_t b = (int64_t)vb * kCdfTile + (int64_t)tid * 8;
  double w[8], run = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    w[q] = b + q < K ? cdf[b + q] : 0.0;
    run += w[q];
    w[q] = run;  // inclusive inside the thread
  }
  double incl = run;  // inclusive scan of the thread totals inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  double wbase = 0.0;
  for (int i = 0; i < wid; ++i) wbase += wsum[i];
  const double base = offset + wbase + (incl - run);
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (b + q < K) cdf[b + q] = base + w[q];
  if (vb == 0) {
    __syncthreads();
    const double t = block_reduce_array(pw, nb, false, sm);
    const double q = block_reduce_array(pq, nb, false, sm);
    if (tid == 0) {
      stats[0] = M + log(t);
      stats[1] = t;
      stats[2] = t * t / q;
      stats[3] = M;
    }
  }
}

}  // namespace vcsmc
