// (d) pair proposal, log-sum-exp normalisation + ESS + multinomial resampling, counter-based uniforms.
//
// propose_pairs : extend_partial_state, vcsmc.py:298-305
// resample      : resample, vcsmc.py:284-285 (tf.random.categorical: running fp64 sum of exp(logit - max),
//                 one uniform per draw, upper_bound(u * total)); logsumexp also feeds compute_log_ZSMC
//                 (vcsmc.py:276).  The CDF is built by ONE CTA in a fixed blocked order, so every GPU that
//                 holds the same K log-weights derives bit-identical ancestors.
#include "launch.h"
#include "smc_device.cuh"

namespace vcsmc {
namespace {

constexpr int kWarpsPerCta = 8;

__global__ void __launch_bounds__(kWarpsPerCta * 32) propose_pairs_kernel(const float* __restrict__ u, int64_t K, int n,
                                                                         int32_t* __restrict__ coal,
                                                                         int32_t* __restrict__ rem) {
  extern __shared__ float su_all[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * kWarpsPerCta + wid;
  if (k >= K) return;
  float* su = su_all + wid * n;
  for (int i = lane; i < n; i += 32) su[i] = u[k * n + i];
  __syncwarp();
  int c0, c1;
  int32_t* rk = rem + k * (int64_t)(n - 2);
  rank_pairs_warp(su, n, lane, c0, c1, [&](int pos, int i) { rk[pos] = i; });
  if (lane == 0) {
    coal[k * 2] = c0;
    coal[k * 2 + 1] = c1;
  }
}

constexpr int kCdfThreads = 1024;

// stats[0] = logsumexp(lw), stats[1] = total = cdf[K-1], stats[2] = ESS, stats[3] = max(lw)
// One CTA, fixed order: warp w owns the contiguous segment [w*seg, (w+1)*seg) and walks it 32 elements at a time
// (coalesced); the running sum inside a segment is a warp shuffle scan plus a carried offset.
__global__ void __launch_bounds__(kCdfThreads) resample_cdf_kernel(const double* __restrict__ lw, int64_t K,
                                                                   double* __restrict__ cdf,
                                                                   double* __restrict__ stats) {
  constexpr int NW = kCdfThreads / 32;
  __shared__ double sm[NW];
  __shared__ double sm2[NW];
  __shared__ double bc[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t seg = ((K + NW - 1) / NW + 31) / 32 * 32;
  const int64_t b = min((int64_t)wid * seg, K), e = min(b + seg, K);

  // max  (4 independent loads in flight per lane: a single CTA is latency-bound, not bandwidth-bound)
  double m = -INFINITY;
  for (int64_t i0 = b + lane; i0 < e; i0 += 128) {
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = (i0 + 32 * q < e) ? lw[i0 + 32 * q] : -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) m = fmax(m, v[q]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) sm[wid] = m;
  __syncthreads();
  if (tid == 0) {
    double t = sm[0];
    for (int i = 1; i < NW; ++i) t = fmax(t, sm[i]);
    bc[0] = t;
  }
  __syncthreads();
  const double M = bc[0];

  // logsumexp (fixed order: lane-strided partial sums, shuffle tree, then warps in order)
  double s = 0.0;
  for (int64_t i0 = b + lane; i0 < e; i0 += 128) {
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = (i0 + 32 * q < e) ? lw[i0 + 32 * q] : -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) s += exp(v[q] - M);   // exp(-inf) = 0 for the padding
  }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) sm[wid] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < NW; ++i) t += sm[i];
    bc[1] = M + log(t);
  }
  __syncthreads();
  const double lse = bc[1];


  // segment totals of w = exp(logit - max) and of w^2
  double run = 0.0, sq = 0.0;
  for (int64_t i0 = b + lane; i0 < e; i0 += 128) {
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = (i0 + 32 * q < e) ? lw[i0 + 32 * q] : -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double w = exp(v[q] - M);
      run += w;
      sq = fma(w, w, sq);
    }
  }
  run = warp_sum(run);
  sq = warp_sum(sq);
  __syncthreads();
  if (lane == 0) {
    sm[wid] = run;
    sm2[wid] = sq;
  }
  __syncthreads();
  if (tid == 0) {
    double t = 0.0, q = 0.0;
    for (int i = 0; i < NW; ++i) {
      const double v = sm[i];
      sm[i] = t;  // exclusive offset of segment i
      t += v;
      q += sm2[i];
    }
    stats[0] = lse;
    stats[1] = t;
    stats[2] = t * t / q;
    stats[3] = M;
  }
  __syncthreads();
  // running sum inside the segment
  double carry = sm[wid];
  for (int64_t i0 = b; i0 < e; i0 += 128) {
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t i = i0 + 32 * q + lane;
      v[q] = i < e ? exp(lw[i] - M) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t i = i0 + 32 * q + lane;
      double w = v[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      if (i < e) cdf[i] = carry + w;
      carry += __shfl_sync(0xffffffffu, w, 31);
    }
  }
}

// ---- the same statistics and CDF with ceil(K / 2048) CTAs (three tiny launches instead of one 80 us single-CTA kernel
// at K = 65,536).  Every reduction runs in a fixed order (cdf_stage_* in smc_device.cuh, shared with the lazy forward's
// event kernel), so every GPU -- and every schedule -- derives bit-identical ancestors.
__global__ void __launch_bounds__(256) cdf_max_kernel(const double* __restrict__ lw, int64_t K, double* __restrict__ pmax) {
  __shared__ double sm[8];
  cdf_stage_max((int)blockIdx.x, lw, K, pmax, sm);
}

__global__ void __launch_bounds__(256) cdf_weights_kernel(const double* __restrict__ lw, int64_t K, int nb,
                                                          const double* __restrict__ pmax, double* __restrict__ w_out,
                                                          double* __restrict__ pw, double* __restrict__ pq) {
  __shared__ double sm[8];
  const double M = block_reduce_array(pmax, nb, true, sm);
  cdf_stage_weights((int)blockIdx.x, lw, K, M, w_out, pw, pq, nullptr, sm);
}

__global__ void __launch_bounds__(256) cdf_scan_kernel(int64_t K, int nb, const double* __restrict__ pmax,
                                                       const double* __restrict__ pw, const double* __restrict__ pq,
                                                       double* __restrict__ cdf, double* __restrict__ stats) {
  __shared__ double sm[8];
  __shared__ double wsum[8];
  const double M = block_reduce_array(pmax, nb, true, sm);
  cdf_stage_scan((int)blockIdx.x, K, nb, M, pw, pq, cdf, stats, sm, wsum);
}

// small K: the same three stages, tile after tile, by ONE CTA in one launch (identical arithmetic)
__global__ void __launch_bounds__(256) cdf_one_cta_kernel(const double* __restrict__ lw, int64_t K, int nb, double* __restrict__ scratch,
                                                          double* __restrict__ cdf, double* __restrict__ stats) {
  __shared__ double sm[8];
  __shared__ double wsum[8];
  double *pmax = scratch, *pw = scratch + 2 * nb, *pq = scratch + 3 * nb;
  for (int vb = 0; vb < nb; ++vb) cdf_stage_max(vb, lw, K, pmax, sm);
  __syncthreads();
  const double M = block_reduce_array(pmax, nb, true, sm);
  for (int vb = 0; vb < nb; ++vb) cdf_stage_weights(vb, lw, K, M, cdf, pw, pq, nullptr, sm);
  __syncthreads();
  for (int vb = 0; vb < nb; ++vb) cdf_stage_scan(vb, K, nb, M, pw, pq, cdf, stats, sm, wsum);
}

__global__ void resample_search_kernel(const double* __restrict__ cdf, const double* __restrict__ stats,
                                       const double* __restrict__ u, int64_t K, int32_t* __restrict__ idx) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= K) return;
  idx[j] = upper_bound_cdf(cdf, K, u[j] * cdf[K - 1]);
}

__global__ void philox_step_kernel(uint64_t seed, const uint64_t* __restrict__ seed_dev, int r, int64_t k0, int64_t K, int n,
                                   float* __restrict__ u_pair, double* __restrict__ u_bl, double* __restrict__ u_br,
                                   double* __restrict__ u_res, double* __restrict__ u_cat) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  if (seed_dev) seed = *seed_dev;  // device-resident seed: a captured launch sequence can be replayed with a new seed
  const uint64_t k = (uint64_t)(k0 + i);  // LOGICAL particle index: identical streams for any GPU count
  const uint32_t s0 = (uint32_t)seed, s1 = (uint32_t)(seed >> 32);
  uint32_t c[4] = {(uint32_t)k, (uint32_t)(k >> 32) | ((uint32_t)r << 8), 0u, 0u};
  philox4x32_10(c, s0, s1);
  const double tiny = 2.2250738585072014e-308;
  if (u_bl) u_bl[i] = fmax(u64_to_unit_f64(c[0], c[1]), tiny);   // tfp Exponential: U in [tiny, 1)
  if (u_br) u_br[i] = fmax(u64_to_unit_f64(c[2], c[3]), tiny);
  if (u_res || u_cat) {
    uint32_t d[4] = {(uint32_t)k, (uint32_t)(k >> 32) | ((uint32_t)r << 8), 1u, 0u};
    philox4x32_10(d, s0, s1);
    if (u_res) u_res[i] = u64_to_unit_f64(d[0], d[1]);
    if (u_cat) u_cat[i] = u64_to_unit_f64(d[2], d[3]);
  }
  if (u_pair) {
    for (int j = 0; j < n; j += 4) {
      uint32_t p[4] = {(uint32_t)k, (uint32_t)(k >> 32) | ((uint32_t)r << 8), 2u, (uint32_t)(j >> 2)};
      philox4x32_10(p, s0, s1);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (j + q < n) {
          // 16 random bits + the 8-bit position: values inside one particle's row are pairwise DISTINCT.  Exact
          // float32 ties make tf.nn.top_k keep one subtree twice and drop another (vcsmc.py:304-305, SURVEY Q-ties);
          // at K = 65,536 that happens a few times per sweep and the truncated forest then wins every resampling.
          // The proposal kernels keep the reference's tie semantics; this generator just never produces a tie.
          const uint32_t bits = ((p[q] >> 8) & 0xFFFF00u) | (uint32_t)((j + q) & 0xFF);
          u_pair[i * n + j + q] = (float)bits * 5.9604644775390625e-8f;
        }
    }
  }
}

}  // namespace

int launch_propose_pairs(const float* u, int64_t K, int n, int32_t* coal, int32_t* rem, cudaStream_t st) {
  if (K <= 0) return VCSMC_OK;
  if (n < 2 || n > kMaxRoots) { set_error("propose_pairs: n=%d out of range [2,%d]", n, kMaxRoots); return VCSMC_ERR_ARG; }
  const size_t smem = (size_t)kWarpsPerCta * n * sizeof(float);
  propose_pairs_kernel<<<(unsigned)((K + kWarpsPerCta - 1) / kWarpsPerCta), kWarpsPerCta * 32, smem, st>>>(u, K, n, coal, rem);
  VCSMC_LAUNCH_CHECK("propose_pairs_kernel");
  return VCSMC_OK;
}

int64_t resample_scratch_doubles(int64_t K) { return 4 * ((K + kCdfTile - 1) / kCdfTile) + 4; }

int launch_resample_cdf(const double* lw, int64_t K, double* cdf, double* stats, double* scratch, cudaStream_t st) {
  if (K <= 0) return VCSMC_OK;
  if (!scratch) {  // no scratch for the tile partials: the single-CTA routine (its own fixed order)
    resample_cdf_kernel<<<1, kCdfThreads, 0, st>>>(lw, K, cdf, stats);
    VCSMC_LAUNCH_CHECK("resample_cdf_kernel");
    return VCSMC_OK;
  }
  const int nb = (int)((K + kCdfTile - 1) / kCdfTile);
  double *pmax = scratch, *pw = scratch + 2 * nb, *pq = scratch + 3 * nb;
  if (nb <= 2) {   // one launch; same arithmetic as the four-launch path and as the lazy forward's event kernel
    cdf_one_cta_kernel<<<1, 256, 0, st>>>(lw, K, nb, scratch, cdf, stats);
    VCSMC_LAUNCH_CHECK("cdf_one_cta_kernel");
    return VCSMC_OK;
  }
  cdf_max_kernel<<<nb, 256, 0, st>>>(lw, K, pmax);
  VCSMC_LAUNCH_CHECK("cdf_max_kernel");
  cdf_weights_kernel<<<nb, 256, 0, st>>>(lw, K, nb, pmax, cdf, pw, pq);
  VCSMC_LAUNCH_CHECK("cdf_weights_kernel");
  cdf_scan_kernel<<<nb, 256, 0, st>>>(K, nb, pmax, pw, pq, cdf, stats);
  VCSMC_LAUNCH_CHECK("cdf_scan_kernel");
  return VCSMC_OK;
}

int launch_resample_search(const double* cdf, const double* stats, const double* u, int64_t K, int32_t* idx,
                           cudaStream_t st) {
  if (K <= 0) return VCSMC_OK;
  resample_search_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(cdf, stats, u, K, idx);
  VCSMC_LAUNCH_CHECK("resample_search_kernel");
  return VCSMC_OK;
}

int launch_philox_step(uint64_t seed, int r, int64_t k0, int64_t K, int n, float* u_pair, double* u_bl, double* u_br,
                       double* u_res, double* u_cat, cudaStream_t st, const uint64_t* seed_dev) {
  if (K <= 0) return VCSMC_OK;
  philox_step_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(seed, seed_dev, r, k0, K, n, u_pair, u_bl, u_br, u_res, u_cat);
  VCSMC_LAUNCH_CHECK("philox_step_kernel");
  return VCSMC_OK;
}

}  // namespace vcsmc
