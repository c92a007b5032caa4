// Device helpers shared by the merge kernels (merge.cu: eager forward + reverse pruning; score.cu: likelihood-only
// scoring, survivor materialisation, peer pulls).
#pragma once
#include <limits.h>

#include "common.cuh"

namespace vcsmc {

constexpr int kRMax = 16;       // particles per group (shared-memory staging)
constexpr int kWarps = kTileThreads / 32;
constexpr int kNone = INT_MIN;  // "no child loaded"

// P for one child.  General: 16 entries.  JC: P = o*1 1^T + (d-o) I, so lp_j = o*sum(L) + (d-o) L_j.
template <bool JC>
struct Trans;
template <>
struct Trans<false> {
  double p[16];
  __device__ __forceinline__ void load(const double* P) {
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = P[i];
  }
  // row-vector convention (quirk Q5): out_j = sum_i L_i P[i][j]
  __device__ __forceinline__ d4 apply(const d4& L) const {
    d4 o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = L.v[0] * p[j];
#pragma unroll
      for (int i = 1; i < 4; ++i) s = fma(L.v[i], p[i * 4 + j], s);
      o.v[j] = s;
    }
    return o;
  }
  // out_i = sum_j P[i][j] g_j
  __device__ __forceinline__ d4 apply_t(const d4& g) const {
    d4 o;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = p[i * 4] * g.v[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) s = fma(p[i * 4 + j], g.v[j], s);
      o.v[i] = s;
    }
    return o;
  }
};
template <>
struct Trans<true> {
  double dmo, o;  // diag - off, off
  __device__ __forceinline__ void load(const double* P) {
    const double d = P[0];
    o = P[1];
    dmo = d - o;
  }
  __device__ __forceinline__ d4 apply(const d4& L) const {
    const double so = o * ((L.v[0] + L.v[1]) + (L.v[2] + L.v[3]));
    d4 r;
#pragma unroll
    for (int j = 0; j < 4; ++j) r.v[j] = fma(dmo, L.v[j], so);
    return r;
  }
  __device__ __forceinline__ d4 apply_t(const d4& g) const { return apply(g); }  // symmetric
};

__device__ __forceinline__ d4 zero4() {
  d4 z;
#pragma unroll
  for (int j = 0; j < 4; ++j) z.v[j] = 0.0;
  return z;
}

struct ChildRef {
  const uint8_t* codes_row;
  const double* node;
};
__device__ __forceinline__ ChildRef child_ref(int src, const uint8_t* codes, int64_t codes_stride, const double* pool,
                                              int64_t slot_sites) {
  ChildRef c;
  c.codes_row = src < 0 ? codes + (int64_t)(-src - 1) * codes_stride : nullptr;
  c.node = src < 0 ? nullptr : pool + (int64_t)src * slot_sites * 4;
  return c;
}
__device__ __forceinline__ d4 load_child(const ChildRef& c, int s) {
  return c.codes_row ? leaf_site(__ldg(c.codes_row + s)) : ld_site(c.node + (int64_t)s * 4);
}

// log(x) of a positive double split as (mantissa in [1,2), unbiased exponent): sum_s log x_s is then the log of a
// running mantissa product plus ln2 times an integer sum -- ONE log per (thread, particle) instead of one per site.
// Zero, subnormal, inf and NaN are passed through unsplit so that log() of the product still yields what
// the reference's log(0) = -inf / NaN would.
__device__ __forceinline__ void split_positive(double x, double& mant, int& ex) {
  const int hi = __double2hiint(x);
  const int e = (hi >> 20) & 0x7ff;
  const bool normal = (e != 0) && (e != 0x7ff) && (hi >= 0);
  mant = normal ? __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x)) : x;
  ex = normal ? e - 1023 : 0;
}

}  // namespace vcsmc
