// (c) merge kernel and (e) its reverse-pruning adjoint.
//
// Forward replaces broadcast_conditional_likelihood_K (vcsmc.py:180-188) fused with the NEW node's term of
// compute_forest_posterior (vcsmc.py:238-242).  Backward is the per-site part of what TF autodiff does to
// those ops (vcsmc.py:488-491).  HBM-bound 4x4 contraction on CUDA cores: one 256-bit access per site
// vector, per-particle P matrices in registers, fixed-order CTA reduction of the site log-likelihoods.
#include "common.cuh"
#include "launch.h"

namespace vcsmc {

namespace {

// P for one child.  General: 16 entries.  JC: P = o*1 1^T + (d-o) I, so lp_j = o*sum(L) + (d-o) L_j.
template <bool JC>
struct Trans;
template <>
struct Trans<false> {
  double p[16];
  __device__ __forceinline__ void load(const double* __restrict__ P) {
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = __ldg(P + i);
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
  __device__ __forceinline__ void load(const double* __restrict__ P) {
    const double d = __ldg(P);
    o = __ldg(P + 1);
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

struct FwdArgs {
  const uint8_t* codes;
  int64_t codes_stride;
  double* pool;
  int64_t slot_sites;
  const int32_t* lsrc;
  const int32_t* rsrc;
  const int32_t* dst;
  const double* P;
  const double* pi;
  int n_sites;
  int tiles;
  int skip_unstored;  // re-forward of the chunked backward: nodes nobody consumes are not materialised
  double* ell_part;
};

__device__ __forceinline__ d4 load_child(const uint8_t* __restrict__ codes_row, const double* __restrict__ node, int s) {
  return codes_row ? leaf_site(__ldg(codes_row + s)) : ld_site(node + (int64_t)s * 4);
}

template <bool JC>
__global__ void __launch_bounds__(kTileThreads) merge_fwd_kernel(const FwdArgs a) {
  __shared__ double red[kTileThreads / 32];
  const int64_t w = blockIdx.x;
  const int64_t k = w / a.tiles;
  const int tile = (int)(w - k * a.tiles);
  const int ls = a.lsrc[k], rs = a.rsrc[k], ds = a.dst ? a.dst[k] : (int)k;
  if (a.skip_unstored && ds < 0) return;

  Trans<JC> Pl, Pr;
  Pl.load(a.P + k * 32);
  Pr.load(a.P + k * 32 + 16);
  double pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = __ldg(a.pi + j);
  const uint8_t* lcodes = ls < 0 ? a.codes + (int64_t)(-ls - 1) * a.codes_stride : nullptr;
  const uint8_t* rcodes = rs < 0 ? a.codes + (int64_t)(-rs - 1) * a.codes_stride : nullptr;
  const double* lnode = ls < 0 ? nullptr : a.pool + (int64_t)ls * a.slot_sites * 4;
  const double* rnode = rs < 0 ? nullptr : a.pool + (int64_t)rs * a.slot_sites * 4;
  double* out = ds < 0 ? nullptr : a.pool + (int64_t)ds * a.slot_sites * 4;

  const int s0 = tile * kTileSites + threadIdx.x;
  d4 Ll[kSitesPerThread], Lr[kSitesPerThread];
#pragma unroll
  for (int it = 0; it < kSitesPerThread; ++it) {
    const int s = s0 + it * kTileThreads;
    if (s < a.n_sites) {
      Ll[it] = load_child(lcodes, lnode, s);
      Lr[it] = load_child(rcodes, rnode, s);
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int it = 0; it < kSitesPerThread; ++it) {
    const int s = s0 + it * kTileThreads;
    if (s < a.n_sites) {
      const d4 lp = Pl.apply(Ll[it]), rp = Pr.apply(Lr[it]);
      d4 nw;
      double x = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        nw.v[j] = lp.v[j] * rp.v[j];
        x = fma(pi[j], nw.v[j], x);
      }
      if (out) st_site(out + (int64_t)s * 4, nw);
      acc += log(x);
    }
  }
  const double t = block_sum<kTileThreads>(acc, red);
  if (threadIdx.x == 0) a.ell_part[k * a.tiles + tile] = t;
}

__global__ void ell_reduce_kernel(const double* __restrict__ part, int tiles, int64_t K, double* __restrict__ ell) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double s = 0.0;
  for (int t = 0; t < tiles; ++t) s += part[k * tiles + t];
  ell[k] = s;
}

struct BwdArgs {
  const uint8_t* codes;
  int64_t codes_stride;
  const double* pool;
  double* gpool;
  int64_t slot_sites;
  const int32_t* lsrc;
  const int32_t* rsrc;
  const int32_t* gsrc;
  const double* P;
  const double* pi;
  const double* coef;
  int n_sites;
  int tiles;
  double* dP;        // [K][32]
  double* dpi_each;  // [K][4] or null
  int skip_zero;
};

constexpr int kAccGeneral = 36;  // dP_l[16] dP_r[16] dpi[4]
constexpr int kAccJC = 8;        // dPl_diag dPl_off dPr_diag dPr_off dpi[4]

template <bool JC>
__global__ void __launch_bounds__(kTileThreads) merge_bwd_kernel(const BwdArgs a) {
  constexpr int NACC = JC ? kAccJC : kAccGeneral;
  __shared__ double red[kTileThreads / 32][NACC];
  const int64_t w = blockIdx.x;
  const int64_t k = w / a.tiles;
  const int tile = (int)(w - k * a.tiles);
  const double c = a.coef[k];
  const int gs = a.gsrc ? a.gsrc[k] : -1;
  if (a.skip_zero && c == 0.0 && gs < 0) return;  // exact zero adjoint: nothing to propagate
  const int ls = a.lsrc[k], rs = a.rsrc[k];

  Trans<JC> Pl, Pr;
  Pl.load(a.P + k * 32);
  Pr.load(a.P + k * 32 + 16);
  double pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = __ldg(a.pi + j);
  const uint8_t* lcodes = ls < 0 ? a.codes + (int64_t)(-ls - 1) * a.codes_stride : nullptr;
  const uint8_t* rcodes = rs < 0 ? a.codes + (int64_t)(-rs - 1) * a.codes_stride : nullptr;
  const double* lnode = ls < 0 ? nullptr : a.pool + (int64_t)ls * a.slot_sites * 4;
  const double* rnode = rs < 0 ? nullptr : a.pool + (int64_t)rs * a.slot_sites * 4;
  double* lg = ls < 0 ? nullptr : a.gpool + (int64_t)ls * a.slot_sites * 4;
  double* rg = rs < 0 ? nullptr : a.gpool + (int64_t)rs * a.slot_sites * 4;
  const double* gnew = gs < 0 ? nullptr : a.gpool + (int64_t)gs * a.slot_sites * 4;

  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

  const int s0 = tile * kTileSites + threadIdx.x;
#pragma unroll 2
  for (int it = 0; it < kSitesPerThread; ++it) {
    const int s = s0 + it * kTileThreads;
    if (s >= a.n_sites) break;
    const d4 Ll = load_child(lcodes, lnode, s), Lr = load_child(rcodes, rnode, s);
    d4 g;
    if (gnew) {
      g = ld_site(gnew + (int64_t)s * 4);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) g.v[j] = 0.0;
    }
    const d4 lp = Pl.apply(Ll), rp = Pr.apply(Lr);
    double nw[4], x = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      nw[j] = lp.v[j] * rp.v[j];
      x = fma(pi[j], nw[j], x);
    }
    const double inv = c / x;
    d4 gl, gr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double gj = fma(inv, pi[j], g.v[j]);
      acc[NACC - 4 + j] = fma(inv, nw[j], acc[NACC - 4 + j]);  // d ell / d pi_j = new_j / x
      gl.v[j] = gj * rp.v[j];
      gr.v[j] = gj * lp.v[j];
    }
    if (lg) {
      const d4 t = Pl.apply_t(gl);
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(lg + (int64_t)s * 4 + i, t.v[i]);
    }
    if (rg) {
      const d4 t = Pr.apply_t(gr);
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(rg + (int64_t)s * 4 + i, t.v[i]);
    }
    if (JC) {
      // dP enters only through (sum_i dP_ii, sum_{i!=j} dP_ij): dP_ij = L_i gl_j
      double dl = 0.0, dr = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        dl = fma(Ll.v[i], gl.v[i], dl);
        dr = fma(Lr.v[i], gr.v[i], dr);
      }
      const double sl = ((Ll.v[0] + Ll.v[1]) + (Ll.v[2] + Ll.v[3])) * ((gl.v[0] + gl.v[1]) + (gl.v[2] + gl.v[3]));
      const double sr = ((Lr.v[0] + Lr.v[1]) + (Lr.v[2] + Lr.v[3])) * ((gr.v[0] + gr.v[1]) + (gr.v[2] + gr.v[3]));
      acc[0] += dl;
      acc[1] += sl - dl;
      acc[2] += dr;
      acc[3] += sr - dr;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i * 4 + j] = fma(Ll.v[i], gl.v[j], acc[i * 4 + j]);
          acc[16 + i * 4 + j] = fma(Lr.v[i], gr.v[j], acc[16 + i * 4 + j]);
        }
    }
  }
  // CTA reduction of the accumulators, then one atomic per value (tiles of one particle race only here)
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    const double t = warp_sum(acc[i]);
    if (lane == 0) red[wid][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < kTileThreads / 32; ++q) t += red[q][threadIdx.x];
    const int i = threadIdx.x;
    if (JC) {
      if (i < 4) {
        // JC layout inside dP[k][32]: [0]=sum diag(dP_l), [1]=sum offdiag(dP_l), [16],[17] same for right
        atomicAdd(a.dP + k * 32 + (i >> 1) * 16 + (i & 1), t);
      } else if (a.dpi_each) {
        atomicAdd(a.dpi_each + k * 4 + (i - 4), t);
      }
    } else {
      if (i < 32) {
        atomicAdd(a.dP + k * 32 + i, t);
      } else if (a.dpi_each) {
        atomicAdd(a.dpi_each + k * 4 + (i - 32), t);
      }
    }
  }
}

}  // namespace

int merge_tiles(int n_sites) { return (n_sites + kTileSites - 1) / kTileSites; }

int launch_merge_fwd(const uint8_t* codes, int64_t codes_stride, double* pool, int64_t slot_sites, const int32_t* lsrc,
                     const int32_t* rsrc, const int32_t* dst, const double* P, const double* pi, int64_t K,
                     int n_sites, int jc, int skip_unstored, double* ell_part, cudaStream_t st) {
  if (K <= 0 || n_sites <= 0) return VCSMC_OK;
  FwdArgs a;
  a.codes = codes; a.codes_stride = codes_stride; a.pool = pool; a.slot_sites = slot_sites;
  a.lsrc = lsrc; a.rsrc = rsrc; a.dst = dst; a.P = P;
  a.pi = pi;
  a.n_sites = n_sites; a.tiles = merge_tiles(n_sites); a.skip_unstored = skip_unstored; a.ell_part = ell_part;
  const int64_t grid = K * a.tiles;
  if (grid > 2147483647LL) { set_error("merge_fwd: grid too large"); return VCSMC_ERR_ARG; }
  if (jc) merge_fwd_kernel<true><<<(unsigned)grid, kTileThreads, 0, st>>>(a);
  else merge_fwd_kernel<false><<<(unsigned)grid, kTileThreads, 0, st>>>(a);
  VCSMC_LAUNCH_CHECK("merge_fwd_kernel");
  return VCSMC_OK;
}

int launch_ell_reduce(const double* ell_part, int tiles, int64_t K, double* ell, cudaStream_t st) {
  ell_reduce_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(ell_part, tiles, K, ell);
  VCSMC_LAUNCH_CHECK("ell_reduce_kernel");
  return VCSMC_OK;
}

int launch_merge_bwd(const uint8_t* codes, int64_t codes_stride, const double* pool, double* gpool, int64_t slot_sites,
                     const int32_t* lsrc, const int32_t* rsrc, const int32_t* gsrc, const double* P,
                     const double* pi, const double* coef, int64_t K, int n_sites, int jc, int skip_zero,
                     double* dP, double* dpi_each, cudaStream_t st) {
  if (K <= 0 || n_sites <= 0) return VCSMC_OK;
  BwdArgs a;
  a.codes = codes; a.codes_stride = codes_stride; a.pool = pool; a.gpool = gpool; a.slot_sites = slot_sites;
  a.lsrc = lsrc; a.rsrc = rsrc; a.gsrc = gsrc; a.P = P;
  a.pi = pi;
  a.coef = coef; a.n_sites = n_sites; a.tiles = merge_tiles(n_sites); a.dP = dP; a.dpi_each = dpi_each;
  a.skip_zero = skip_zero;
  const int64_t grid = K * a.tiles;
  if (grid > 2147483647LL) { set_error("merge_bwd: grid too large"); return VCSMC_ERR_ARG; }
  if (jc) merge_bwd_kernel<true><<<(unsigned)grid, kTileThreads, 0, st>>>(a);
  else merge_bwd_kernel<false><<<(unsigned)grid, kTileThreads, 0, st>>>(a);
  VCSMC_LAUNCH_CHECK("merge_bwd_kernel");
  return VCSMC_OK;
}

}  // namespace vcsmc
