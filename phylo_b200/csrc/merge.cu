// (c) merge kernel and (e) its reverse-pruning adjoint.
//
// Forward replaces broadcast_conditional_likelihood_K (vcsmc.py:180-188) fused with the NEW node's term of
// compute_forest_posterior (vcsmc.py:238-242).  Backward is the per-site part of what TF autodiff does to
// those ops (vcsmc.py:488-491).  A 4x4 contraction per site on CUDA cores (no tensor cores: 2 flop/B in fp64).
//
// Work decomposition (both directions): particles are visited in an order sorted by their (unordered) pair of
// child nodes; one work item = a GROUP of up to R consecutive sorted particles x one TILE of 256*SPT sites.
// A thread keeps its sites of the two children in registers (one 256-bit access per site vector) and only
// reloads a child when the sorted order moves on to a different node.  After resampling most particles descend
// from few ancestors and share children, so a child is read once per group instead of once per particle; with
// all-distinct children the kernel degenerates to a plain streaming merge.  The per-particle P matrices of a
// group are staged in shared memory with one cooperative load.  Backward additionally accumulates the adjoint
// of a shared child in registers across the run and issues ONE set of fp64 atomics per run instead of one per
// particle, and reduces the per-particle 4x4 adjoints with a halving-butterfly warp transpose.
#include <limits.h>
#include <stdlib.h>

#include "launch.h"
#include "merge_device.cuh"

namespace vcsmc {

namespace {

// Sum of v[i] over the 32 lanes of a warp for 32 values at once: afterwards lane L holds the total of v[L]
// in v[0].  Halving butterfly: 31 exchanges instead of 32 x 5.
__device__ __forceinline__ void warp_transpose_sum32(double (&v)[32], int lane) {
#pragma unroll
  for (int s = 16, cnt = 16; s >= 1; s >>= 1, cnt >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < cnt; ++i) {
      const double send = up ? v[i] : v[i + cnt];
      const double keep = up ? v[i + cnt] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

struct FwdArgs {
  const uint8_t* codes;
  int64_t codes_stride;
  double* pool;
  int64_t slot_sites;
  const int32_t* lsrc;
  const int32_t* rsrc;
  const int32_t* dst;
  const int32_t* order;  // sorted visiting order (null: identity)
  const int32_t* count;  // number of leading entries of `order` to process (null: K)
  const double* P;
  const double* pi;
  int64_t K;
  int n_sites;
  int tiles;
  int tiles_per_item;  // a work item = one group x tiles_per_item consecutive tiles (staging amortised)
  int n_chunks;
  int R;
  int skip_unstored;  // re-forward of the chunked backward: nodes nobody consumes are not materialised
  double* ell_part;   // [K][n_chunks][kWarps]
};

// stage the group's particle descriptors and P matrices in shared memory (canonical child order a <= b)
template <typename Extra>
__device__ __forceinline__ void stage_group(const int32_t* order, const int32_t* lsrc, const int32_t* rsrc,
                                            const double* P, int64_t j0, int nj, int* s_k, int* s_a, int* s_b,
                                            double (*sP)[32], Extra&& extra) {
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid < nj) {
    const int k = order ? order[j0 + tid] : (int)(j0 + tid);
    const int ls = lsrc[k], rs = rsrc[k];
    const bool sw = ls > rs;
    s_a[tid] = sw ? rs : ls;
    s_b[tid] = sw ? ls : rs;
    s_k[tid] = sw ? ~k : k;  // swap flag in the sign
    extra(tid, k);
  }
  __syncthreads();
  for (int e = tid; e < nj * 32; e += kTileThreads) {
    const int j = e >> 5, i = e & 31;
    const int kk = s_k[j];
    const bool sw = kk < 0;
    const int k = sw ? ~kk : kk;
    sP[j][i] = __ldg(P + (int64_t)k * 32 + (i ^ (sw ? 16 : 0)));
  }
  __syncthreads();
}

constexpr int kFwdSmemBytes = kRMax * 32 * 8 + kRMax * kTileThreads * (8 + 4);

template <bool JC, int SPT, int MINB>
__global__ void __launch_bounds__(kTileThreads, MINB) merge_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double(*sP)[32] = reinterpret_cast<double(*)[32]>(smem_raw);
  double* s_prod = reinterpret_cast<double*>(smem_raw + kRMax * 32 * 8);   // [R][256] running mantissa products
  int* s_exp = reinterpret_cast<int*>(s_prod + kRMax * kTileThreads);      // [R][256] running exponent sums
  __shared__ int s_k[kRMax], s_a[kRMax], s_b[kRMax], s_dst[kRMax];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t count = a.count ? (int64_t)*a.count : a.K;
  const int R = a.R;
  const int64_t total = ((count + R - 1) / R) * a.n_chunks;
  double pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = __ldg(a.pi + j);

  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t g = w / a.n_chunks;
    const int tc = (int)(w - g * a.n_chunks);
    const int64_t j0 = g * R;
    const int nj = (int)min((int64_t)R, count - j0);
    stage_group(a.order, a.lsrc, a.rsrc, a.P, j0, nj, s_k, s_a, s_b, sP,
                [&](int i, int k) { s_dst[i] = a.dst ? a.dst[k] : k; });
    for (int j = 0; j < nj; ++j) {
      s_prod[j * kTileThreads + tid] = 1.0;
      s_exp[j * kTileThreads + tid] = 0;
    }
    const int t_end = min(a.tiles, (tc + 1) * a.tiles_per_item);
    for (int t = tc * a.tiles_per_item; t < t_end; ++t) {
      const int sbase = t * (kTileThreads * SPT) + tid;
      int pa = kNone, pb = kNone;
      d4 La[SPT], Lb[SPT];
      for (int j = 0; j < nj; ++j) {
        const int ds = s_dst[j];
        if (a.skip_unstored && ds < 0) continue;
        const int ca = s_a[j], cb = s_b[j];
        if (ca != pa) {
          const ChildRef c = child_ref(ca, a.codes, a.codes_stride, a.pool, a.slot_sites);
#pragma unroll
          for (int q = 0; q < SPT; ++q) {
            const int s = sbase + q * kTileThreads;
            if (s < a.n_sites) La[q] = load_child(c, s);
          }
          pa = ca;
        }
        if (cb != pb) {
          const ChildRef c = child_ref(cb, a.codes, a.codes_stride, a.pool, a.slot_sites);
#pragma unroll
          for (int q = 0; q < SPT; ++q) {
            const int s = sbase + q * kTileThreads;
            if (s < a.n_sites) Lb[q] = load_child(c, s);
          }
          pb = cb;
        }
        Trans<JC> Pa, Pb;
        Pa.load(sP[j]);
        Pb.load(sP[j] + 16);
        double* out = ds < 0 ? nullptr : a.pool + (int64_t)ds * a.slot_sites * 4;
        double pr = s_prod[j * kTileThreads + tid];
        int ex = s_exp[j * kTileThreads + tid];
#pragma unroll
        for (int q = 0; q < SPT; ++q) {
          const int s = sbase + q * kTileThreads;
          if (s < a.n_sites) {
            const d4 lp = Pa.apply(La[q]), rp = Pb.apply(Lb[q]);
            d4 nw;
            double x = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              nw.v[i] = lp.v[i] * rp.v[i];
              x = fma(pi[i], nw.v[i], x);
            }
            if (out) st_site(out + (int64_t)s * 4, nw);
            double m;
            int e;
            split_positive(x, m, e);
            pr *= m;
            ex += e;
          }
        }
        s_prod[j * kTileThreads + tid] = pr;
        s_exp[j * kTileThreads + tid] = ex;
      }
    }
    // one log per (thread, particle): sum_s log x_s = log(prod mantissas) + ln2 * sum exponents
    for (int j = 0; j < nj; ++j) {
      if (a.skip_unstored && s_dst[j] < 0) continue;
      const double pr = s_prod[j * kTileThreads + tid];
      const double ex = (double)s_exp[j * kTileThreads + tid];
      double acc = fma(ex, 6.93147180369123816490e-01, fma(ex, 1.90821492927058770002e-10, log(pr)));  // ln2 hi + lo
      acc = warp_sum(acc);
      if (lane == 0) {
        const int kk = s_k[j];
        const int64_t k = kk < 0 ? ~kk : kk;
        a.ell_part[(k * a.n_chunks + tc) * kWarps + wid] = acc;
      }
    }
  }
}

__global__ void ell_reduce_kernel(const double* __restrict__ part, int n_part, int64_t K, double* __restrict__ ell) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double s = 0.0;
  for (int t = 0; t < n_part; ++t) s += part[k * n_part + t];
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
  const int32_t* order;
  const int32_t* count;
  const double* P;
  const double* pi;
  const double* coef;
  int64_t K;
  int n_sites;
  int tiles;
  int tiles_per_item;
  int n_chunks;
  int R;
  double* dP;        // [K][32]
  double* dpi_acc;   // [4] accumulated over all particles, or null
  int skip_zero;
  double skip_below;  // |coefficient| at or below this counts as zero (0: exact zeros only)
};

template <int SPT>
__device__ __forceinline__ void flush_adjoint(int src, double* gpool, int64_t slot_sites, int sbase, int n_sites,
                                              const d4 (&G)[SPT]) {
  if (src < 0) return;  // leaves (and kNone) have no adjoint buffer
  double* g = gpool + (int64_t)src * slot_sites * 4;
#pragma unroll
  for (int q = 0; q < SPT; ++q) {
    const int s = sbase + q * kTileThreads;
    if (s < n_sites) {
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(g + (int64_t)s * 4 + i, G[q].v[i]);
    }
  }
}

template <bool JC, int SPT>
__global__ void __launch_bounds__(kTileThreads, JC ? 2 : 1) merge_bwd_kernel(const BwdArgs a) {
  __shared__ __align__(16) double sP[kRMax][32];
  __shared__ double s_c[kRMax];
  __shared__ int s_k[kRMax], s_a[kRMax], s_b[kRMax], s_g[kRMax];
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t count = a.count ? (int64_t)*a.count : a.K;
  const int R = a.R;
  const int64_t total = ((count + R - 1) / R) * a.n_chunks;
  double pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = __ldg(a.pi + j);

  double dpi[4] = {0.0, 0.0, 0.0, 0.0};  // d/dpi is a sum over ALL particles: reduced once per CTA, not per particle
  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t g = w / a.n_chunks;
    const int tc = (int)(w - g * a.n_chunks);
    const int64_t j0 = g * R;
    const int nj = (int)min((int64_t)R, count - j0);
    stage_group(a.order, a.lsrc, a.rsrc, a.P, j0, nj, s_k, s_a, s_b, sP, [&](int i, int k) {
      s_c[i] = a.coef[k];
      s_g[i] = a.gsrc ? a.gsrc[k] : -1;
    });
    const int t_end = min(a.tiles, (tc + 1) * a.tiles_per_item);
    for (int t = tc * a.tiles_per_item; t < t_end; ++t) {
    const int sbase = t * (kTileThreads * SPT) + tid;
    int pa = kNone, pb = kNone;
    d4 La[SPT], Lb[SPT], Ga[SPT], Gb[SPT];
    for (int j = 0; j < nj; ++j) {
      const double c = s_c[j];
      const int gs = s_g[j];
      if (a.skip_zero && fabs(c) <= a.skip_below && gs < 0) continue;  // (numerically) zero adjoint: nothing to propagate
      const int ca = s_a[j], cb = s_b[j];
      if (ca != pa) {
        flush_adjoint<SPT>(pa, a.gpool, a.slot_sites, sbase, a.n_sites, Ga);
        const ChildRef cr = child_ref(ca, a.codes, a.codes_stride, a.pool, a.slot_sites);
#pragma unroll
        for (int q = 0; q < SPT; ++q) {
          const int s = sbase + q * kTileThreads;
          if (s < a.n_sites) La[q] = load_child(cr, s);
          Ga[q] = zero4();
        }
        pa = ca;
      }
      if (cb != pb) {
        flush_adjoint<SPT>(pb, a.gpool, a.slot_sites, sbase, a.n_sites, Gb);
        const ChildRef cr = child_ref(cb, a.codes, a.codes_stride, a.pool, a.slot_sites);
#pragma unroll
        for (int q = 0; q < SPT; ++q) {
          const int s = sbase + q * kTileThreads;
          if (s < a.n_sites) Lb[q] = load_child(cr, s);
          Gb[q] = zero4();
        }
        pb = cb;
      }
      Trans<JC> Pa, Pb;
      Pa.load(sP[j]);
      Pb.load(sP[j] + 16);
      const double* gnew = gs < 0 ? nullptr : a.gpool + (int64_t)gs * a.slot_sites * 4;
      const int kk = s_k[j];
      const bool sw = kk < 0;
      const int64_t k = sw ? ~kk : kk;

      double acc[JC ? 4 : 32];
#pragma unroll
      for (int i = 0; i < (JC ? 4 : 32); ++i) acc[i] = 0.0;
#pragma unroll
      for (int q = 0; q < SPT; ++q) {
        const int s = sbase + q * kTileThreads;
        if (s < a.n_sites) {
          const d4 gin = gnew ? ld_site(gnew + (int64_t)s * 4) : zero4();
          const d4 lp = Pa.apply(La[q]), rp = Pb.apply(Lb[q]);
          double nw[4], x = 0.0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            nw[i] = lp.v[i] * rp.v[i];
            x = fma(pi[i], nw[i], x);
          }
          const double inv = c / x;
          d4 gl, gr;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double gi = fma(inv, pi[i], gin.v[i]);
            dpi[i] = fma(inv, nw[i], dpi[i]);  // d ell / d pi_i = new_i / x
            gl.v[i] = gi * rp.v[i];
            gr.v[i] = gi * lp.v[i];
          }
          if (ca >= 0) {
            const d4 tt = Pa.apply_t(gl);
#pragma unroll
            for (int i = 0; i < 4; ++i) Ga[q].v[i] += tt.v[i];
          }
          if (cb >= 0) {
            const d4 tt = Pb.apply_t(gr);
#pragma unroll
            for (int i = 0; i < 4; ++i) Gb[q].v[i] += tt.v[i];
          }
          if (JC) {
            // dP enters only through (sum_i dP_ii, sum_{i!=j} dP_ij) with dP_ij = L_i g_j
            double dl = 0.0, dr = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              dl = fma(La[q].v[i], gl.v[i], dl);
              dr = fma(Lb[q].v[i], gr.v[i], dr);
            }
            const double sl = ((La[q].v[0] + La[q].v[1]) + (La[q].v[2] + La[q].v[3])) * ((gl.v[0] + gl.v[1]) + (gl.v[2] + gl.v[3]));
            const double sr = ((Lb[q].v[0] + Lb[q].v[1]) + (Lb[q].v[2] + Lb[q].v[3])) * ((gr.v[0] + gr.v[1]) + (gr.v[2] + gr.v[3]));
            acc[0] += dl;
            acc[1] += sl - dl;
            acc[2] += dr;
            acc[3] += sr - dr;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                acc[i * 4 + jj] = fma(La[q].v[i], gl.v[jj], acc[i * 4 + jj]);
                acc[16 + i * 4 + jj] = fma(Lb[q].v[i], gr.v[jj], acc[16 + i * 4 + jj]);
              }
          }
        }
      }
      // per-particle reduction over the warp, one atomic per value per warp (a/b are in canonical order: undo swap)
      if (JC) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);  // totals land in lane 0
        if (lane == 0) {
          // JC layout inside dP[k][32]: [0]=sum diag(dP_l), [1]=sum offdiag(dP_l), [16],[17] same for right
          const int sa = sw ? 16 : 0, sb = sw ? 0 : 16;
          atomicAdd(a.dP + k * 32 + sa, acc[0]);
          atomicAdd(a.dP + k * 32 + sa + 1, acc[1]);
          atomicAdd(a.dP + k * 32 + sb, acc[2]);
          atomicAdd(a.dP + k * 32 + sb + 1, acc[3]);
        }
      } else {
        double(&v32)[32] = *reinterpret_cast<double(*)[32]>(acc);
        warp_transpose_sum32(v32, lane);
        atomicAdd(a.dP + k * 32 + (lane ^ (sw ? 16 : 0)), v32[0]);
      }
    }
    flush_adjoint<SPT>(pa, a.gpool, a.slot_sites, sbase, a.n_sites, Ga);
    flush_adjoint<SPT>(pb, a.gpool, a.slot_sites, sbase, a.n_sites, Gb);
    }
  }
  if (a.dpi_acc) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dpi[i] = warp_sum(dpi[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (dpi[i] != 0.0) atomicAdd(a.dpi_acc + i, dpi[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The per-site pass of the reverse sweep when only a handful of particles per rank event carry an adjoint (the rule with
// ESS ~ 1: one or two lineages).  Sites are independent, so a THREAD owns its site through the whole pass: it recomputes
// the consumed nodes of the visited particles in event order (reading what it wrote itself an event earlier), then walks
// the events backwards accumulating child adjoints with plain read-modify-writes -- no launch per rank event, no grid
// barrier, no atomics on the node adjoints.  The per-particle 4x4 adjoints dP = sum_s L (x) (g * rp) are reduced per CTA
// (halving-butterfly transpose in the warp, one shared-memory round) and added to the table once per CTA.
// Same arithmetic per site as merge_fwd_kernel / merge_bwd_kernel.
// ---------------------------------------------------------------------------------------------
struct SparseArgs {
  const uint8_t* codes;
  int64_t codes_stride;
  double* pool;
  double* gpool;
  int64_t slot_sites;
  int n_sites, N, recompute, skip_zero;
  int64_t K;
  double skip_below;
  const int32_t* order;   // [N-1][K] visit lists
  const int32_t* count;   // [N-1]
  const int32_t* lsrc;    // [N-1][K] child / adjoint / output slots of every event
  const int32_t* rsrc;
  const int32_t* gsrc;
  const int32_t* dst;
  const double* P;        // [N-1][K][32]
  const double* pi;
  const double* coef;     // [N-1][K]
  double* dP;             // [N-1][K][32]
  double* dpi_acc;
};

template <bool JC>
__global__ void __launch_bounds__(kTileThreads) bwd_sparse_kernel(const SparseArgs a) {
  __shared__ double s_red[32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int s = blockIdx.x * kTileThreads + tid;
  const bool valid = s < a.n_sites;
  const int64_t K = a.K;
  double pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = __ldg(a.pi + j);
  if (tid < 32) s_red[tid] = 0.0;
  // ---- the adjoint slots of the consumed nodes, this thread's site
  for (int r = 0; r < a.N - 1; ++r) {
    const int cnt = a.count[r];
    for (int i = 0; i < cnt; ++i) {
      const int64_t e = (int64_t)r * K + a.order[(int64_t)r * K + i];
      const int g = a.gsrc[e];
      if (g >= 0 && valid) st_site(a.gpool + ((int64_t)g * a.slot_sites + s) * 4, zero4());
    }
  }
  // ---- forward: the consumed nodes of the visited particles, in event order
  if (a.recompute) {
    for (int r = 0; r < a.N - 1; ++r) {
      const int cnt = a.count[r];
      for (int i = 0; i < cnt; ++i) {
        const int64_t e = (int64_t)r * K + a.order[(int64_t)r * K + i];
        const int d = a.dst[e];
        if (d < 0 || !valid) continue;
        const ChildRef ra = child_ref(a.lsrc[e], a.codes, a.codes_stride, a.pool, a.slot_sites);
        const ChildRef rb = child_ref(a.rsrc[e], a.codes, a.codes_stride, a.pool, a.slot_sites);
        Trans<JC> Pa, Pb;
        Pa.load(a.P + e * 32);
        Pb.load(a.P + e * 32 + 16);
        const d4 lp = Pa.apply(load_child(ra, s)), rp = Pb.apply(load_child(rb, s));
        d4 nw;
#pragma unroll
        for (int j = 0; j < 4; ++j) nw.v[j] = lp.v[j] * rp.v[j];
        st_site(a.pool + ((int64_t)d * a.slot_sites + s) * 4, nw);
      }
    }
  }
  // ---- backward
  double dpi[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r = a.N - 2; r >= 0; --r) {
    const int cnt = a.count[r];
    for (int i = 0; i < cnt; ++i) {
      const int64_t e = (int64_t)r * K + a.order[(int64_t)r * K + i];
      const double c = a.coef[e];
      const int gs = a.gsrc[e];
      if (a.skip_zero && fabs(c) <= a.skip_below && gs < 0) continue;  // (numerically) zero adjoint: nothing to propagate
      const int ca = a.lsrc[e], cb = a.rsrc[e];
      double acc[JC ? 4 : 32];
#pragma unroll
      for (int j = 0; j < (JC ? 4 : 32); ++j) acc[j] = 0.0;
      if (valid) {
        const ChildRef ra = child_ref(ca, a.codes, a.codes_stride, a.pool, a.slot_sites);
        const ChildRef rb = child_ref(cb, a.codes, a.codes_stride, a.pool, a.slot_sites);
        const d4 La = load_child(ra, s), Lb = load_child(rb, s);
        Trans<JC> Pa, Pb;
        Pa.load(a.P + e * 32);
        Pb.load(a.P + e * 32 + 16);
        const d4 gin = gs >= 0 ? ld_site(a.gpool + ((int64_t)gs * a.slot_sites + s) * 4) : zero4();
        const d4 lp = Pa.apply(La), rp = Pb.apply(Lb);
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
          const double gi = fma(inv, pi[j], gin.v[j]);
          dpi[j] = fma(inv, nw[j], dpi[j]);  // d ell / d pi_j = new_j / x
          gl.v[j] = gi * rp.v[j];
          gr.v[j] = gi * lp.v[j];
        }
        if (ca >= 0) {   // this thread is the only one that touches site s of any slot
          double* g = a.gpool + ((int64_t)ca * a.slot_sites + s) * 4;
          const d4 tt = Pa.apply_t(gl);
          d4 cur = ld_site(g);
#pragma unroll
          for (int j = 0; j < 4; ++j) cur.v[j] += tt.v[j];
          st_site(g, cur);
        }
        if (cb >= 0) {
          double* g = a.gpool + ((int64_t)cb * a.slot_sites + s) * 4;
          const d4 tt = Pb.apply_t(gr);
          d4 cur = ld_site(g);
#pragma unroll
          for (int j = 0; j < 4; ++j) cur.v[j] += tt.v[j];
          st_site(g, cur);
        }
        if (JC) {
          double dl = 0.0, dr = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            dl = fma(La.v[j], gl.v[j], dl);
            dr = fma(Lb.v[j], gr.v[j], dr);
          }
          const double sl = ((La.v[0] + La.v[1]) + (La.v[2] + La.v[3])) * ((gl.v[0] + gl.v[1]) + (gl.v[2] + gl.v[3]));
          const double sr = ((Lb.v[0] + Lb.v[1]) + (Lb.v[2] + Lb.v[3])) * ((gr.v[0] + gr.v[1]) + (gr.v[2] + gr.v[3]));
          acc[0] = dl;
          acc[1] = sl - dl;
          acc[2] = dr;
          acc[3] = sr - dr;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              acc[j * 4 + jj] = La.v[j] * gl.v[jj];
              acc[16 + j * 4 + jj] = Lb.v[j] * gr.v[jj];
            }
        }
      }
      // the CTA's share of this particle's dP: warp totals into shared memory, one add per value to the table
      if (JC) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
        if (lane == 0) {
          atomicAdd(&s_red[0], acc[0]);
          atomicAdd(&s_red[1], acc[1]);
          atomicAdd(&s_red[16], acc[2]);
          atomicAdd(&s_red[17], acc[3]);
        }
      } else {
        double(&v32)[32] = *reinterpret_cast<double(*)[32]>(acc);
        warp_transpose_sum32(v32, lane);
        atomicAdd(&s_red[lane], v32[0]);
      }
      __syncthreads();
      if (tid < 32) {
        const double v = s_red[tid];
        if (v != 0.0) atomicAdd(a.dP + e * 32 + tid, v);
        s_red[tid] = 0.0;
      }
      __syncthreads();
    }
  }
  if (a.dpi_acc) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dpi[j] = warp_sum(dpi[j]);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (dpi[j] != 0.0) atomicAdd(a.dpi_acc + j, dpi[j]);
    }
  }
}


constexpr int kSptBwdJC = 2;
constexpr int kSptBwdGeneral = 2;  // measured at 64x10kx65,536 dense: 1078 ms (SPT 2) vs 1278 ms (SPT 1)

constexpr int64_t kTargetItems = 148 * 16;  // work items wanted per launch (persistent grid cap)

int pick_group(int64_t K, int tiles) {
  // enough work items to fill 148 SMs several times over; larger groups amortise shared children
  int64_t R = (K * tiles) / (148 * 32);
  if (R < 1) R = 1;
  if (R > kRMax) R = kRMax;
  return (int)R;
}

// split the tiles of a group into n_chunks work items only when the groups alone cannot fill the machine
void pick_chunks(int64_t K, int R, int tiles, int* tiles_per_item, int* n_chunks) {
  const int64_t groups = (K + R - 1) / R;
  int64_t nc = (kTargetItems + groups - 1) / groups;
  if (nc < 1) nc = 1;
  if (nc > tiles) nc = tiles;
  const int tpi = (int)((tiles + nc - 1) / nc);
  *tiles_per_item = tpi;
  *n_chunks = (tiles + tpi - 1) / tpi;
}

unsigned pick_grid(int64_t K, int R, int n_chunks) {
  const int64_t total = ((K + R - 1) / R) * n_chunks;
  return (unsigned)(total < kTargetItems ? (total > 0 ? total : 1) : kTargetItems);
}

}  // namespace

int merge_fwd_tiles(int n_sites) { return (n_sites + kTileThreads - 1) / kTileThreads; }  // upper bound (SPT = 1)
int merge_ell_parts(int n_sites) { return merge_fwd_tiles(n_sites) * kWarps; }

int launch_merge_fwd(const uint8_t* codes, int64_t codes_stride, double* pool, int64_t slot_sites, const int32_t* lsrc,
                     const int32_t* rsrc, const int32_t* dst, const int32_t* order, const int32_t* count,
                     const double* P, const double* pi, int64_t K, int64_t n_active, int n_sites, int jc,
                     int skip_unstored, double* ell_part, int* n_parts, cudaStream_t st) {
  // n_active: host-side knowledge of *count (exact), or < 0 when only the device knows (then K bounds the grid)
  if (n_parts) *n_parts = 0;
  if (K <= 0 || n_sites <= 0 || n_active == 0) return VCSMC_OK;
  const int64_t Kw = n_active > 0 ? n_active : K;
  // Measured on B200 at 64 x 10k x 65,536 (ms per sweep of merge_fwd): general Q 265 (SPT 2, 2 CTAs/SM), 266 (SPT 1, 3),
  // 264 (SPT 1, 4), 283 (SPT 1, 2); JC 208 / 213 / 214 / 211 -- occupancy is not the limiter.  Default: general -> SPT 1
  // with 4 CTAs/SM (also best at 27 x 1949 x 8192), JC -> SPT 2 with 2 CTAs/SM.  VCSMC_FWD_VARIANT=0..3 overrides.
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("VCSMC_FWD_VARIANT");
    forced = e ? atoi(e) : -1;
    if (forced < -1 || forced > 3) forced = -1;
    VCSMC_CUDA(cudaFuncSetAttribute(merge_fwd_kernel<true, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    VCSMC_CUDA(cudaFuncSetAttribute(merge_fwd_kernel<false, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    VCSMC_CUDA(cudaFuncSetAttribute(merge_fwd_kernel<true, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    VCSMC_CUDA(cudaFuncSetAttribute(merge_fwd_kernel<false, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    VCSMC_CUDA(cudaFuncSetAttribute(merge_fwd_kernel<true, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    VCSMC_CUDA(cudaFuncSetAttribute(merge_fwd_kernel<false, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    VCSMC_CUDA(cudaFuncSetAttribute(merge_fwd_kernel<true, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    VCSMC_CUDA(cudaFuncSetAttribute(merge_fwd_kernel<false, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
  }
  const int variant = forced >= 0 ? forced : (jc ? 0 : 2);
  const int spt = variant >= 1 ? 1 : 2;
  FwdArgs a;
  a.codes = codes; a.codes_stride = codes_stride; a.pool = pool; a.slot_sites = slot_sites;
  a.lsrc = lsrc; a.rsrc = rsrc; a.dst = dst; a.order = order; a.count = count; a.P = P; a.pi = pi; a.K = K;
  a.n_sites = n_sites; a.tiles = (n_sites + kTileThreads * spt - 1) / (kTileThreads * spt); a.R = pick_group(Kw, a.tiles);
  a.skip_unstored = skip_unstored; a.ell_part = ell_part;
  pick_chunks(Kw, a.R, a.tiles, &a.tiles_per_item, &a.n_chunks);
  if (n_parts) *n_parts = a.n_chunks * kWarps;
  const unsigned grid = pick_grid(Kw, a.R, a.n_chunks);
#define VCSMC_FWD_LAUNCH(SPT, MINB)                                                                       \
  do {                                                                                                    \
    if (jc) merge_fwd_kernel<true, SPT, MINB><<<grid, kTileThreads, kFwdSmemBytes, st>>>(a);              \
    else merge_fwd_kernel<false, SPT, MINB><<<grid, kTileThreads, kFwdSmemBytes, st>>>(a);                \
  } while (0)
  switch (variant) {
    case 1: VCSMC_FWD_LAUNCH(1, 3); break;
    case 2: VCSMC_FWD_LAUNCH(1, 4); break;
    case 3: VCSMC_FWD_LAUNCH(1, 2); break;
    default: VCSMC_FWD_LAUNCH(2, 2); break;
  }
#undef VCSMC_FWD_LAUNCH
  VCSMC_LAUNCH_CHECK("merge_fwd_kernel");
  return VCSMC_OK;
}

int launch_ell_reduce(const double* ell_part, int n_part, int64_t K, double* ell, cudaStream_t st) {
  ell_reduce_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(ell_part, n_part, K, ell);
  VCSMC_LAUNCH_CHECK("ell_reduce_kernel");
  return VCSMC_OK;
}

int launch_merge_bwd(const uint8_t* codes, int64_t codes_stride, const double* pool, double* gpool, int64_t slot_sites,
                     const int32_t* lsrc, const int32_t* rsrc, const int32_t* gsrc, const int32_t* order,
                     const int32_t* count, const double* P, const double* pi, const double* coef, int64_t K,
                     int64_t n_active, int n_sites, int jc, int skip_zero, double skip_below, double* dP, double* dpi_acc, cudaStream_t st) {
  if (K <= 0 || n_sites <= 0 || n_active == 0) return VCSMC_OK;
  const int64_t Kw = n_active > 0 ? n_active : K;
  BwdArgs a;
  a.codes = codes; a.codes_stride = codes_stride; a.pool = pool; a.gpool = gpool; a.slot_sites = slot_sites;
  a.lsrc = lsrc; a.rsrc = rsrc; a.gsrc = gsrc; a.order = order; a.count = count; a.P = P; a.pi = pi;
  a.coef = coef; a.K = K; a.n_sites = n_sites; a.dP = dP; a.dpi_acc = dpi_acc; a.skip_zero = skip_zero; a.skip_below = skip_below;
  static int gspt = -1;  // tuning knob (debug): VCSMC_BWD_SPT = 1 | 2 sites per thread in the general-Q reverse merge
  if (gspt < 0) {
    const char* e = getenv("VCSMC_BWD_SPT");
    gspt = e ? atoi(e) : kSptBwdGeneral;
    if (gspt != 1 && gspt != 2) gspt = kSptBwdGeneral;
  }
  const int spt = jc ? kSptBwdJC : gspt;
  a.tiles = (n_sites + kTileThreads * spt - 1) / (kTileThreads * spt);
  a.R = pick_group(Kw, a.tiles);
  pick_chunks(Kw, a.R, a.tiles, &a.tiles_per_item, &a.n_chunks);
  const unsigned grid = pick_grid(Kw, a.R, a.n_chunks);
  if (jc) merge_bwd_kernel<true, kSptBwdJC><<<grid, kTileThreads, 0, st>>>(a);
  else if (spt == 2) merge_bwd_kernel<false, 2><<<grid, kTileThreads, 0, st>>>(a);
  else merge_bwd_kernel<false, 1><<<grid, kTileThreads, 0, st>>>(a);
  VCSMC_LAUNCH_CHECK("merge_bwd_kernel");
  return VCSMC_OK;
}

int launch_bwd_sparse(const uint8_t* codes, int64_t codes_stride, double* pool, double* gpool, int64_t slot_sites, int n_sites,
                      int N, int64_t K, int recompute, int jc, int skip_zero, double skip_below, const int32_t* order,
                      const int32_t* count, const int32_t* lsrc, const int32_t* rsrc, const int32_t* gsrc, const int32_t* dst,
                      const double* P, const double* pi, const double* coef, double* dP, double* dpi_acc, cudaStream_t st) {
  if (n_sites <= 0 || N < 2) return VCSMC_OK;
  SparseArgs a;
  a.codes = codes; a.codes_stride = codes_stride; a.pool = pool; a.gpool = gpool; a.slot_sites = slot_sites;
  a.n_sites = n_sites; a.N = N; a.recompute = recompute; a.skip_zero = skip_zero; a.K = K; a.skip_below = skip_below;
  a.order = order; a.count = count; a.lsrc = lsrc; a.rsrc = rsrc; a.gsrc = gsrc; a.dst = dst; a.P = P; a.pi = pi;
  a.coef = coef; a.dP = dP; a.dpi_acc = dpi_acc;
  const unsigned grid = (unsigned)((n_sites + kTileThreads - 1) / kTileThreads);
  if (jc) bwd_sparse_kernel<true><<<grid, kTileThreads, 0, st>>>(a);
  else bwd_sparse_kernel<false><<<grid, kTileThreads, 0, st>>>(a);
  VCSMC_LAUNCH_CHECK("bwd_sparse_kernel");
  return VCSMC_OK;
}

}  // namespace vcsmc
