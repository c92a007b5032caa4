// VNCSMC: the nested look-ahead proposal of vncsmc.py:295-416 ("--nested=true" / "--twisting").
//
// For every particle, every pair (r1 < r2) of its n live subtrees and M sub-samples, the reference samples a branch
// pair, merges the two subtrees, and scores   [ell(merged) + prior] - [ell(r1) + prior] - [ell(r2) + prior]; one
// categorical draw over the C(n,2)*M log-softmaxed scores picks the pair and its branches (vncsmc.py:298-320).
// The reference tiles [K,S,4] tensors C*M times in a Python-built double while-loop; here one CTA owns one particle:
// the n root vectors of a site tile are staged in shared memory ONCE and every (pair, m) combination is evaluated
// on them read-only.  A THREAD owns a combination: its two 4x4 P matrices and the running product of the
// site likelihoods (mantissa product + integer exponent sum, one log at the end) stay in registers, so there is no
// cross-thread reduction at all.  The kernel is FP64-ALU-bound, not HBM-bound.
#include <limits.h>

#include "launch.h"
#include "smc_device.cuh"

namespace vcsmc {
namespace {

constexpr int kLookThreads = 256;
constexpr int kLookTile = 128;   // sites per staged tile when n <= 48: n * tile * 32 B of shared memory must fit 192 KB
constexpr int kLookSmem = 48 * kLookTile * 32;
constexpr int kMaxNestedRoots = kMaxRoots;   // larger forests stage shorter tiles (slower: every round of 256 (pair, sub-sample) combinations re-reads the roots)

__device__ __forceinline__ void pair_of(int t, int n, int& r1, int& r2) {
  // r1-major enumeration of vncsmc.py:324-377: t = sum_{i<r1} (n-1-i) + (r2 - r1 - 1)
  int a = 0, rem = t;
  while (rem >= n - 1 - a) {
    rem -= n - 1 - a;
    ++a;
  }
  r1 = a;
  r2 = a + 1 + rem;
}

__device__ __forceinline__ void look_uniforms(const double* u_bl, const double* u_br, uint64_t seed, int r, int64_t k,
                                              int64_t K, int M, int c, double& ul, double& ur) {
  if (u_bl) {
    const int t = c / M, m = c - t * M;
    const int64_t i = (int64_t)t * M * K + (int64_t)m * K + k;  // [C][M*K], column m*K + k (vncsmc.py:346-353)
    ul = u_bl[i];
    ur = u_br[i];
  } else {
    uint32_t ctr[4] = {(uint32_t)k, (uint32_t)((uint64_t)k >> 32) | ((uint32_t)r << 8), 3u, (uint32_t)c};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double tiny = 2.2250738585072014e-308;
    ul = fmax(u64_to_unit_f64(ctr[0], ctr[1]), tiny);
    ur = fmax(u64_to_unit_f64(ctr[2], ctr[3]), tiny);
  }
}

__device__ __forceinline__ void transition_of(const double* Q, double t, int jc, double (&P)[16]) {
  if (jc) {
    const double o = -0.25 * expm1(-t), d = 0.25 + 0.75 * exp(-t);
#pragma unroll
    for (int e = 0; e < 16; ++e) P[e] = (e % 5 == 0) ? d : o;
  } else {
    M4 A;
#pragma unroll
    for (int e = 0; e < 16; ++e) A.a[e] = __ldg(Q + e) * t;
    const M4 X = m4_expm(A);
#pragma unroll
    for (int e = 0; e < 16; ++e) P[e] = X.a[e];
  }
}

struct InheritArgs {
  int r, n, N;
  int64_t K;
  const double* cdf;
  const double* u_res;
  const int32_t* ids_prev;
  const int32_t* cnt_prev;
  const int32_t* slot_prev;
  int32_t* ids;
  int32_t* cnt;
  int32_t* slot;
  int32_t* rows_all;  // [K][N] of this rank event (kept for the reverse sweep) or null
  const double* LL_prev;
  int32_t* anc;
  double* ll_tilde;
};

// resample (vncsmc.py:283-293,:418-425): every particle inherits its ancestor's whole forest row
__global__ void nested_inherit_kernel(const InheritArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t k = i / a.n;
  const int p = (int)(i - k * a.n);
  if (k >= a.K) return;
  const bool first = a.r == 0;
  const int idx = first ? (int)k : upper_bound_cdf(a.cdf, a.K, a.u_res[k] * a.cdf[a.K - 1]);
  const int id = first ? p : a.ids_prev[(int64_t)idx * a.N + p];
  a.ids[k * a.N + p] = id;
  a.cnt[k * a.N + p] = first ? 1 : a.cnt_prev[(int64_t)idx * a.N + p];
  a.slot[k * a.N + p] = first ? -1 : a.slot_prev[(int64_t)idx * a.N + p];
  if (a.rows_all) a.rows_all[k * a.N + p] = id;
  if (p == 0) {
    a.anc[k] = idx;
    a.ll_tilde[k] = first ? log(1.0 / (double)a.K) : a.LL_prev[idx];
  }
}

struct LookArgs {
  int r, n, N, M, jc, gc, S, tile;
  int64_t K;
  const int32_t* ids;
  const int32_t* cnt;
  const int32_t* slot;
  const uint8_t* codes;
  int64_t codes_stride;
  const double* pool;
  int64_t slot_sites;
  const double* ell_node;
  const double* ldf;
  const double* Q;
  const double* pi;
  const double* lam_l;
  const double* lam_r;
  const double* u_bl;
  const double* u_br;
  uint64_t seed;
  double* pot;  // [K][C*M] raw potentials
  double share; // site sharding: fraction of the site-independent terms this rank contributes before the all-reduce (1 otherwise)
};

template <bool JC>
__global__ void __launch_bounds__(kLookThreads) lookahead_kernel(const LookArgs a) {
  extern __shared__ __align__(32) double roots[];  // [n][tile][4]
  __shared__ int s_ref[kMaxNestedRoots];            // < 0: leaf -(ref+1); else pool slot
  __shared__ double s_ell[kMaxNestedRoots], s_prior[kMaxNestedRoots];
  __shared__ int s_cnt[kMaxNestedRoots];
  const int tid = threadIdx.x;
  const int64_t k = blockIdx.x;
  const int n = a.n, N = a.N, M = a.M;
  const int combos = n * (n - 1) / 2 * M;
  if (tid < n) {
    const int id = a.ids[k * N + tid];
    s_ref[tid] = id < N ? -(id + 1) : (a.gc ? a.slot[k * N + tid] : id - N);
    s_ell[tid] = a.ell_node[id];
    const int c = a.cnt[k * N + tid];
    s_cnt[tid] = c;
    s_prior[tid] = -a.ldf[2 * max(c, 2) - 3];
  }
  double pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = __ldg(a.pi + j);
  const double laml = a.lam_l[a.r], lamr = a.lam_r[a.r];
  // G site-slices when there are fewer combinations than threads
  const int G = combos >= kLookThreads ? 1 : kLookThreads / combos;
  const int per_round = G == 1 ? kLookThreads : combos;
  const int g = G == 1 ? 0 : tid / combos;
  const bool worker = G == 1 ? true : (g < G);
  __syncthreads();

  for (int c0 = 0; c0 < combos; c0 += per_round) {
    const int c = c0 + (G == 1 ? tid : tid % combos);
    const bool on = worker && c < combos;
    int r1 = 0, r2 = 1;
    // The site likelihood of a merge is a bilinear form of the two children (score.cu):
    //   x[s] = sum_{j,m} L1[s][j] L2[s][m] Mx[j][m],  Mx[j][m] = sum_i pi_i Pl[j][i] Pr[m][i]
    // -- 20 DFMA per site (t = Mx L2, then L1 . t) instead of 36, and 16 registers instead of the two P matrices.
    double Mx[16];
    if (on) {
      pair_of(c / M, n, r1, r2);
      double ul, ur;
      look_uniforms(a.u_bl, a.u_br, a.seed, a.r, k, a.K, M, c, ul, ur);
      double Pl[16], Pr[16];
      transition_of(a.Q, -log(ul) / laml, JC, Pl);
      transition_of(a.Q, -log(ur) / lamr, JC, Pr);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          double v = pi[0] * Pl[j * 4] * Pr[m * 4];
#pragma unroll
          for (int i = 1; i < 4; ++i) v = fma(pi[i] * Pl[j * 4 + i], Pr[m * 4 + i], v);
          Mx[j * 4 + m] = v;
        }
    }
    double pr = 1.0;
    int ex = 0;
    const int tile = a.tile;
    for (int s0 = 0; s0 < a.S; s0 += tile) {
      const int nt = min(tile, a.S - s0);
      __syncthreads();
      for (int e = tid; e < n * tile; e += kLookThreads) {
        const int p = e / tile, sl = e - p * tile;
        if (sl < nt) {
          const int ref = s_ref[p];
          const d4 v = ref < 0 ? leaf_site(__ldg(a.codes + (int64_t)(-ref - 1) * a.codes_stride + s0 + sl))
                               : ld_site(a.pool + ((int64_t)ref * a.slot_sites + s0 + sl) * 4);
          *reinterpret_cast<d4*>(roots + ((int64_t)p * tile + sl) * 4) = v;
        }
      }
      __syncthreads();
      if (on) {
        const double* A1 = roots + (int64_t)r1 * tile * 4;
        const double* A2 = roots + (int64_t)r2 * tile * 4;
        for (int sl = g; sl < nt; sl += G) {
          const d4 L1 = *reinterpret_cast<const d4*>(A1 + sl * 4), L2 = *reinterpret_cast<const d4*>(A2 + sl * 4);
          double x = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            double t = Mx[j * 4] * L2.v[0];
#pragma unroll
            for (int m = 1; m < 4; ++m) t = fma(Mx[j * 4 + m], L2.v[m], t);
            x = fma(L1.v[j], t, x);
          }
          const int hi = __double2hiint(x);
          const int e = (hi >> 20) & 0x7ff;
          const bool normal = (e != 0) && (e != 0x7ff) && (hi >= 0);
          pr *= normal ? __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x)) : x;
          ex += normal ? e - 1023 : 0;
        }
        // a thread multiplies up to 128 mantissas in [1,2) per tile: fold the product's exponent back once per tile
        const int hi = __double2hiint(pr);
        const int e = (hi >> 20) & 0x7ff;
        if (e != 0 && e != 0x7ff && hi >= 0) {
          pr = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(pr));
          ex += e - 1023;
        }
      }
    }
    double val = 0.0;
    if (on) {
      const double e = (double)ex;
      val = fma(e, 6.93147180369123816490e-01, fma(e, 1.90821492927058770002e-10, log(pr)));
    }
    if (G > 1) {
      // fixed-order sum of the site slices: reuse the root buffer as scratch [G][combos]
      __syncthreads();
      if (on) roots[g * combos + c] = val;
      __syncthreads();
      if (on && g == 0) {
        double t = 0.0;
        for (int q = 0; q < G; ++q) t += roots[q * combos + c];
        val = t;
      }
    }
    if (on && g == 0) {
      const int cm = s_cnt[r1] + s_cnt[r2];
      const double prior12 = -a.ldf[2 * max(cm, 2) - 3];
      // vncsmc.py:363-365: joint(merged) - joint(left) - joint(right)
      if (a.share == 1.0)
        a.pot[k * combos + c] = ((val + prior12) - (s_ell[r1] + s_prior[r1])) - (s_ell[r2] + s_prior[r2]);
      else   // this rank's sites only: the all-reduce completes the site sum, one rank contributes the site-free terms
        a.pot[k * combos + c] = val + a.share * ((prior12 - (s_ell[r1] + s_prior[r1])) - (s_ell[r2] + s_prior[r2]));
    }
  }
}

struct ChooseArgs {
  int r, n, N, M, gc;
  int64_t K;
  double* pot;  // in: raw, out: log-softmax (vncsmc.py:407)
  const double* u_cat;
  const double* u_bl;
  const double* u_br;
  uint64_t seed;
  const double* lam_l;
  const double* lam_r;
  const int32_t* ids;
  const int32_t* cnt;
  const int32_t* slot;
  int32_t* ids_new;
  int32_t* cnt_new;
  int32_t* slot_new;
  int32_t* lref;
  int32_t* rref;
  int32_t* nleaf;
  uint8_t* rempos;
  int32_t* choice;
  double* b_l;
  double* b_r;
  double* t2;
  double* qlog;
  int32_t* lsrc;
  int32_t* rsrc;
  int32_t* dst;
};

// log-softmax of the potentials, one categorical draw per particle (tf.random.categorical, vncsmc.py:298), and the
// forest-row update of extend_partial_state (vncsmc.py:299-320) + state update (:458-470).  One warp per particle.
__global__ void __launch_bounds__(256) nested_choose_kernel(const ChooseArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (k >= a.K) return;
  const int n = a.n, N = a.N, M = a.M;
  const int combos = n * (n - 1) / 2 * M;
  double* pot = a.pot + k * combos;
  double mx = -INFINITY;
  for (int c = lane; c < combos; c += 32) mx = fmax(mx, pot[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  double s = 0.0;
  for (int c = lane; c < combos; c += 32) s += exp(pot[c] - mx);
  s = warp_sum(s);
  s = __shfl_sync(0xffffffffu, s, 0);
  const double lse = mx + log(s);
  for (int c = lane; c < combos; c += 32) pot[c] = pot[c] - lse;
  __syncwarp();
  const double lmax = mx - lse;
  // total of exp(logit - max) in column order (32-wide chunks, shuffle scan, carried offset)
  double carry = 0.0;
  for (int c0 = 0; c0 < combos; c0 += 32) {
    const int c = c0 + lane;
    double w = c < combos ? exp(pot[c] - lmax) : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    carry += __shfl_sync(0xffffffffu, w, 31);
  }
  const double target = a.u_cat[k] * carry;
  int pick = combos - 1;
  carry = 0.0;
  for (int c0 = 0; c0 < combos; c0 += 32) {
    const int c = c0 + lane;
    double w = c < combos ? exp(pot[c] - lmax) : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, c < combos && carry + w > target);
    if (hit) {
      pick = c0 + __ffs(hit) - 1;
      break;
    }
    carry += __shfl_sync(0xffffffffu, w, 31);
  }
  int r1, r2;
  pair_of(pick / M, n, r1, r2);
  const int32_t* io = a.ids + k * N;
  const int32_t* co = a.cnt + k * N;
  const int32_t* so = a.slot + k * N;
  // remaining subtrees in DESCENDING index order (vncsmc.py:302-305), then the new node
  for (int p = lane; p < n - 2; p += 32) {
    int i = n - 1 - p;            // p-th largest index overall ...
    if (i <= r2) --i;             // ... skipping r2 (> r1)
    if (i <= r1) --i;             // ... and r1
    a.rempos[k * (int64_t)(n - 2) + p] = (uint8_t)i;
    a.ids_new[k * N + p] = io[i];
    a.cnt_new[k * N + p] = co[i];
    if (a.gc) a.slot_new[k * N + p] = so[i];
  }
  if (lane == 0) {
    const int64_t e = (int64_t)a.r * a.K + k;
    const int lid = io[r1], rid = io[r2];
    a.ids_new[k * N + n - 2] = (int32_t)(N + e);
    const int nl = co[r1] + co[r2];
    a.cnt_new[k * N + n - 2] = nl;
    a.nleaf[k] = nl;
    a.lref[k] = lid;
    a.rref[k] = rid;
    a.lsrc[k] = lid < N ? -(lid + 1) : (a.gc ? so[r1] : lid - N);
    a.rsrc[k] = rid < N ? -(rid + 1) : (a.gc ? so[r2] : rid - N);
    if (!a.gc) a.dst[k] = (int32_t)e;
    double ul, ur;
    look_uniforms(a.u_bl, a.u_br, a.seed, a.r, k, a.K, M, pick, ul, ur);
    const double bl = -log(ul) / a.lam_l[a.r], br = -log(ur) / a.lam_r[a.r];
    a.b_l[k] = bl;
    a.b_r[k] = br;
    a.t2[2 * k] = bl;
    a.t2[2 * k + 1] = br;
    a.qlog[k] = pot[pick];
    a.choice[k] = pick;
  }
}

// ---------------------------------------------------------------------------------------------
// reverse sweep of the look-ahead: lw_r[k] contains -(pot[k,choice] - lse_k), so with W = dELBO/dlw_r[k]
//   kappa_c = dELBO/d raw_c = W * (softmax_c - [c == choice]).
// Every (active particle, combination) becomes a VIRTUAL merge event (children = roots r1, r2, coefficient kappa_c,
// no output node) fed to the ordinary merge_bwd / transition_bwd kernels; the -ell(r1) - ell(r2) terms adjust the
// node coefficients of the roots.
// ---------------------------------------------------------------------------------------------
struct VirtArgs {
  double thresh;
  int r, n, N, M;
  int64_t K;
  double grad;
  int dense;               // 1: keep combinations whose adjoint is exactly zero too
  const double* lw;
  const double* stats;
  const double* pot;       // [K][combos] log-softmax
  const int32_t* choice;
  const int32_t* index;    // [K*combos] exclusive scan of the keep flags
  const int32_t* rows;     // [K][N] inherited forest (node ids)
  const int32_t* slot_of;  // compact slot of a consumed node (or null: direct)
  const double* u_bl;
  const double* u_br;
  uint64_t seed;
  const double* lam_l;
  const double* lam_r;
  int64_t v0, v1;          // window of virtual events generated by this launch
  int32_t* v_lsrc;
  int32_t* v_rsrc;
  double* v_coef;
  double* v_t2;
};

__device__ __forceinline__ double kappa_of(const double* lw, const double* stats, const double* pot, const int32_t* choice,
                                           int r, int64_t k, int combos, int c, double grad, double thresh) {
  const double W = exp(lw[k] - stats[r * 4]) * grad;
  const double kappa = W * (exp(pot[k * combos + c]) - (c == choice[k] ? 1.0 : 0.0));
  return fabs(kappa) <= thresh ? 0.0 : kappa;   // thresh = skip_below * |dELBO| (0: exact zeros only)
}

// keep[k*combos + c] = 1 when the virtual event has to be visited (exact-zero adjoints are dropped unless dense)
__global__ void nested_keep_kernel(int r, int n, int M, int64_t K, double grad, int dense, double thresh, const double* __restrict__ lw,
                                   const double* __restrict__ stats, const double* __restrict__ pot,
                                   const int32_t* __restrict__ choice, int32_t* __restrict__ keep) {
  const int combos = n * (n - 1) / 2 * M;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * combos) return;
  const int64_t k = i / combos;
  const int c = (int)(i - k * combos);
  keep[i] = dense ? 1 : (kappa_of(lw, stats, pot, choice, r, k, combos, c, grad, thresh) != 0.0);
}

__global__ void nested_virtual_kernel(const VirtArgs a) {
  const int combos = a.n * (a.n - 1) / 2 * a.M;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.K * combos) return;
  const int64_t k = i / combos;
  const int c = (int)(i - k * combos);
  const double kappa = kappa_of(a.lw, a.stats, a.pot, a.choice, a.r, k, combos, c, a.grad, a.dense ? 0.0 : a.thresh);
  if (!a.dense && kappa == 0.0) return;
  const int64_t v = a.index[i];
  if (v < a.v0 || v >= a.v1) return;
  const int64_t o = v - a.v0;
  int r1, r2;
  pair_of(c / a.M, a.n, r1, r2);
  const int id1 = a.rows[k * a.N + r1], id2 = a.rows[k * a.N + r2];
  a.v_lsrc[o] = id1 < a.N ? -(id1 + 1) : (a.slot_of ? a.slot_of[id1 - a.N] : id1 - a.N);
  a.v_rsrc[o] = id2 < a.N ? -(id2 + 1) : (a.slot_of ? a.slot_of[id2 - a.N] : id2 - a.N);
  a.v_coef[o] = kappa;
  double ul, ur;
  look_uniforms(a.u_bl, a.u_br, a.seed, a.r, k, a.K, a.M, c, ul, ur);
  a.v_t2[2 * o] = -log(ul) / a.lam_l[a.r];
  a.v_t2[2 * o + 1] = -log(ur) / a.lam_r[a.r];
}

// root coefficient adjustments: d/d ell(root p) = -sum_{c containing p} kappa_c, scattered like the D table
__global__ void nested_coef_kernel(int r, int n, int N, int M, int64_t K, double grad, const double* __restrict__ lw,
                                   const double* __restrict__ stats, const double* __restrict__ pot,
                                   const int32_t* __restrict__ choice, const int32_t* __restrict__ anc,
                                   double* __restrict__ Dacc_next) {
  const int combos = n * (n - 1) / 2 * M;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t k = i / combos;
  const int c = (int)(i - k * combos);
  if (k >= K) return;
  const double W = exp(lw[k] - stats[r * 4]) * grad;
  if (W == 0.0) return;
  const double kappa = W * (exp(pot[k * combos + c]) - (c == choice[k] ? 1.0 : 0.0));
  if (kappa == 0.0) return;
  int r1, r2;
  pair_of(c / M, n, r1, r2);
  const int64_t A = r > 0 ? anc[k] : k;
  atomicAdd(Dacc_next + A * N + r1, -kappa);
  atomicAdd(Dacc_next + A * N + r2, -kappa);
}

__global__ void nested_active_kernel(int r, int64_t K, int skip_zero, double rel, const double* __restrict__ lw,
                                     const double* __restrict__ stats, int32_t* __restrict__ active) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  active[k] = skip_zero ? (exp(lw[k] - stats[r * 4]) > rel) : 1;   // |kappa| <= W: below the threshold nothing is kept
}

// roots of the active particles must be materialised (and own adjoint slots) in the reverse sweep
__global__ void nested_mark_roots_kernel(int n, int N, int64_t K, const int32_t* __restrict__ active,
                                         const int32_t* __restrict__ rows, int32_t* __restrict__ consumed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t k = i / n;
  const int p = (int)(i - k * n);
  if (k >= K || !active[k]) return;
  const int id = rows[k * N + p];
  if (id >= N) consumed[id - N] = 1;
}

// db -> dlam (b = -log U / lam) and dQ of the virtual events, reduced over the batch
__global__ void __launch_bounds__(256) nested_reduce_kernel(int r, int64_t V, int jc, const double* __restrict__ dt,
                                                            const double* __restrict__ dQ_each, const double* __restrict__ t2,
                                                            const double* __restrict__ lam_l, const double* __restrict__ lam_r,
                                                            double* __restrict__ dlam_l, double* __restrict__ dlam_r,
                                                            double* __restrict__ dQ) {
  __shared__ double red[8];
  double dl = 0.0, dr = 0.0, q[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) q[e] = 0.0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
    dl += dt[2 * v] * (-t2[2 * v] / lam_l[r]);
    dr += dt[2 * v + 1] * (-t2[2 * v + 1] / lam_r[r]);
    if (!jc) {
#pragma unroll
      for (int e = 0; e < 16; ++e) q[e] += dQ_each[(2 * v) * 16 + e] + dQ_each[(2 * v + 1) * 16 + e];
    }
  }
  const double tl = block_sum<256>(dl, red);
  const double tr = block_sum<256>(dr, red);
  if (threadIdx.x == 0) {
    atomicAdd(dlam_l + r, tl);
    atomicAdd(dlam_r + r, tr);
  }
  if (!jc) {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const double t = block_sum<256>(q[e], red);
      if (threadIdx.x == 0) atomicAdd(dQ + e, t);
    }
  }
}

}  // namespace

int nested_max_roots() { return kMaxNestedRoots; }

int launch_nested_inherit(int r, int n, int N, int64_t K, const double* cdf, const double* u_res, const int32_t* ids_prev,
                          const int32_t* cnt_prev, const int32_t* slot_prev, int32_t* ids, int32_t* cnt, int32_t* slot,
                          int32_t* rows_all, const double* LL_prev, int32_t* anc, double* ll_tilde, cudaStream_t st) {
  InheritArgs a{r, n, N, K, cdf, u_res, ids_prev, cnt_prev, slot_prev, ids, cnt, slot, rows_all, LL_prev, anc, ll_tilde};
  nested_inherit_kernel<<<(unsigned)((K * n + 255) / 256), 256, 0, st>>>(a);
  VCSMC_LAUNCH_CHECK("nested_inherit_kernel");
  return VCSMC_OK;
}

int launch_lookahead(int r, int n, int N, int M, int jc, int gc, int S, int64_t K, const int32_t* ids, const int32_t* cnt,
                     const int32_t* slot, const uint8_t* codes, int64_t codes_stride, const double* pool, int64_t slot_sites,
                     const double* ell_node, const double* ldf, const double* Q, const double* pi, const double* lam_l,
                     const double* lam_r, const double* u_bl, const double* u_br, uint64_t seed, double* pot,
                     double share, cudaStream_t st) {
  if (n > kMaxNestedRoots) { set_error("nested look-ahead supports at most %d live subtrees (got %d)", kMaxNestedRoots, n); return VCSMC_ERR_ARG; }
  int tile = kLookSmem / (n * 32);
  if (tile > kLookTile) tile = kLookTile;
  if (tile < 8) tile = 8;
  LookArgs a{r, n, N, M, jc, gc, S, tile, K, ids, cnt, slot, codes, codes_stride, pool, slot_sites, ell_node, ldf, Q, pi,
             lam_l, lam_r, u_bl, u_br, seed, pot, share};
  size_t smem = (size_t)n * tile * 32;
  if (smem < 256 * 8 * 2) smem = 256 * 8 * 2;   // (the buffer doubles as the [G][combos] scratch of the site-slice sum)
  static bool configured = false;
  if (!configured) {
    VCSMC_CUDA(cudaFuncSetAttribute(lookahead_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxRoots * 8 * 32 > kLookSmem ? kMaxRoots * 8 * 32 : kLookSmem));
    VCSMC_CUDA(cudaFuncSetAttribute(lookahead_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxRoots * 8 * 32 > kLookSmem ? kMaxRoots * 8 * 32 : kLookSmem));
    configured = true;
  }
  if (jc) lookahead_kernel<true><<<(unsigned)K, kLookThreads, smem, st>>>(a);
  else lookahead_kernel<false><<<(unsigned)K, kLookThreads, smem, st>>>(a);
  VCSMC_LAUNCH_CHECK("lookahead_kernel");
  return VCSMC_OK;
}

int launch_nested_choose(int r, int n, int N, int M, int gc, int64_t K, double* pot, const double* u_cat, const double* u_bl,
                         const double* u_br, uint64_t seed, const double* lam_l, const double* lam_r, const int32_t* ids,
                         const int32_t* cnt, const int32_t* slot, int32_t* ids_new, int32_t* cnt_new, int32_t* slot_new,
                         int32_t* lref, int32_t* rref, int32_t* nleaf, uint8_t* rempos, int32_t* choice, double* b_l,
                         double* b_r, double* t2, double* qlog, int32_t* lsrc, int32_t* rsrc, int32_t* dst, cudaStream_t st) {
  ChooseArgs a{r, n, N, M, gc, K, pot, u_cat, u_bl, u_br, seed, lam_l, lam_r, ids, cnt, slot, ids_new, cnt_new, slot_new,
               lref, rref, nleaf, rempos, choice, b_l, b_r, t2, qlog, lsrc, rsrc, dst};
  nested_choose_kernel<<<(unsigned)((K + 7) / 8), 256, 0, st>>>(a);
  VCSMC_LAUNCH_CHECK("nested_choose_kernel");
  return VCSMC_OK;
}

int launch_nested_active(int r, int64_t K, int skip_zero, double rel, const double* lw, const double* stats, int32_t* active, cudaStream_t st) {
  nested_active_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(r, K, skip_zero, rel, lw, stats, active);
  VCSMC_LAUNCH_CHECK("nested_active_kernel");
  return VCSMC_OK;
}

int launch_nested_mark_roots(int n, int N, int64_t K, const int32_t* active, const int32_t* rows, int32_t* consumed, cudaStream_t st) {
  nested_mark_roots_kernel<<<(unsigned)((K * n + 255) / 256), 256, 0, st>>>(n, N, K, active, rows, consumed);
  VCSMC_LAUNCH_CHECK("nested_mark_roots_kernel");
  return VCSMC_OK;
}

int launch_nested_coef(int r, int n, int N, int M, int64_t K, double grad, const double* lw, const double* stats,
                       const double* pot, const int32_t* choice, const int32_t* anc, double* Dacc_next, cudaStream_t st) {
  const int64_t total = K * (int64_t)(n * (n - 1) / 2 * M);
  nested_coef_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(r, n, N, M, K, grad, lw, stats, pot, choice, anc, Dacc_next);
  VCSMC_LAUNCH_CHECK("nested_coef_kernel");
  return VCSMC_OK;
}

int launch_nested_keep(int r, int n, int M, int64_t K, double grad, int dense, double thresh, const double* lw, const double* stats,
                       const double* pot, const int32_t* choice, int32_t* keep, cudaStream_t st) {
  const int64_t total = K * (int64_t)(n * (n - 1) / 2 * M);
  nested_keep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(r, n, M, K, grad, dense, thresh, lw, stats, pot, choice, keep);
  VCSMC_LAUNCH_CHECK("nested_keep_kernel");
  return VCSMC_OK;
}

int launch_nested_virtual(int r, int n, int N, int M, int64_t K, double grad, int dense, double thresh, const double* lw, const double* stats,
                          const double* pot, const int32_t* choice, const int32_t* index, const int32_t* rows,
                          const int32_t* slot_of, const double* u_bl, const double* u_br, uint64_t seed, const double* lam_l,
                          const double* lam_r, int64_t v0, int64_t v1, int32_t* v_lsrc, int32_t* v_rsrc, double* v_coef,
                          double* v_t2, cudaStream_t st) {
  VirtArgs a{thresh, r, n, N, M, K, grad, dense, lw, stats, pot, choice, index, rows, slot_of, u_bl, u_br, seed, lam_l, lam_r,
             v0, v1, v_lsrc, v_rsrc, v_coef, v_t2};
  const int64_t total = K * (int64_t)(n * (n - 1) / 2 * M);
  nested_virtual_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
  VCSMC_LAUNCH_CHECK("nested_virtual_kernel");
  return VCSMC_OK;
}

int launch_nested_reduce(int r, int64_t V, int jc, const double* dt, const double* dQ_each, const double* t2,
                         const double* lam_l, const double* lam_r, double* dlam_l, double* dlam_r, double* dQ,
                         cudaStream_t st) {
  if (V <= 0) return VCSMC_OK;
  int64_t blocks = (V + 255) / 256;
  if (blocks > 592) blocks = 592;
  nested_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(r, V, jc, dt, dQ_each, t2, lam_l, lam_r, dlam_l, dlam_r, dQ);
  VCSMC_LAUNCH_CHECK("nested_reduce_kernel");
  return VCSMC_OK;
}

}  // namespace vcsmc
