// The SMC sweep: sample_phylogenies / body_rank_update (vcsmc.py:332-451) forward, and the reverse sweep that
// TF autodiff of cost = -elbo performs (vcsmc.py:488-491) backward.
//
// B200-first data model (not the reference's): a subtree's partial-likelihood vector never changes once
// computed, so it is written ONCE into a node pool; a particle is a row of int32 node references plus cached
// per-node scalars.  The reference's three gathers + concat + resample gather per rank event (vcsmc.py:286,
// :361-368), which copy the whole [K,n,S,4] state, become int-table gathers; compute_forest_posterior
// (vcsmc.py:231-245) re-reads nothing: only the NEW node's sum_s log(pi.L) is computed (fused in the merge).
//
// Memory plan (H2): if every node of the sweep fits, nodes are direct-mapped (slot = r*K + k) and kept for
// the backward pass.  Otherwise the forward runs on a garbage-collected slot pool (dead nodes are recycled
// after every resampling) and the backward recomputes the forward per SITE CHUNK (sites are independent
// once ancestors, pairs, branch lengths and the per-node coefficients are known from the scalar tables).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include <cooperative_groups.h>

#include "sweep_state.h"

namespace vcsmc {
namespace {

// ---------------------------------------------------------------------------------------------
// forward kernels
// ---------------------------------------------------------------------------------------------
__global__ void leaf_ell_kernel(const uint8_t* __restrict__ codes, int64_t stride, int S, const double* __restrict__ pi,
                                double* __restrict__ ell_node) {
  // ell_leaf = sum_s log(pi . leaf[s])   (the leaves' share of compute_forest_posterior, vcsmc.py:238-242)
  __shared__ double red[8];
  const int leaf = blockIdx.x;
  double p[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) p[j] = pi[j];
  double acc = 0.0;
  for (int s = threadIdx.x; s < S; s += 256) {
    const d4 L = leaf_site(codes[(int64_t)leaf * stride + s]);
    double x = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) x = fma(p[j], L.v[j], x);
    acc += log(x);
  }
  const double t = block_sum<256>(acc, red);
  if (threadIdx.x == 0) ell_node[leaf] = t;
}

struct PrepArgs {
  int r, n, N, gc;
  int64_t K;
  const double* cdf;
  const double* u_res;
  const float* u_pair;
  const double* u_bl;
  const double* u_br;
  const double* lam_l;
  const double* lam_r;
  const int32_t* ids_old;
  const int32_t* cnt_old;
  const int32_t* slot_old;
  int32_t* ids_new;
  int32_t* cnt_new;
  int32_t* slot_new;
  const double* LL_prev;
  int32_t* anc;
  int32_t* lref;
  int32_t* rref;
  int32_t* nleaf;
  uint8_t* rempos;
  double* b_l;
  double* b_r;
  double* t2;
  double* ll_tilde;
  int32_t* lsrc;
  int32_t* rsrc;
  int32_t* dst;
};

constexpr int kPrepWarps = 8;

// One warp per particle: resampling draw, pair proposal, forest-row update, branch lengths.
// (resample vcsmc.py:284-289,318-325; extend_partial_state :298-305; Exponential sample :351-358; state update :361-373)
__global__ void __launch_bounds__(kPrepWarps * 32) step_prepare_kernel(const PrepArgs a) {
  extern __shared__ float su_all[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * kPrepWarps + wid;
  if (k >= a.K) return;
  const int n = a.n, N = a.N, r = a.r;
  float* su = su_all + wid * n;
  for (int i = lane; i < n; i += 32) su[i] = a.u_pair[k * n + i];

  int idx = (int)k;
  if (r > 0) {
    if (lane == 0) idx = upper_bound_cdf(a.cdf, a.K, a.u_res[k] * a.cdf[a.K - 1]);
    idx = __shfl_sync(0xffffffffu, idx, 0);
  }
  __syncwarp();
  const int32_t* io = a.ids_old + (int64_t)idx * N;
  const int32_t* co = a.cnt_old + (int64_t)idx * N;
  const int32_t* so = a.slot_old + (int64_t)idx * N;
  int32_t* in_ = a.ids_new + k * N;
  int32_t* cn = a.cnt_new + k * N;
  int32_t* sn = a.slot_new + k * N;
  uint8_t* rp = a.rempos + k * (int64_t)(n - 2);
  const bool first = (r == 0);  // initial forest = the N leaves, one each (vcsmc.py:414-415)
  const int gc = a.gc;
  int c0, c1;
  rank_pairs_warp(su, n, lane, c0, c1, [&](int pos, int i) {
    rp[pos] = (uint8_t)i;
    in_[pos] = first ? i : io[i];
    cn[pos] = first ? 1 : co[i];
    if (gc) sn[pos] = first ? -1 : so[i];
  });
  if (lane == 0) {
    const int lid = first ? c0 : io[c0], rid = first ? c1 : io[c1];
    const int64_t e = (int64_t)r * a.K + k;
    in_[n - 2] = (int32_t)(N + e);
    const int nl = (first ? 1 : co[c0]) + (first ? 1 : co[c1]);
    cn[n - 2] = nl;
    a.nleaf[k] = nl;
    a.anc[k] = idx;
    a.lref[k] = lid;
    a.rref[k] = rid;
    a.lsrc[k] = lid < N ? -(lid + 1) : (gc ? so[c0] : lid - N);
    a.rsrc[k] = rid < N ? -(rid + 1) : (gc ? so[c1] : rid - N);
    if (!gc) a.dst[k] = (int32_t)e;  // direct map: slot = r*K + k (GC mode: gc_alloc_kernel assigns it)
    const double bl = -log(a.u_bl[k]) / a.lam_l[r];
    const double br = -log(a.u_br[k]) / a.lam_r[r];
    a.b_l[k] = bl;
    a.b_r[k] = br;
    a.t2[2 * k] = bl;
    a.t2[2 * k + 1] = br;
    a.ll_tilde[k] = first ? log(1.0 / (double)a.K) : a.LL_prev[idx];
  }
}

// --- slot pool garbage collection (forward, GC mode) ---
__global__ void gc_mark_kernel(const int32_t* __restrict__ ids_new, const int32_t* __restrict__ slot_new, int N, int n,
                               int64_t K, const int32_t* __restrict__ lsrc, const int32_t* __restrict__ rsrc,
                               int32_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t k = i / n;
  const int p = (int)(i - k * n);
  if (k >= K) return;
  if (p < n - 2) {
    if (ids_new[k * N + p] >= N) flags[slot_new[k * N + p]] = 1;
  } else if (p == n - 2) {
    if (lsrc[k] >= 0) flags[lsrc[k]] = 1;
  } else {
    if (rsrc[k] >= 0) flags[rsrc[k]] = 1;
  }
}

// The K lowest free slots, in order, go to particles 0..K-1.  Single CTA, fixed order.  Each warp owns a contiguous
// segment of the flag array and walks it 32 flags at a time (coalesced), compacting with ballot/popc.
__global__ void __launch_bounds__(1024) gc_alloc_kernel(const int32_t* __restrict__ flags, int64_t P, int64_t K, int N, int n,
                                                        int32_t* __restrict__ dst, int32_t* __restrict__ slot_new,
                                                        int32_t* __restrict__ status) {
  __shared__ int64_t warp_off[33];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // Lowest-free-first allocation keeps every slot ever used below the running peak (live + K), and the live slots of
  // this event are a subset of last event's (live + new) <= previous peak: the K lowest free slots are < peak + K.
  const int64_t Pfull = P;
  P = min(Pfull, (int64_t)status[1] + K);
  const int64_t seg = ((P + 31) / 32 + 31) / 32 * 32;  // per-warp segment length, multiple of 32
  const int64_t b = min((int64_t)wid * seg, P), e = min(b + seg, P);
  int cnt = 0;
  for (int64_t i0 = b + lane; i0 < e; i0 += 128) {
    int f[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) f[q] = (i0 + 32 * q < e) ? flags[i0 + 32 * q] : 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) cnt += f[q] == 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) warp_off[wid + 1] = cnt;
  __syncthreads();
  if (tid == 0) {
    int64_t t = 0;
    warp_off[0] = 0;
    for (int i = 1; i <= 32; ++i) {
      t += warp_off[i];
      warp_off[i] = t;
    }
    if (t < K) status[0] = VCSMC_ERR_POOL;
    const int64_t used = P - t + (t < K ? t : K);
    if (used > status[1]) status[1] = (int32_t)used;
    for (int64_t j = t; j < K; ++j) {  // pool exhausted: these particles compute ell only
      dst[j] = -1;
      slot_new[j * N + (n - 2)] = -1;
    }
  }
  __syncthreads();
  int64_t j = warp_off[wid];
  for (int64_t i0 = b; i0 < e && j < K; i0 += 128) {
    int f[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) f[q] = (i0 + 32 * q + lane < e) ? flags[i0 + 32 * q + lane] : 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t i = i0 + 32 * q + lane;
      const bool is_free = f[q] == 0;
      const unsigned m = __ballot_sync(0xffffffffu, is_free);
      const int64_t mine = j + __popc(m & ((1u << lane) - 1));
      if (is_free && mine < K) {
        dst[mine] = (int32_t)i;
        slot_new[mine * N + (n - 2)] = (int32_t)i;
      }
      j += __popc(m);
    }
  }
}


// compute_forest_posterior with cached per-node scalars + branch priors + v^- + weight (vcsmc.py:376-395)
__global__ void step_weights_kernel(const WeightArgs a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.K) return;
  const int N = a.N, n = a.n, r = a.r;
  double ell = 0.0;
  for (int t = 0; t < a.tiles; ++t) ell += a.ell_part[k * a.tiles + t];
  a.ell_node[a.e_off + k] = ell;
  double F = 0.0, topo = 0.0;
  int vm = 0;
  for (int p = 0; p < n - 2; ++p) {
    F += a.ell_node[a.ids_new[k * N + p]];
    const int c = a.cnt_new[k * N + p];
    topo -= a.ldf[2 * max(c, 2) - 3];
    vm += c - (c == 1);
  }
  {
    F += ell;
    const int c = a.cnt_new[k * N + n - 2];
    topo -= a.ldf[2 * max(c, 2) - 3];
    vm += c - (c == 1);
  }
  const double laml = a.lam_l[r], lamr = a.lam_r[r];
  const double bl = a.b_l[k], br = a.b_r[k];
  const double cl = (r > 0 ? a.cum_l_prev[k] : 0.0) + bl;   // quirk Q1: slot-wise, un-resampled histories
  const double cr = (r > 0 ? a.cum_r_prev[k] : 0.0) + br;
  a.cum_l[k] = cl;
  a.cum_r[k] = cr;
  const double llog = log(laml), rlog = log(lamr);
  // quirk Q2: the CURRENT step's rate multiplies ALL earlier branches (vcsmc.py:380-383)
  const double LLr = (F + topo) + (-laml * cl + (double)(r + 1) * llog) + (-lamr * cr + (double)(r + 1) * rlog);
  // quirk Q3: q = 1/C(n,2) is subtracted raw (vcsmc.py:298,392)
  const double lw = LLr - a.ll_tilde[k] - (llog - laml * bl + rlog - lamr * br) + log((double)vm) - (a.qlog ? a.qlog[k] : a.q);
  a.LL[k] = LLr;
  a.lw[k] = lw;
  a.vminus[k] = vm;
}

// ELBO (vcsmc.py:276) and log_likelihood_R (vcsmc.py:254-268, incl. quirk Q4)
__global__ void finalize_kernel(int N, int64_t K, const double* __restrict__ stats, const double* __restrict__ LL_last,
                                const double* __restrict__ b_l, const double* __restrict__ b_r,
                                const double* __restrict__ lam_l, const double* __restrict__ lam_r, double ldf_root,
                                double* __restrict__ llR, double* __restrict__ elbo, double* __restrict__ logz,
                                double* __restrict__ ess) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0) {
    double s = 0.0;
    const double lk = log((double)K);
    for (int r = 0; r < N - 1; ++r) {
      const double z = stats[r * 4] - lk;
      logz[r] = z;
      ess[r] = stats[r * 4 + 2];
      s += z;
    }
    elbo[0] = s;
  }
  if (k >= K) return;
  double lp = 0.0, rp = 0.0;
  for (int r = 0; r < N - 1; ++r) {
    const double ll = log(lam_l[r]);
    lp += ll - b_l[(int64_t)r * K + k] * lam_l[r];
    rp += ll - b_r[(int64_t)r * K + k] * lam_r[r];  // log(LEFT param): vcsmc.py:262
  }
  llR[k] = LL_last[k] + ldf_root - lp - rp;
}

// ---------------------------------------------------------------------------------------------
// backward kernels (scalar tables)
// ---------------------------------------------------------------------------------------------
__global__ void mark_consumed_kernel(const int32_t* __restrict__ lref, const int32_t* __restrict__ rref, int64_t n, int N,
                                     int32_t* __restrict__ consumed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (lref[i] >= N) consumed[lref[i] - N] = 1;
  if (rref[i] >= N) consumed[rref[i] - N] = 1;
}

// child / adjoint / output slots of every event for the reverse sweep.  slot_of == null: direct map (slot = event
// index, nodes retained from the forward); otherwise the compact numbering of the consumed nodes (chunked mode).
__global__ void bwd_src_kernel(const int32_t* __restrict__ lref, const int32_t* __restrict__ rref,
                               const int32_t* __restrict__ consumed, const int32_t* __restrict__ slot_of, int64_t n, int N,
                               int32_t* __restrict__ lsrc, int32_t* __restrict__ rsrc, int32_t* __restrict__ gsrc,
                               int32_t* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = lref[i], r = rref[i];
  lsrc[i] = l < N ? -(l + 1) : (slot_of ? slot_of[l - N] : l - N);
  rsrc[i] = r < N ? -(r + 1) : (slot_of ? slot_of[r - N] : r - N);
  const int c = consumed[i];
  const int mine = slot_of ? slot_of[i] : (int32_t)i;
  gsrc[i] = c ? mine : -1;
  dst[i] = c ? mine : -1;
}

struct CoefArgs {
  int r, n, N;
  int64_t K;
  double grad;
  double share;  // fraction of the site-independent gradient terms this rank owns (site sharding)
  const double* lw;
  const double* stats;  // row r: lse
  const int32_t* anc;
  const uint8_t* rempos;
  const double* childsum_cur;
  double* childsum_next;
  const double* Dacc_cur;
  double* Dacc_next;
  double* cnew;
  const double* lam_l;
  const double* lam_r;
  const double* b_l;
  const double* b_r;
  const double* cum_l;
  const double* cum_r;
  double* suf_l;
  double* suf_r;
  double* dlam_l;  // [N-1]
  double* dlam_r;
};

// Adjoint of the weight algebra for rank event r (vcsmc.py:376-395 reversed):
//   W = softmax(lw_r) * grad;  a = dELBO/dLL_r[k] = W - sum_{children k' at r+1} W_{r+1}[k'];
//   D[p] = a + (what the descendants' forests still hold of entry p);  c_new = D[last].
__global__ void __launch_bounds__(256) bwd_coef_kernel(const CoefArgs a) {
  __shared__ double red[8];
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = a.N, n = a.n, r = a.r;
  double dl = 0.0, dr = 0.0;
  if (k < a.K) {
    const double W = exp(a.lw[k] - a.stats[r * 4]) * a.grad;
    const double av = W - a.childsum_cur[k];
    const int64_t A = r > 0 ? a.anc[k] : k;
    const uint8_t* rp = a.rempos + k * (int64_t)(n - 2);
    for (int p = 0; p < n - 2; ++p) {
      const double D = av + a.Dacc_cur[k * N + p];
      if (D != 0.0) atomicAdd(a.Dacc_next + A * N + rp[p], D);
    }
    a.cnew[k] = av + a.Dacc_cur[k * N + n - 2];
    if (r > 0 && W != 0.0) atomicAdd(a.childsum_next + A, W);
    const double laml = a.lam_l[r], lamr = a.lam_r[r];
    const double sl = a.suf_l[k] + av * laml, sr = a.suf_r[k] + av * lamr;  // sum_{r' >= r} a_{r'}[k] lam_{r'}
    a.suf_l[k] = sl;
    a.suf_r[k] = sr;
    // site-independent part of dELBO/d b[r][k] (priors and proposal density), pushed through b = -log(U)/lam at once
    const double gBl = W * laml - sl, gBr = W * lamr - sr;
    dl = a.share * (av * (-a.cum_l[k] + (double)(r + 1) / laml) + W * (a.b_l[k] - 1.0 / laml) + gBl * (-a.b_l[k] / laml));
    dr = a.share * (av * (-a.cum_r[k] + (double)(r + 1) / lamr) + W * (a.b_r[k] - 1.0 / lamr) + gBr * (-a.b_r[k] / lamr));
  }
  const double tl = block_sum<256>(dl, red);
  const double tr = block_sum<256>(dr, red);
  if (threadIdx.x == 0) {
    if (tl != 0.0) atomicAdd(a.dlam_l + r, tl);
    if (tr != 0.0) atomicAdd(a.dlam_r + r, tr);
  }
}

// which particles of rank event r the reverse sweep / the chunk recompute must visit
__global__ void bwd_active_kernel(const double* __restrict__ cnew, const int32_t* __restrict__ consumed, int64_t K,
                                  int skip_zero, double skip_below, int32_t* __restrict__ act_bwd, int32_t* __restrict__ act_rec) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int c = consumed[k];
  act_rec[k] = c;
  act_bwd[k] = c || !(skip_zero && fabs(cnew[k]) <= skip_below);  // one visiting order serves the recompute and the reverse sweep
}

// ---------------------------------------------------------------------------------------------
// The whole scalar pass of the reverse sweep in ONE cooperative launch (VCSMC proposal).  Rank events are strictly
// sequential (the coefficients of event r need what event r+1 scattered), so the grid walks r = N-2 ... 0 with a grid
// barrier between events instead of 5 launches + two 33 MB memsets per event.  Per event a thread does what
// bwd_coef_kernel does for one particle, zeroes the accumulator entries it consumed (they are the scatter targets two
// events later), decides whether the reverse merge has to visit the particle and appends it to the event's list.
// ---------------------------------------------------------------------------------------------
struct ScalarArgs {
  int N, skip_zero;
  int64_t K;
  double grad, share, thresh;
  const double* lw;
  const double* stats;
  const int32_t* anc;
  const uint8_t* rempos;
  double* childsum[2];
  double* Dacc[2];
  int32_t* dirty[2];   // [K]: accumulator row k of that buffer holds non-zeros
  double* cnew;
  const double* lam_l;
  const double* lam_r;
  const double* b_l;
  const double* b_r;
  const double* cum_l;
  const double* cum_r;
  double* suf_l;
  double* suf_r;
  double* dlam_l;
  double* dlam_r;
  const int32_t* consumed;
  int32_t* act_all;    // [N-1][K]
  int32_t* order_bwd;  // [N-1][K]: the particles the reverse merge of event r has to visit (in no particular order)
  int32_t* count_bwd;  // [N-1]
};

__global__ void __launch_bounds__(256) bwd_scalar_kernel(const ScalarArgs a) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  __shared__ double red[8];
  const int N = a.N, lane = threadIdx.x & 31;
  const int64_t K = a.K;
  const int64_t stride = (int64_t)gridDim.x * 256;
  const int64_t Kround = (K + 31) / 32 * 32;
  // byte offset of event r's block of kept positions: sum_{r' < r} align16(K (N - r' - 2))
  int64_t rem_off = 0;
  for (int q = 0; q < N - 2; ++q) rem_off += (K * (int64_t)(N - q - 2) + 15) / 16 * 16;
  for (int r = N - 2; r >= 0; --r) {
    const int n = N - r, cur = r & 1, nxt = cur ^ 1;
    const double laml = a.lam_l[r], lamr = a.lam_r[r], lse = a.stats[r * 4];
    double* Dcur = a.Dacc[cur];
    double* Dnxt = a.Dacc[nxt];
    double dl = 0.0, dr = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < Kround; k += stride) {
      bool act = false, need = false, is_dirty = false;
      const bool valid = k < K;
      const int64_t e = (int64_t)r * K + (valid ? k : 0);
      double W = 0.0, av = 0.0, c = 0.0;
      int A = 0;
      if (valid) {
        W = exp(a.lw[e] - lse) * a.grad;
        av = W - a.childsum[cur][k];
        a.childsum[cur][k] = 0.0;
        A = r > 0 ? a.anc[e] : (int)k;
        is_dirty = a.dirty[cur][k] != 0;   // somebody scattered into this particle's accumulator row two events ago
        if (is_dirty) a.dirty[cur][k] = 0;
        need = is_dirty || av != 0.0;
        c = av;                            // clean row: D = av at every position
      }
      // rows that hold or produce something are walked by the whole warp (coalesced); with ESS ~ 1 that is a handful
      unsigned todo = __ballot_sync(0xffffffffu, need);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t kk = __shfl_sync(0xffffffffu, (int)k, src);
        const double avv = __shfl_sync(0xffffffffu, av, src);
        const int AA = __shfl_sync(0xffffffffu, A, src);
        const bool dd = __shfl_sync(0xffffffffu, (int)is_dirty, src) != 0;
        const uint8_t* rp = a.rempos + rem_off + kk * (int64_t)(n - 2);
        double* row = Dcur + kk * N;
        double last = 0.0;
        bool wrote = false;
        for (int p = lane; p < n - 1; p += 32) {
          double d = 0.0;
          if (dd) {
            d = row[p];
            row[p] = 0.0;
          }
          if (p < n - 2) {
            const double D = avv + d;
            if (D != 0.0) {
              atomicAdd(Dnxt + (int64_t)AA * N + rp[p], D);
              wrote = true;
            }
          } else {
            last = d;
          }
        }
        last = __shfl_sync(0xffffffffu, last, (n - 2) & 31);
        if (lane == src) c = av + last;
        if (__any_sync(0xffffffffu, wrote) && lane == 0) a.dirty[nxt][AA] = 1;
      }
      if (valid) {
        a.cnew[e] = c;
        if (r > 0 && W != 0.0) atomicAdd(a.childsum[nxt] + A, W);
        const double sl = a.suf_l[k] + av * laml, sr = a.suf_r[k] + av * lamr;  // sum_{r' >= r} a_{r'}[k] lam_{r'}
        a.suf_l[k] = sl;
        a.suf_r[k] = sr;
        const double bl = a.b_l[e], br = a.b_r[e];
        const double gBl = W * laml - sl, gBr = W * lamr - sr;
        dl += a.share * (av * (-a.cum_l[e] + (double)(r + 1) / laml) + W * (bl - 1.0 / laml) + gBl * (-bl / laml));
        dr += a.share * (av * (-a.cum_r[e] + (double)(r + 1) / lamr) + W * (br - 1.0 / lamr) + gBr * (-br / lamr));
        act = a.consumed[e] || !(a.skip_zero && fabs(c) <= a.thresh);
        a.act_all[e] = act;
      }
      const unsigned m = __ballot_sync(0xffffffffu, act);
      if (m) {
        int base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(a.count_bwd + r, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (act) a.order_bwd[(int64_t)r * K + base + __popc(m & ((1u << lane) - 1))] = (int32_t)k;
      }
    }
    const double tl = block_sum<256>(dl, red);
    const double tr = block_sum<256>(dr, red);
    if (threadIdx.x == 0) {
      if (tl != 0.0) atomicAdd(a.dlam_l + r, tl);
      if (tr != 0.0) atomicAdd(a.dlam_r + r, tr);
    }
    if (r > 0) rem_off -= (K * (int64_t)(N - (r - 1) - 2) + 15) / 16 * 16;
    grid.sync();
  }
}

// particle sharding: the forward filled P only for this rank's particles; the reverse sweep needs the matrices of the
// particles it visits (any rank's), from the gathered branch lengths -- matrix i of event blockIdx.y is child (i & 1) of
// particle order[r][i >> 1]
__global__ void __launch_bounds__(64) transition_visited_kernel(const double* __restrict__ Q, const double* __restrict__ t2,
                                                                const int32_t* __restrict__ order, const int32_t* __restrict__ count,
                                                                int64_t K, int jc, double* __restrict__ P) {
  __shared__ __align__(16) double s_tab[kExpmTableDoubles];   // the table of m4_expm_tq: the forward's own matrices, bit for bit
  if (!jc) {
    if (threadIdx.x == 0) expm_tq_table(Q, s_tab);
    __syncthreads();
  }
  const int r = blockIdx.y;
  const int64_t n = 2 * (int64_t)count[r];
  for (int64_t i = (int64_t)blockIdx.x * 64 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 64) {
    const int64_t m = 2 * ((int64_t)r * K + order[(int64_t)r * K + (i >> 1)]) + (i & 1);
    const double ti = t2[m];
    double* out = P + m * 16;
    if (jc) {
      const double o = -0.25 * expm1(-ti);
      const double d = 0.25 + 0.75 * exp(-ti);
#pragma unroll
      for (int e = 0; e < 16; ++e) out[e] = (e % 5 == 0) ? d : o;
    } else {
      const M4 X = m4_expm_tq(s_tab, ti);
#pragma unroll
      for (int e = 0; e < 16; ++e) out[e] = X.a[e];
    }
  }
}

// dP rows of the particles the reverse merge is going to visit (instead of clearing the whole [N-1][K][32] table)
__global__ void __launch_bounds__(256) zero_dP_rows_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ count,
                                                           int64_t K, double* __restrict__ dP) {
  const int r = blockIdx.y;
  const int64_t cnt = count[r];
  for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < cnt; i += (int64_t)gridDim.x * 8)
    dP[((int64_t)r * K + order[(int64_t)r * K + i]) * 32 + (threadIdx.x & 31)] = 0.0;
}

// dP -> (db, dQ) -> dlam for the visited particles of EVERY rank event in one launch: matrix i of event blockIdx.y is
// child (i & 1) of particle order[r][i >> 1].  (vcsmc.py:353-358 and the expm of :183-184, reversed.)
__global__ void __launch_bounds__(64) bwd_transition_all_kernel(const double* __restrict__ Q, const double* __restrict__ t2,
                                                                const double* __restrict__ dP, const int32_t* __restrict__ order,
                                                                const int32_t* __restrict__ count, int64_t K, int jc,
                                                                const double* __restrict__ b_l, const double* __restrict__ b_r,
                                                                const double* __restrict__ lam_l, const double* __restrict__ lam_r,
                                                                double* __restrict__ dQ_acc, double* __restrict__ dlam_l,
                                                                double* __restrict__ dlam_r) {
  const int r = blockIdx.y;
  const int64_t n = 2 * (int64_t)count[r];
  for (int64_t i = (int64_t)blockIdx.x * 64 + threadIdx.x; i < (n + 63) / 64 * 64; i += (int64_t)gridDim.x * 64) {
    double dt = 0.0;
    int side = 0;
    int64_t k = 0;
    if (i < n) {
      k = order[(int64_t)r * K + (i >> 1)];
      side = (int)(i & 1);
      const int64_t m = ((int64_t)r * K + k) * 2 + side;
      const double ti = t2[m];
      const double* G = dP + m * 16;
      if (jc) {
        const double e = exp(-ti);
        dt = e * (0.25 * G[1] - 0.75 * G[0]);
      } else {
        double any = 0.0;
#pragma unroll
        for (int e = 0; e < 16; ++e) any += fabs(G[e]);
        if (any != 0.0) {
          M4 At, E, X, Y;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int c = 0; c < 4; ++c) At.a[rr * 4 + c] = __ldg(Q + c * 4 + rr) * ti;
#pragma unroll
          for (int e = 0; e < 16; ++e) E.a[e] = G[e];
          m4_expm_frechet(At, E, X, Y);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            dt = fma(Y.a[e], __ldg(Q + e), dt);
            const double v = ti * Y.a[e];
            if (v != 0.0) atomicAdd(dQ_acc + k * 16 + e, v);
          }
        }
      }
      dt *= side ? (-b_r[(int64_t)r * K + k] / lam_r[r]) : (-b_l[(int64_t)r * K + k] / lam_l[r]);
    }
    // lanes alternate left / right children: reduce the two parities separately
    double vl = side ? 0.0 : dt, vr = side ? dt : 0.0;
    vl = warp_sum(vl);
    vr = warp_sum(vr);
    if ((threadIdx.x & 31) == 0) {
      if (vl != 0.0) atomicAdd(dlam_l + r, vl);
      if (vr != 0.0) atomicAdd(dlam_r + r, vr);
    }
  }
}

// zero the adjoint slots of the consumed nodes of one rank event (the first *count entries of `order`)
__global__ void zero_consumed_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ count, int K,
                                     const int32_t* __restrict__ gsrc, int64_t slot_sites, int n_sites,
                                     double* __restrict__ gpool) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sites) return;
  const int n = count ? *count : K;
  d4 z;
#pragma unroll
  for (int j = 0; j < 4; ++j) z.v[j] = 0.0;
  for (int j = blockIdx.y; j < n; j += gridDim.y) {
    const int g = gsrc[order ? order[j] : j];
    if (g >= 0) st_site(gpool + ((int64_t)g * slot_sites + s) * 4, z);
  }
}

// per-site part of db -> dlam through b = -log(U)/lam, and accumulation of the per-matrix dQ (vcsmc.py:353-358 reversed).
// Visits the `count` particles of `list` (null: all K) -- only particles the reverse merge visited carry a dP.
__global__ void __launch_bounds__(256) bwd_branch_kernel(int r, int64_t count, const int32_t* __restrict__ list, int jc,
                                                         const double* __restrict__ dt, const double* __restrict__ dQ_each,
                                                         const double* __restrict__ b_l, const double* __restrict__ b_r,
                                                         const double* __restrict__ lam_l, const double* __restrict__ lam_r,
                                                         double* __restrict__ dQ_acc, double* __restrict__ dlam_l,
                                                         double* __restrict__ dlam_r) {
  __shared__ double red[8];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double dl = 0.0, dr = 0.0;
  if (i < count) {
    const int64_t k = list ? list[i] : i;
    const int64_t m = list ? 2 * i : 2 * k;   // a listed run keeps its per-matrix results compact
    dl = dt[m] * (-b_l[k] / lam_l[r]);
    dr = dt[m + 1] * (-b_r[k] / lam_r[r]);
    if (!jc) {
#pragma unroll
      for (int e = 0; e < 16; ++e) dQ_acc[k * 16 + e] += dQ_each[m * 16 + e] + dQ_each[(m + 1) * 16 + e];
    }
  }
  const double tl = block_sum<256>(dl, red);
  const double tr = block_sum<256>(dr, red);
  if (threadIdx.x == 0) {
    if (tl != 0.0) atomicAdd(dlam_l + r, tl);
    if (tr != 0.0) atomicAdd(dlam_r + r, tr);
  }
}

// out[c] = sum_k in[k][c] (+ add[c]) in a fixed order: one CTA per column
__global__ void __launch_bounds__(256) column_sum_kernel(const double* __restrict__ in, int64_t K, int C, int ld,
                                                         double* __restrict__ out) {
  __shared__ double red[8];
  const int c = blockIdx.x;
  double s = 0.0;
  for (int64_t k = threadIdx.x; k < K; k += 256) s += in[k * ld + c];
  const double t = block_sum<256>(s, red);
  if (threadIdx.x == 0) out[c] += t;
}

// d/dpi of the leaves' sum_s log(pi . leaf[s]) weighted by c_leaf (column sums of the step-0 scatter)
__global__ void __launch_bounds__(256) leaf_pi_grad_kernel(const uint8_t* __restrict__ codes, int64_t stride, int S,
                                                           const double* __restrict__ pi, const double* __restrict__ cleaf,
                                                           double* __restrict__ dpi) {
  __shared__ double red[8];
  const int leaf = blockIdx.x;
  const double c = cleaf[leaf];
  double p[4], acc[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 4; ++j) p[j] = pi[j];
  for (int s = threadIdx.x; s < S; s += 256) {
    const d4 L = leaf_site(codes[(int64_t)leaf * stride + s]);
    double x = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) x = fma(p[j], L.v[j], x);
    const double inv = c / x;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = fma(inv, L.v[j], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double t = block_sum<256>(acc[j], red);
    if (threadIdx.x == 0) atomicAdd(dpi + j, t);
  }
}

}  // namespace
}  // namespace vcsmc


namespace vcsmc {

// Sorting the particles by child pair only pays when groups of several particles are formed (see pick_group in
// merge.cu: R = K * tiles / 4736); below that the visiting order is the identity and ~9 launches per event are saved.
bool use_sorted_order(int64_t K, int n_sites) {
  const int64_t tiles = (n_sites + 511) / 512;
  return K * tiles >= 2 * 148 * 32;
}

double log_double_factorial_host(int m) {  // vcsmc.py:30-57
  double r = 0.0;
  for (int j = m; j >= 2; j -= 2) r += log((double)j);
  return r;
}


int group_particles(vcsmc_sweep* h, const int32_t* lsrc, const int32_t* rsrc, const int32_t* active, int64_t K,
                    int32_t* order_out, int32_t* count_out, cudaStream_t st, int skip_leaf_pairs) {
  return launch_group_order(lsrc, rsrc, active, skip_leaf_pairs, K, h->p<unsigned long long>(h->o_gtab), h->p<int32_t>(h->o_gcnt), h->p<int32_t>(h->o_goff),
                            h->p<int32_t>(h->o_gslot), h->p<int32_t>(h->o_grank), order_out, count_out, h->p<char>(h->o_sort_temp),
                            h->sort_temp, st);
}

int launch_leaf_ell(const uint8_t* codes, int64_t stride, int N, int S, const double* pi, double* ell_node, cudaStream_t st) {
  leaf_ell_kernel<<<N, 256, 0, st>>>(codes, stride, S, pi, ell_node);
  VCSMC_LAUNCH_CHECK("leaf_ell_kernel");
  return VCSMC_OK;
}

int launch_step_weights(const WeightArgs& w, cudaStream_t st) {
  if (w.K <= 0) return VCSMC_OK;
  step_weights_kernel<<<(unsigned)((w.K + 127) / 128), 128, 0, st>>>(w);
  VCSMC_LAUNCH_CHECK("step_weights_kernel");
  return VCSMC_OK;
}

int launch_finalize(int N, int64_t K, const double* stats, const double* LL_last, const double* b_l, const double* b_r,
                    const double* lam_l, const double* lam_r, double ldf_root, double* llR, double* elbo, double* logz,
                    double* ess, cudaStream_t st) {
  finalize_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(N, K, stats, LL_last, b_l, b_r, lam_l, lam_r, ldf_root, llR, elbo, logz, ess);
  VCSMC_LAUNCH_CHECK("finalize_kernel");
  return VCSMC_OK;
}

}  // namespace vcsmc

// ---------------------------------------------------------------------------------------------
// host object
// ---------------------------------------------------------------------------------------------
using namespace vcsmc;

namespace {

int64_t full_pool_bytes(const vcsmc_sweep_config& c) {
  return (int64_t)(c.n_taxa - 1) * c.n_particles * c.n_sites * 32;
}

// Carves the workspace.  pool_bytes < 0: only measure the table bytes.
int64_t plan(vcsmc_sweep* h) {
  const int N = h->N;
  const int64_t K = h->K;
  const int64_t E = (int64_t)(N - 1) * K;
  Layout L;
  h->o_status = L.take<int32_t>(8);
  h->o_sig = L.take<int32_t>(64);
  h->o_epoch_dev = L.take<int32_t>(4);
  h->o_seed_dev = L.take<uint64_t>(2);
  h->o_model = L.take<double>(2 * (int64_t)N + 24 + kExpmTableDoubles);   // lam_l, lam_r, Q, pi; the table of m4_expm_tq
  h->o_elbo = L.take<double>(1);
  h->o_anc = L.take<int32_t>(E);
  h->o_lref = L.take<int32_t>(E);
  h->o_rref = L.take<int32_t>(E);
  h->o_nleaf = L.take<int32_t>(E);
  h->rem_off.assign(N - 1, 0);
  int64_t rem_total = 0;
  for (int r = 0; r < N - 1; ++r) {
    h->rem_off[r] = rem_total;
    rem_total += align_up(K * (int64_t)(N - r - 2), 16);
  }
  h->o_rempos = L.take<uint8_t>(rem_total + 16);
  h->o_b_l = L.take<double>(E);
  h->o_b_r = L.take<double>(E);
  h->o_t2 = L.take<double>(2 * E);
  h->o_cum_l = L.take<double>(E);
  h->o_cum_r = L.take<double>(E);
  h->o_lw = L.take<double>(E);
  h->o_LL = L.take<double>(E);
  h->o_lltilde = L.take<double>(K);
  h->o_llR = L.take<double>(K);
  h->o_vminus = L.take<int32_t>(K);
  h->o_ell_node = L.take<double>(N + E);
  h->o_stats = L.take<double>(4 * (int64_t)N);
  h->o_logz = L.take<double>(N);
  h->o_ess = L.take<double>(N);
  h->o_P = L.take<double>(32 * E);
  for (int i = 0; i < 2; ++i) {
    h->o_ids[i] = L.take<int32_t>(K * N);
    h->o_cnt[i] = L.take<int32_t>(K * N);
    h->o_slot[i] = L.take<int32_t>(K * N);
  }
  h->o_cdf = L.take<double>(K);
  h->o_cdf_scratch = L.take<double>(resample_scratch_doubles(K));
  h->o_u_pair = L.take<float>(K * N);
  h->o_u_bl = L.take<double>(K);
  h->o_u_br = L.take<double>(K);
  h->o_u_res = L.take<double>(K);
  h->tiles_max = merge_ell_parts(h->S);
  h->o_ell_part = L.take<double>(K * h->tiles_max);
  h->o_ell_new = L.take<double>(K);
  h->o_lsrc = L.take<int32_t>(K);
  h->o_rsrc = L.take<int32_t>(K);
  h->o_dst = L.take<int32_t>(K);
  h->o_ldf = L.take<double>(2 * (int64_t)N + 4);
  if (h->M > 0) {
    h->o_inh_ids = L.take<int32_t>(K * N);
    h->o_inh_cnt = L.take<int32_t>(K * N);
    h->o_inh_slot = L.take<int32_t>(K * N);
    h->pot_off.assign(N, 0);
    int64_t tot = 0;
    for (int r = 0; r < N - 1; ++r) {
      h->pot_off[r] = tot;
      tot += K * (int64_t)((N - r) * (N - r - 1) / 2) * h->M;
    }
    h->o_pot = L.take<double>(tot + 1);
    h->o_choice = L.take<int32_t>(E);
    h->o_qlog = L.take<double>(K);
    h->o_u_cat = L.take<double>(K);
  }
  if (h->M == 0) {
    // lazy forward / particle sharding (lazy.cu): node -> local slot map, survivor and fetch lists, inherited rows,
    // the per-event record that is all-gathered across ranks
    h->o_loc = L.take<int32_t>(E);
    h->o_pend = L.take<int32_t>(E);
    h->o_surv = L.take<int32_t>(K);
    h->o_mat_list = L.take<int32_t>(K);
    h->fetch_cap = K * (int64_t)N < E ? K * (int64_t)N : E;
    h->o_fetch_e = L.take<int32_t>(h->fetch_cap);
    h->o_fetch_src = L.take<int32_t>(h->fetch_cap);
    h->o_counts = L.take<int32_t>(8);
    h->o_pF = L.take<double>(K);      // per-particle work arrays of the event kernel (lazy.cu)
    h->o_pT = L.take<double>(K);
    h->o_pV = L.take<int32_t>(K);
    h->o_pLLt = L.take<double>(K);
    h->o_pEll = L.take<double>(K);
    h->o_mat_ls = L.take<int32_t>(K);
    h->o_mat_rs = L.take<int32_t>(K);
    h->o_pDirect = L.take<int32_t>(K);
    h->o_lsrc2 = L.take<int32_t>(K);
    h->o_rsrc2 = L.take<int32_t>(K);
    h->o_F0 = L.take<double>(2);
    h->o_live = L.take<int32_t>(K);
    h->o_haskid = L.take<int32_t>(K);
    h->o_gocc = L.take<int32_t>(K);
    h->o_ev_timing = L.take<unsigned long long>(16 * (int64_t)N);
    h->o_leaf_perm = L.take<int32_t>((int64_t)N * leaf_sort_stride(h->S));
    h->o_leaf_tstate = L.take<uint8_t>((int64_t)N * (leaf_sort_stride(h->S) / 128) + 16);
    h->o_leaf_hist = L.take<int32_t>(leaf_pair_hist_ints(N));
    for (int i = 0; i < 2; ++i) {   // forest scalars that travel with a particle (lazy.cu)
      h->o_F[i] = L.take<double>(K);
      h->o_topo[i] = L.take<double>(K);
      h->o_vm[i] = L.take<int32_t>(K);
    }
    h->rec_stride = align_up((h->Kl > 0 ? h->Kl : K) * (int64_t)(72 + N), 16);
    h->o_rec = L.take<char>(2 * (K * (int64_t)(72 + N) + 16 * kMaxPeers + 256));   // two buffers (by launch parity)
  }
  h->o_keys_in = L.take<uint64_t>(K);
  h->o_keys_out = L.take<uint64_t>(K);
  h->o_vals_in = L.take<int32_t>(K);
  h->o_order = L.take<int32_t>(K);
  h->o_count = L.take<int32_t>(4);
  {
    const int64_t T = group_table_entries(K);
    h->o_gtab = L.take<unsigned long long>(T + T / 2);   // hash keys [T] u64 immediately followed by the counts [T] i32
    h->o_gcnt = h->o_gtab + T * (int64_t)sizeof(unsigned long long);
    h->o_goff = L.take<int32_t>(T);
    h->o_gslot = L.take<int32_t>(K);
    h->o_grank = L.take<int32_t>(K);
  }
  h->sort_temp = sort_temp_bytes(K);
  { const size_t sc = scan_temp_bytes(group_table_entries(K)); if (sc > h->sort_temp) h->sort_temp = sc; }
  if (h->keep) { const size_t sc = scan_temp_bytes(E); if (sc > h->sort_temp) h->sort_temp = sc; }
  h->o_sort_temp = L.take<char>((int64_t)h->sort_temp + 256);
  if (h->keep) {
    for (int i = 0; i < 2; ++i) {
      h->o_childsum[i] = L.take<double>(K);
      h->o_Dacc[i] = L.take<double>(K * N);
    }
    h->o_cnew = L.take<double>(E);
    h->o_consumed = L.take<int32_t>(E);
    h->o_bsrc_l = L.take<int32_t>(E);
    h->o_bsrc_r = L.take<int32_t>(E);
    h->o_bsrc_g = L.take<int32_t>(E);
    h->o_bdst = L.take<int32_t>(E);
    h->o_dP = L.take<double>(32 * E);
    h->o_dpi_each = L.take<double>(4 * K);
    h->o_dQ_acc = L.take<double>(16 * K);
    h->o_dQ_each = L.take<double>(32 * K);
    h->o_dt = L.take<double>(2 * K);
    h->o_suf_l = L.take<double>(K);
    h->o_suf_r = L.take<double>(K);
    h->o_cleaf = L.take<double>(N);
    h->o_order_bwd = L.take<int32_t>(E);
    h->o_order_rec = L.take<int32_t>(E);
    h->o_count_bwd = L.take<int32_t>(N);
    h->o_count_rec = L.take<int32_t>(N);
    h->o_act_all = L.take<int32_t>(h->M == 0 ? E : 1);
    h->o_dirty = L.take<int32_t>(2 * K);
    h->o_act_bwd = L.take<int32_t>(K);
    h->o_act_rec = L.take<int32_t>(K);
    h->o_cslot = L.take<int32_t>(E + 1);
    if (h->M > 0) {
      h->o_rows_all = L.take<int32_t>(E * N);
      h->o_nact = L.take<int32_t>(K);
      h->o_nbase = L.take<int32_t>(K + 1);
      const int64_t all = K * (int64_t)(N * (N - 1) / 2) * h->M;
      h->v_batch = all < (1 << 20) ? all : (1 << 20);
      const int64_t V = h->v_batch;
      h->o_v_lsrc = L.take<int32_t>(V);
      h->o_v_rsrc = L.take<int32_t>(V);
      h->o_v_coef = L.take<double>(V);
      h->o_v_t2 = L.take<double>(2 * V);
      h->o_v_P = L.take<double>(32 * V);
      h->o_v_dP = L.take<double>(32 * V);
      h->o_v_dt = L.take<double>(2 * V);
      h->o_v_dQ = L.take<double>(32 * V);
      h->o_v_dpi = L.take<double>(4 * V);
      h->o_v_order = L.take<int32_t>(V);
      h->o_v_keys_in = L.take<uint64_t>(V);
      h->o_v_keys_out = L.take<uint64_t>(V);
      h->o_v_vals = L.take<int32_t>(V);
      h->o_v_count = L.take<int32_t>(4);
      h->v_temp = sort_temp_bytes(V);
      h->o_v_temp = L.take<char>((int64_t)h->v_temp + 256);
      h->o_v_keep = L.take<int32_t>(all + 1);
      h->o_v_index = L.take<int32_t>(all + 1);
      h->v_scan = scan_temp_bytes(all);
      h->o_v_scan = L.take<char>((int64_t)h->v_scan + 256);
    }
  }
  return L.off;
}

int decide_modes(vcsmc_sweep* h, int64_t tables, int64_t ws_bytes, bool report) {
  const int N = h->N, S = h->S;
  const int64_t K = h->K;
  const int64_t node_bytes = (int64_t)S * 32;
  const int64_t full = (int64_t)(N - 1) * K * node_bytes;
  const int64_t avail = ws_bytes - tables;
  const int64_t need_retain = h->keep ? 2 * full : full;
  if (avail >= need_retain && h->world == 1 && !h->force_gc) {
    h->fwd_gc = false;
    h->retain = true;
    h->chunk_sites = S;
    h->pool_slots = (int64_t)(N - 1) * K;
    h->o_flags = tables;  // unused
    h->o_pool = tables;
    h->pool_bytes = need_retain;
    return VCSMC_OK;
  }
  // GC forward: flags[P] + slot_id[P] + P slots
  h->fwd_gc = true;
  h->retain = false;
  int64_t P = (avail - 8192) / (node_bytes + 8);
  const int64_t Pmax = (int64_t)(N - 1) * K;
  if (P > Pmax) P = Pmax;
  const int64_t P_need = 2 * K < Pmax ? 2 * K : Pmax;   // (a one-event sweep never holds more than its K nodes)
  if (P < P_need) {
    if (report) set_error("workspace too small: GC pool would hold %lld slots, need >= %lld", (long long)P, (long long)P_need);
    return VCSMC_ERR_ARG;
  }
  h->pool_slots = P;
  h->o_flags = tables;
  h->o_slot_id = align_up(tables + P * 4);
  h->o_pool = align_up(h->o_slot_id + P * 4);
  h->pool_bytes = ws_bytes - h->o_pool;
  if (h->keep) {
    int64_t Sc = (h->pool_bytes / 2) / ((int64_t)(N - 1) * K * 32);
    if (Sc >= S) Sc = S;
    else Sc = Sc / 256 * 256;
    if (Sc < 256 && Sc < S) {
      if (report) set_error("workspace too small for a 256-site backward chunk");
      return VCSMC_ERR_ARG;
    }
    h->chunk_sites = (int)Sc;
  } else {
    h->chunk_sites = 0;
  }
  return VCSMC_OK;
}

int check_cfg(const vcsmc_sweep_config* c) {
  if (!c) { set_error("null config"); return VCSMC_ERR_ARG; }
  if (c->n_taxa < 2 || c->n_taxa > kMaxRoots) { set_error("n_taxa=%d out of range [2,%d]", c->n_taxa, kMaxRoots); return VCSMC_ERR_ARG; }
  if (c->n_sites < 1) { set_error("n_sites must be >= 1"); return VCSMC_ERR_ARG; }
  if (c->n_particles < 1) { set_error("n_particles must be >= 1"); return VCSMC_ERR_ARG; }
  if (c->n_sub < 0) { set_error("n_sub must be >= 0"); return VCSMC_ERR_ARG; }
  if (c->n_sub > 0 && c->n_taxa > nested_max_roots()) { set_error("nested look-ahead supports at most %d taxa (got %d)", nested_max_roots(), c->n_taxa); return VCSMC_ERR_ARG; }
  if ((int64_t)(c->n_taxa - 1) * c->n_particles + c->n_taxa > 2147483000LL) { set_error("too many nodes for int32 references"); return VCSMC_ERR_ARG; }
  return VCSMC_OK;
}

}  // namespace

extern "C" {

int vcsmc_sweep_query(const vcsmc_sweep_config* cfg, vcsmc_sweep_sizes* out) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!out) { set_error("null out"); return VCSMC_ERR_ARG; }
  vcsmc_sweep tmp;
  tmp.N = cfg->n_taxa; tmp.S = cfg->n_sites; tmp.K = cfg->n_particles; tmp.jc = cfg->jc; tmp.keep = cfg->keep_for_backward; tmp.M = cfg->n_sub;
  const int64_t tables = plan(&tmp);
  const int64_t full = full_pool_bytes(*cfg);
  out->retain_bytes = tables + (cfg->keep_for_backward ? 2 * full : full);
  const int64_t node = (int64_t)cfg->n_sites * 32;
  const int64_t K = cfg->n_particles;
  int64_t gc_min = 8192 + 2 * K * (node + 8) + 512;
  if (cfg->keep_for_backward) {
    const int64_t sc = cfg->n_sites < 256 ? cfg->n_sites : 256;
    const int64_t chunk_min = 2 * (int64_t)(cfg->n_taxa - 1) * K * 32 * sc + 4 * 2 * K + 8192;
    if (chunk_min > gc_min) gc_min = chunk_min;
  }
  out->min_bytes = tables + gc_min;
  if (out->min_bytes > out->retain_bytes) out->min_bytes = out->retain_bytes;
  return VCSMC_OK;
}

int vcsmc_sweep_create(const vcsmc_sweep_config* cfg, void* workspace, vcsmc_sweep_t** out) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!workspace || !out) { set_error("null workspace/out"); return VCSMC_ERR_ARG; }
  if (((uintptr_t)workspace & 255) != 0) { set_error("workspace must be 256-byte aligned"); return VCSMC_ERR_ARG; }
  vcsmc_sweep* h = new (std::nothrow) vcsmc_sweep();
  if (!h) { set_error("out of host memory"); return VCSMC_ERR_ARG; }
  h->N = cfg->n_taxa; h->S = cfg->n_sites; h->K = cfg->n_particles; h->jc = cfg->jc; h->keep = cfg->keep_for_backward;
  h->M = cfg->n_sub;
  h->ws = (char*)workspace; h->ws_bytes = cfg->workspace_bytes;
  h->Kl = h->K; h->k0 = 0; h->peer_ws[0] = h->ws;
  const int64_t tables = plan(h);
  rc = decide_modes(h, tables, cfg->workspace_bytes, true);
  if (rc) { delete h; return rc; }
  if (cudaMemset(h->ws + h->o_epoch_dev, 0, 4 * sizeof(int32_t)) != cudaSuccess) cudaGetLastError();  // (no device: entry points fail later)
  *out = h;
  return VCSMC_OK;
}

void vcsmc_sweep_destroy(vcsmc_sweep_t* h) { delete h; }

int vcsmc_sweep_set_allreduce(vcsmc_sweep_t* h, vcsmc_allreduce_fn fn, void* user) {
  if (!h) return VCSMC_ERR_ARG;
  h->allreduce = fn;
  h->allreduce_user = user;
  return VCSMC_OK;
}

int vcsmc_sweep_set_comm(vcsmc_sweep_t* h, int rank, int world, vcsmc_comm_fn fn, void* user, void* const* peer_ws_host) {
  if (!h) return VCSMC_ERR_ARG;
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) { set_error("set_comm: rank %d / world %d out of range (max %d ranks)", rank, world, kMaxPeers); return VCSMC_ERR_ARG; }
  if (h->M > 0 && world > 1) { set_error("particle sharding supports the VCSMC proposal only (n_sub == 0)"); return VCSMC_ERR_STATE; }
  if (h->K % world != 0) { set_error("n_particles = %lld is not divisible by %d ranks", (long long)h->K, world); return VCSMC_ERR_ARG; }
  if (world > 1 && !peer_ws_host) { set_error("set_comm: null peer table"); return VCSMC_ERR_ARG; }
  if (world > 1 && !fn) h->peer_sync = 1;
  if (h->allreduce && world > 1) { set_error("site sharding (set_allreduce) and particle sharding (set_comm) are exclusive"); return VCSMC_ERR_STATE; }
  h->rank = rank; h->world = world; h->comm = fn; h->comm_user = user;
  h->Kl = h->K / world; h->k0 = h->Kl * rank;
  for (int g = 0; g < kMaxPeers; ++g) h->peer_ws[g] = (world > 1 && g < world) ? (char*)peer_ws_host[g] : nullptr;
  h->peer_ws[rank] = h->ws;
  if (world > 1) h->lazy = 1;
  h->forward_done = false;
  const int64_t tables = plan(h);
  const int rc = decide_modes(h, tables, h->ws_bytes, true);
  if (rc) return rc;
  // flag array of the peer barrier: zero before anybody signals (the caller synchronises the ranks after this call)
  VCSMC_CUDA(cudaMemset(h->ws + h->o_sig, 0, 64 * sizeof(int32_t)));
  VCSMC_CUDA(cudaMemset(h->ws + h->o_epoch_dev, 0, 4 * sizeof(int32_t)));
  if (h->fwd_graph) { cudaGraphExecDestroy(h->fwd_graph); h->fwd_graph = nullptr; }
  return VCSMC_OK;
}

int vcsmc_sweep_set_option(vcsmc_sweep_t* h, const char* name, double value) {
  if (!h || !name) return VCSMC_ERR_ARG;
  if (!strcmp(name, "scalar_share")) h->scalar_share = value;
  else if (!strcmp(name, "skip_zero")) h->skip_zero = value != 0.0;
  else if (!strcmp(name, "skip_below")) h->skip_below = value < 0.0 ? 0.0 : value;
  else if (!strcmp(name, "max_chunk_sites")) h->max_chunk_sites = (int)value;
  else if (!strcmp(name, "lazy")) {
    if (value == 0.0 && h->world > 1) { set_error("particle sharding needs the lazy forward"); return VCSMC_ERR_STATE; }
    h->lazy = value != 0.0;
  }
  else if (!strcmp(name, "force_gc")) {  // testing aid: garbage-collected pool + recompute backward even when every node would fit
    h->force_gc = value != 0.0;
    h->forward_done = false;
    if (h->fwd_graph) { cudaGraphExecDestroy(h->fwd_graph); h->fwd_graph = nullptr; }
    const int64_t tables = plan(h);
    return decide_modes(h, tables, h->ws_bytes, true);
  }
  else if (!strcmp(name, "leaf_patterns")) { h->leaf_patterns = value != 0.0; if (h->fwd_graph) { cudaGraphExecDestroy(h->fwd_graph); h->fwd_graph = nullptr; } }
  else if (!strcmp(name, "force_sorted")) { h->force_sorted = value != 0.0; if (h->fwd_graph) { cudaGraphExecDestroy(h->fwd_graph); h->fwd_graph = nullptr; } }
  else if (!strcmp(name, "sparse_bwd")) h->sparse_bwd = value != 0.0;
  else if (!strcmp(name, "score_streams")) { h->score_streams = value != 0.0; if (h->fwd_graph) { cudaGraphExecDestroy(h->fwd_graph); h->fwd_graph = nullptr; } }
  else if (!strcmp(name, "leaf_rows")) { h->leaf_rows = value != 0.0; if (h->fwd_graph) { cudaGraphExecDestroy(h->fwd_graph); h->fwd_graph = nullptr; } }
  else if (!strcmp(name, "graph")) h->use_graph = value != 0.0;
  else if (!strcmp(name, "event_timing")) { h->event_timing = value != 0.0; if (h->fwd_graph) { cudaGraphExecDestroy(h->fwd_graph); h->fwd_graph = nullptr; } }
  else if (!strcmp(name, "peer_sync")) {
    if (value == 0.0 && !h->comm && h->world > 1) { set_error("peer_sync = 0 needs the collective hook"); return VCSMC_ERR_STATE; }
    h->peer_sync = value != 0.0;
    if (h->fwd_graph) { cudaGraphExecDestroy(h->fwd_graph); h->fwd_graph = nullptr; }
  }
  else if (!strcmp(name, "site_begin")) h->site_begin = (int)value;
  else if (!strcmp(name, "site_end")) h->site_end = (int)value;
  else if (!strcmp(name, "profile")) { h->profile = value != 0.0; h->ev_used = 0; h->ev_kind.clear(); }
  else { set_error("unknown option %s", name); return VCSMC_ERR_ARG; }
  return VCSMC_OK;
}

int vcsmc_sweep_set_uniforms(vcsmc_sweep_t* h, const float* u_pair, const double* u_bl, const double* u_br,
                             const double* u_res) {
  if (!h || !u_pair || !u_bl || !u_br || !u_res) { set_error("null uniforms"); return VCSMC_ERR_ARG; }
  h->x_pair = u_pair; h->x_bl = u_bl; h->x_br = u_br; h->x_res = u_res;
  h->use_seed = false;
  return VCSMC_OK;
}

int vcsmc_sweep_set_uniforms_nested(vcsmc_sweep_t* h, const double* look_bl, const double* look_br, const double* cat,
                                    const double* res) {
  if (!h || !look_bl || !look_br || !cat || !res) { set_error("null uniforms"); return VCSMC_ERR_ARG; }
  if (h->M <= 0) { set_error("set_uniforms_nested on a sweep created with n_sub = 0"); return VCSMC_ERR_STATE; }
  h->x_look_bl = look_bl; h->x_look_br = look_br; h->x_cat = cat; h->x_res = res;
  h->use_seed = false;
  return VCSMC_OK;
}

int vcsmc_sweep_set_seed(vcsmc_sweep_t* h, uint64_t seed) {
  if (!h) return VCSMC_ERR_ARG;
  h->seed = seed;
  h->use_seed = true;
  return VCSMC_OK;
}

int vcsmc_sweep_forward(vcsmc_sweep_t* h, const uint8_t* codes, const double* lam_l, const double* lam_r,
                        const double* Q, const double* pi, void* stream) {
  if (!h || !codes || !lam_l || !lam_r || !pi || (!h->jc && !Q)) { set_error("sweep_forward: null argument"); return VCSMC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->N, S = h->S;
  const int64_t K = h->K;
  h->codes = codes; h->lam_l = lam_l; h->lam_r = lam_r; h->Q = Q; h->pi = pi;
  h->forward_done = false;
  if (h->M == 0 && h->lazy) {
    const int rc = sweep_forward_lazy(h, codes, lam_l, lam_r, Q, pi, st);
    if (rc == VCSMC_OK) h->forward_done = true;
    return rc;
  }
  if (h->world > 1) { set_error("particle sharding needs the lazy forward"); return VCSMC_ERR_STATE; }

  // status + log-double-factorial table (host -> device, tiny)
  VCSMC_CUDA(cudaMemsetAsync(h->p<int32_t>(h->o_status), 0, 8 * sizeof(int32_t), st));
  {
    std::vector<double> ldf(2 * N + 4, 0.0);
    for (int m = 0; m < 2 * N + 4; ++m) ldf[m] = log_double_factorial_host(m);
    VCSMC_CUDA(cudaMemcpyAsync(h->p<double>(h->o_ldf), ldf.data(), ldf.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    VCSMC_CUDA(cudaStreamSynchronize(st));  // ldf is a stack-lifetime host buffer
  }
  double* ell_node = h->p<double>(h->o_ell_node);
  leaf_ell_kernel<<<N, 256, 0, st>>>(codes, S, S, pi, ell_node);
  VCSMC_LAUNCH_CHECK("leaf_ell_kernel");
  if (h->allreduce) {
    int rc = h->allreduce(h->allreduce_user, ell_node, N, st);
    if (rc) { set_error("allreduce hook failed (%d)", rc); return VCSMC_ERR_CUDA; }
  }
  double* pool = h->p<double>(h->o_pool);
  int32_t* flags = h->p<int32_t>(h->o_flags);
  int64_t pair_off = 0, look_off = 0;
  if (!h->use_seed && ((h->M > 0) != (h->x_look_bl != nullptr))) { set_error("uniforms were set for the other proposal (nested vs plain)"); return VCSMC_ERR_STATE; }

  for (int r = 0; r < N - 1; ++r) {
    const int n = N - r;
    const int cur = r & 1, prev = cur ^ 1;
    // uniforms of this rank event
    const float* u_pair = nullptr; const double *u_bl = nullptr, *u_br = nullptr, *u_res = nullptr, *u_cat = nullptr;
    const double *lk_bl = nullptr, *lk_br = nullptr;
    if (h->M > 0) {
      if (h->use_seed) {
        int rc = launch_philox_step(h->seed, r, 0, K, 0, nullptr, nullptr, nullptr, h->p<double>(h->o_u_res), h->p<double>(h->o_u_cat), st);
        if (rc) return rc;
        u_res = h->p<double>(h->o_u_res); u_cat = h->p<double>(h->o_u_cat);
      } else {
        lk_bl = h->x_look_bl + look_off; lk_br = h->x_look_br + look_off;
        look_off += (int64_t)(n * (n - 1) / 2) * h->M * K;
        u_cat = h->x_cat + (int64_t)r * K; u_res = h->x_res + (int64_t)r * K;
      }
    } else if (h->use_seed) {
      int rc = launch_philox_step(h->seed, r, 0, K, n, h->p<float>(h->o_u_pair), h->p<double>(h->o_u_bl),
                                  h->p<double>(h->o_u_br), h->p<double>(h->o_u_res), nullptr, st);
      if (rc) return rc;
      u_pair = h->p<float>(h->o_u_pair); u_bl = h->p<double>(h->o_u_bl); u_br = h->p<double>(h->o_u_br); u_res = h->p<double>(h->o_u_res);
    } else {
      u_pair = h->x_pair + pair_off; u_bl = h->x_bl + (int64_t)r * K; u_br = h->x_br + (int64_t)r * K; u_res = h->x_res + (int64_t)r * K;
      pair_off += K * n;
    }
    PrepArgs a;
    a.r = r; a.n = n; a.N = N; a.gc = h->fwd_gc; a.K = K;
    a.cdf = h->p<double>(h->o_cdf); a.u_res = u_res; a.u_pair = u_pair; a.u_bl = u_bl; a.u_br = u_br;
    a.lam_l = lam_l; a.lam_r = lam_r;
    a.ids_old = h->p<int32_t>(h->o_ids[prev]); a.cnt_old = h->p<int32_t>(h->o_cnt[prev]); a.slot_old = h->p<int32_t>(h->o_slot[prev]);
    a.ids_new = h->p<int32_t>(h->o_ids[cur]); a.cnt_new = h->p<int32_t>(h->o_cnt[cur]); a.slot_new = h->p<int32_t>(h->o_slot[cur]);
    a.LL_prev = r > 0 ? h->p<double>(h->o_LL) + (int64_t)(r - 1) * K : nullptr;
    a.anc = h->p<int32_t>(h->o_anc) + (int64_t)r * K;
    a.lref = h->p<int32_t>(h->o_lref) + (int64_t)r * K;
    a.rref = h->p<int32_t>(h->o_rref) + (int64_t)r * K;
    a.nleaf = h->p<int32_t>(h->o_nleaf) + (int64_t)r * K;
    a.rempos = h->p<uint8_t>(h->o_rempos) + h->rem_off[r];
    a.b_l = h->p<double>(h->o_b_l) + (int64_t)r * K;
    a.b_r = h->p<double>(h->o_b_r) + (int64_t)r * K;
    a.t2 = h->p<double>(h->o_t2) + (int64_t)r * 2 * K;
    a.ll_tilde = h->p<double>(h->o_lltilde);
    a.lsrc = h->p<int32_t>(h->o_lsrc); a.rsrc = h->p<int32_t>(h->o_rsrc); a.dst = h->p<int32_t>(h->o_dst);
    int rc;
    if (h->M > 0) {
      // VNCSMC: inherit the ancestor's forest, score every (pair, sub-sample), draw one option per particle
      int32_t* inh_ids = h->p<int32_t>(h->o_inh_ids);
      int32_t* inh_cnt = h->p<int32_t>(h->o_inh_cnt);
      int32_t* inh_slot = h->p<int32_t>(h->o_inh_slot);
      rc = launch_nested_inherit(r, n, N, K, a.cdf, u_res, a.ids_old, a.cnt_old, a.slot_old, inh_ids, inh_cnt, inh_slot,
                                 h->keep ? h->p<int32_t>(h->o_rows_all) + (int64_t)r * K * N : nullptr, a.LL_prev, a.anc, a.ll_tilde, st);
      if (rc) return rc;
      double* pot = h->p<double>(h->o_pot) + h->pot_off[r];
      h->prof_begin(3, st);
      rc = launch_lookahead(r, n, N, h->M, h->jc, h->fwd_gc, S, K, inh_ids, inh_cnt, inh_slot, codes, S, pool, S, ell_node,
                            h->p<double>(h->o_ldf), Q, pi, lam_l, lam_r, lk_bl, lk_br, h->seed, pot, h->allreduce ? h->scalar_share : 1.0, st);
      h->prof_end(st);
      if (rc) return rc;
      if (h->allreduce) {   // site sharding: the look-ahead's site sums are completed across ranks before the options are drawn
        rc = h->allreduce(h->allreduce_user, pot, K * (int64_t)(n * (n - 1) / 2) * h->M, st);
        if (rc) { set_error("allreduce hook failed (%d)", rc); return VCSMC_ERR_CUDA; }
      }
      rc = launch_nested_choose(r, n, N, h->M, h->fwd_gc, K, pot, u_cat, lk_bl, lk_br, h->seed, lam_l, lam_r, inh_ids, inh_cnt,
                                inh_slot, a.ids_new, a.cnt_new, a.slot_new, a.lref, a.rref, a.nleaf, a.rempos,
                                h->p<int32_t>(h->o_choice) + (int64_t)r * K, a.b_l, a.b_r, a.t2, h->p<double>(h->o_qlog),
                                a.lsrc, a.rsrc, a.dst, st);
      if (rc) return rc;
    } else {
      step_prepare_kernel<<<(unsigned)((K + kPrepWarps - 1) / kPrepWarps), kPrepWarps * 32, (size_t)kPrepWarps * n * sizeof(float), st>>>(a);
      VCSMC_LAUNCH_CHECK("step_prepare_kernel");
    }

    double* P = h->p<double>(h->o_P) + (int64_t)r * K * 32;
    rc = launch_transition_fwd(Q, a.t2, 2 * K, h->jc, P, st);
    if (rc) return rc;

    if (h->fwd_gc) {
      VCSMC_CUDA(cudaMemsetAsync(flags, 0, (size_t)h->pool_slots * sizeof(int32_t), st));
      count_launch();
      gc_mark_kernel<<<(unsigned)((K * n + 255) / 256), 256, 0, st>>>(a.ids_new, a.slot_new, N, n, K, a.lsrc, a.rsrc, flags);
      VCSMC_LAUNCH_CHECK("gc_mark_kernel");
      gc_alloc_kernel<<<1, 1024, 0, st>>>(flags, h->pool_slots, K, N, n, a.dst, a.slot_new, h->p<int32_t>(h->o_status));
      VCSMC_LAUNCH_CHECK("gc_alloc_kernel");
    }
    // visiting order: particles sorted by their pair of child nodes (shared children are read once per group)
    const bool sorted = use_sorted_order(K, S);
    if (sorted) {
      // (a full sort, not just grouping: consecutive groups then also share their first child)
      rc = launch_sort_order(a.lsrc, a.rsrc, nullptr, K, h->pool_slots, h->p<uint64_t>(h->o_keys_in), h->p<uint64_t>(h->o_keys_out),
                             h->p<int32_t>(h->o_vals_in), h->p<int32_t>(h->o_order), h->p<int32_t>(h->o_count),
                             h->p<char>(h->o_sort_temp), h->sort_temp, st);
      if (rc) return rc;
    }
    int tiles = 0;  // partial sums per particle written by the merge
    h->prof_begin(0, st);
    rc = launch_merge_fwd(codes, S, pool, S, a.lsrc, a.rsrc, a.dst, sorted ? h->p<int32_t>(h->o_order) : nullptr, nullptr, P, pi, K, -1, S, h->jc, 0,
                          h->p<double>(h->o_ell_part), &tiles, st);
    h->prof_end(st);
    if (rc) return rc;

    WeightArgs w;
    w.r = r; w.n = n; w.N = N; w.tiles = tiles; w.K = K;
    w.ell_part = h->p<double>(h->o_ell_part);
    if (h->allreduce) {
      rc = launch_ell_reduce(h->p<double>(h->o_ell_part), tiles, K, h->p<double>(h->o_ell_new), st);
      if (rc) return rc;
      rc = h->allreduce(h->allreduce_user, h->p<double>(h->o_ell_new), K, st);
      if (rc) { set_error("allreduce hook failed (%d)", rc); return VCSMC_ERR_CUDA; }
      w.ell_part = h->p<double>(h->o_ell_new);
      w.tiles = 1;
    }
    w.ids_new = a.ids_new; w.cnt_new = a.cnt_new; w.ldf = h->p<double>(h->o_ldf);
    w.lam_l = lam_l; w.lam_r = lam_r; w.b_l = a.b_l; w.b_r = a.b_r;
    w.cum_l_prev = r > 0 ? h->p<double>(h->o_cum_l) + (int64_t)(r - 1) * K : nullptr;
    w.cum_r_prev = r > 0 ? h->p<double>(h->o_cum_r) + (int64_t)(r - 1) * K : nullptr;
    w.cum_l = h->p<double>(h->o_cum_l) + (int64_t)r * K;
    w.cum_r = h->p<double>(h->o_cum_r) + (int64_t)r * K;
    w.ll_tilde = a.ll_tilde; w.ell_node = ell_node;
    w.lw = h->p<double>(h->o_lw) + (int64_t)r * K;
    w.LL = h->p<double>(h->o_LL) + (int64_t)r * K;
    w.vminus = h->p<int32_t>(h->o_vminus);
    w.q = 1.0 / ((double)n * (double)(n - 1) / 2.0);
    w.e_off = N + (int64_t)r * K;
    w.qlog = h->M > 0 ? h->p<double>(h->o_qlog) : nullptr;
    step_weights_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(w);
    VCSMC_LAUNCH_CHECK("step_weights_kernel");

    // log-sum-exp + CDF of this step's weights: logZ_r now, ancestors of the next rank event
    rc = launch_resample_cdf(w.lw, K, h->p<double>(h->o_cdf), h->p<double>(h->o_stats) + r * 4, h->p<double>(h->o_cdf_scratch), st);
    if (rc) return rc;
  }
  finalize_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(
      N, K, h->p<double>(h->o_stats), h->p<double>(h->o_LL) + (int64_t)(N - 2) * K, h->p<double>(h->o_b_l),
      h->p<double>(h->o_b_r), lam_l, lam_r, log_double_factorial_host(2 * N - 3), h->p<double>(h->o_llR),
      h->p<double>(h->o_elbo), h->p<double>(h->o_logz), h->p<double>(h->o_ess));
  VCSMC_LAUNCH_CHECK("finalize_kernel");
  h->forward_done = true;
  return VCSMC_OK;
}

int vcsmc_sweep_backward(vcsmc_sweep_t* h, double grad_elbo, double* dlam_l, double* dlam_r, double* dQ, double* dpi,
                         void* stream) {
  if (!h || !dlam_l || !dlam_r || !dpi || (!h->jc && !dQ)) { set_error("sweep_backward: null argument"); return VCSMC_ERR_ARG; }
  if (!h->keep) { set_error("sweep was created with keep_for_backward = 0"); return VCSMC_ERR_STATE; }
  if (!h->forward_done) { set_error("sweep_backward called before a successful sweep_forward"); return VCSMC_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->N, S = h->S;
  const int64_t K = h->K, E = (int64_t)(N - 1) * K;
  int rc;
  // site slice of this rank's reverse sweep (particle-sharded runs gather the scalar tables and shard the backward by site)
  const int sb = h->site_begin < 0 ? 0 : (h->site_begin > S ? S : h->site_begin);
  const int se = (h->site_end < 0 || h->site_end > S) ? S : (h->site_end < sb ? sb : h->site_end);
  if (h->retain && (sb != 0 || se != S)) { set_error("a site slice needs the recompute backward (nodes are not retained per site slice)"); return VCSMC_ERR_STATE; }

  // ---- zero accumulators
  VCSMC_CUDA(cudaMemsetAsync(dlam_l, 0, (N - 1) * sizeof(double), st));
  VCSMC_CUDA(cudaMemsetAsync(dlam_r, 0, (N - 1) * sizeof(double), st));
  if (dQ) VCSMC_CUDA(cudaMemsetAsync(dQ, 0, 16 * sizeof(double), st));
  VCSMC_CUDA(cudaMemsetAsync(dpi, 0, 4 * sizeof(double), st));
  VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_childsum[0]), 0, K * sizeof(double), st));
  VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_childsum[1]), 0, K * sizeof(double), st));
  VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_Dacc[0]), 0, K * N * sizeof(double), st));
  VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_Dacc[1]), 0, K * N * sizeof(double), st));
  VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_consumed), 0, E * sizeof(int32_t), st));
  if (h->M > 0) VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_dP), 0, 32 * E * sizeof(double), st));  // (VCSMC: only the visited rows, below)
  VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_count_bwd), 0, N * sizeof(int32_t), st));
  VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_dQ_acc), 0, 16 * K * sizeof(double), st));
  VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_suf_l), 0, K * sizeof(double), st));
  VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_suf_r), 0, K * sizeof(double), st));
  count_launch(14);

  if (h->world > 1 && h->M > 0) {
    // the forward filled P only for this rank's particles: rebuild all of it from the gathered branch lengths
    // (the VCSMC path rebuilds the visited particles' matrices only, once the visit lists are known)
    rc = launch_transition_fwd(h->Q, h->p<double>(h->o_t2), 2 * E, h->jc, h->p<double>(h->o_P), st);
    if (rc) return rc;
  }
  // ---- which nodes are ever consumed as a child; child/adjoint slots of every event
  mark_consumed_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(h->p<int32_t>(h->o_lref), h->p<int32_t>(h->o_rref), E, N, h->p<int32_t>(h->o_consumed));
  VCSMC_LAUNCH_CHECK("mark_consumed_kernel");
  std::vector<int64_t> n_act(N, 0);
  if (h->M > 0) {
    // VNCSMC: the look-ahead of every active particle reads (and sends adjoints to) ALL roots of its inherited forest
    std::vector<int32_t> last_base(N, 0), last_flag(N, 0);
    for (int r = 0; r < N - 1; ++r) {
      const int n = N - r;
      const int64_t tot = K * (int64_t)(n * (n - 1) / 2) * h->M;
      rc = launch_nested_active(r, K, h->skip_zero, h->skip_below, h->p<double>(h->o_lw) + (int64_t)r * K, h->p<double>(h->o_stats), h->p<int32_t>(h->o_nact), st);
      if (rc) return rc;
      rc = launch_nested_mark_roots(n, N, K, h->p<int32_t>(h->o_nact), h->p<int32_t>(h->o_rows_all) + (int64_t)r * K * N, h->p<int32_t>(h->o_consumed), st);
      if (rc) return rc;
      rc = launch_nested_keep(r, n, h->M, K, grad_elbo, !h->skip_zero, h->skip_below * fabs(grad_elbo), h->p<double>(h->o_lw) + (int64_t)r * K, h->p<double>(h->o_stats),
                              h->p<double>(h->o_pot) + h->pot_off[r], h->p<int32_t>(h->o_choice) + (int64_t)r * K, h->p<int32_t>(h->o_v_keep), st);
      if (rc) return rc;
      rc = launch_exclusive_scan_i32(h->p<int32_t>(h->o_v_keep), h->p<int32_t>(h->o_v_index), tot, h->p<char>(h->o_v_scan), h->v_scan, st);
      if (rc) return rc;
      VCSMC_CUDA(cudaMemcpyAsync(&last_base[r], h->p<int32_t>(h->o_v_index) + (tot - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      VCSMC_CUDA(cudaMemcpyAsync(&last_flag[r], h->p<int32_t>(h->o_v_keep) + (tot - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    VCSMC_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < N - 1; ++r) n_act[r] = (int64_t)last_base[r] + last_flag[r];   // virtual events to visit
  }
  const int32_t* slot_of = nullptr;
  if (!h->retain) {
    rc = launch_exclusive_scan_i32(h->p<int32_t>(h->o_consumed), h->p<int32_t>(h->o_cslot), E, h->p<char>(h->o_sort_temp), h->sort_temp, st);
    if (rc) return rc;
    slot_of = h->p<int32_t>(h->o_cslot);
  }
  bwd_src_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(h->p<int32_t>(h->o_lref), h->p<int32_t>(h->o_rref), h->p<int32_t>(h->o_consumed), slot_of, E, N,
                                                              h->p<int32_t>(h->o_bsrc_l), h->p<int32_t>(h->o_bsrc_r), h->p<int32_t>(h->o_bsrc_g), h->p<int32_t>(h->o_bdst));
  VCSMC_LAUNCH_CHECK("bwd_src_kernel");

  // ---- scalar pass: coefficients of every node, site-independent gradient terms
  const bool fused = h->M == 0;
  if (fused) {
    ScalarArgs a;
    a.N = N; a.skip_zero = h->skip_zero; a.K = K; a.grad = grad_elbo; a.share = h->scalar_share;
    a.thresh = h->skip_below * fabs(grad_elbo);
    a.lw = h->p<double>(h->o_lw); a.stats = h->p<double>(h->o_stats); a.anc = h->p<int32_t>(h->o_anc);
    a.rempos = h->p<uint8_t>(h->o_rempos);
    for (int i = 0; i < 2; ++i) {
      a.childsum[i] = h->p<double>(h->o_childsum[i]); a.Dacc[i] = h->p<double>(h->o_Dacc[i]);
      a.dirty[i] = h->p<int32_t>(h->o_dirty) + (int64_t)i * K;
    }
    VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_dirty), 0, 2 * K * sizeof(int32_t), st));
    a.cnew = h->p<double>(h->o_cnew); a.lam_l = h->lam_l; a.lam_r = h->lam_r;
    a.b_l = h->p<double>(h->o_b_l); a.b_r = h->p<double>(h->o_b_r);
    a.cum_l = h->p<double>(h->o_cum_l); a.cum_r = h->p<double>(h->o_cum_r);
    a.suf_l = h->p<double>(h->o_suf_l); a.suf_r = h->p<double>(h->o_suf_r);
    a.dlam_l = dlam_l; a.dlam_r = dlam_r;
    a.consumed = h->p<int32_t>(h->o_consumed); a.act_all = h->p<int32_t>(h->o_act_all);
    a.order_bwd = h->p<int32_t>(h->o_order_bwd); a.count_bwd = h->p<int32_t>(h->o_count_bwd);
    static int max_blocks = 0;
    if (max_blocks == 0) {
      int per_sm = 0, sms = 0, dev = 0;
      VCSMC_CUDA(cudaGetDevice(&dev));
      VCSMC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      VCSMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bwd_scalar_kernel, 256, 0));
      max_blocks = per_sm * sms;
      if (max_blocks < 1) { set_error("bwd_scalar_kernel cannot be launched cooperatively"); return VCSMC_ERR_CUDA; }
    }
    int64_t blocks = (K + 255) / 256;
    if (blocks > max_blocks) blocks = max_blocks;
    void* kargs[] = {(void*)&a};
    VCSMC_CUDA(cudaLaunchCooperativeKernel((const void*)bwd_scalar_kernel, dim3((unsigned)blocks), dim3(256), kargs, 0, st));
    count_launch();
  } else {
  for (int r = N - 2; r >= 0; --r) {
    const int cur = r & 1, nxt = cur ^ 1;
    CoefArgs a;
    a.r = r; a.n = N - r; a.N = N; a.K = K; a.grad = grad_elbo; a.share = h->scalar_share;
    a.lw = h->p<double>(h->o_lw) + (int64_t)r * K;
    a.stats = h->p<double>(h->o_stats);
    a.anc = h->p<int32_t>(h->o_anc) + (int64_t)r * K;
    a.rempos = h->p<uint8_t>(h->o_rempos) + h->rem_off[r];
    a.childsum_cur = h->p<double>(h->o_childsum[cur]); a.childsum_next = h->p<double>(h->o_childsum[nxt]);
    a.Dacc_cur = h->p<double>(h->o_Dacc[cur]); a.Dacc_next = h->p<double>(h->o_Dacc[nxt]);
    a.cnew = h->p<double>(h->o_cnew) + (int64_t)r * K;
    a.lam_l = h->lam_l; a.lam_r = h->lam_r;
    a.b_l = h->p<double>(h->o_b_l) + (int64_t)r * K; a.b_r = h->p<double>(h->o_b_r) + (int64_t)r * K;
    a.cum_l = h->p<double>(h->o_cum_l) + (int64_t)r * K; a.cum_r = h->p<double>(h->o_cum_r) + (int64_t)r * K;
    a.suf_l = h->p<double>(h->o_suf_l); a.suf_r = h->p<double>(h->o_suf_r);
    a.dlam_l = dlam_l; a.dlam_r = dlam_r;
    bwd_coef_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(a);
    VCSMC_LAUNCH_CHECK("bwd_coef_kernel");
    if (h->M > 0) {
      rc = launch_nested_coef(r, N - r, N, h->M, K, grad_elbo, a.lw, a.stats, h->p<double>(h->o_pot) + h->pot_off[r],
                              h->p<int32_t>(h->o_choice) + (int64_t)r * K, a.anc, a.Dacc_next, st);
      if (rc) return rc;
    }
    // the buffers just consumed become the accumulation targets of step r-2
    VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_childsum[cur]), 0, K * sizeof(double), st));
    VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_Dacc[cur]), 0, K * N * sizeof(double), st));
    count_launch(2);
  }
  }
  // after r = 0 the scatter target was Dacc[(0&1)^1] = Dacc[1]: per-particle coefficients of the N leaves
  VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_cleaf), 0, N * sizeof(double), st));
  column_sum_kernel<<<N, 256, 0, st>>>(h->p<double>(h->o_Dacc[1]), K, N, N, h->p<double>(h->o_cleaf));
  VCSMC_LAUNCH_CHECK("column_sum_kernel");
  if (se > sb) {
    leaf_pi_grad_kernel<<<N, 256, 0, st>>>(h->codes + sb, S, se - sb, h->pi, h->p<double>(h->o_cleaf), dpi);
    VCSMC_LAUNCH_CHECK("leaf_pi_grad_kernel");
  }

  // ---- visiting orders of every rank event
  const bool sorted_b = fused || use_sorted_order(K, S);
  std::vector<int32_t> cnt_bwd(N, (int32_t)K), cnt_rec(N, (int32_t)K);
  if (fused) {
    // the scalar pass listed the particles to visit; one small D2H + sync tells the host how many per event
    VCSMC_CUDA(cudaMemcpyAsync(cnt_bwd.data(), h->p<int32_t>(h->o_count_bwd), (N - 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VCSMC_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < N - 1; ++r) {
      // long lists are worth arranging by child pair (shared children are then read once per run, adjoints flushed
      // once per run); the dense reverse sweep gets the full sort, as before
      if (h->skip_zero && cnt_bwd[r] < 1024) continue;
      const int32_t* bl = h->p<int32_t>(h->o_bsrc_l) + (int64_t)r * K;
      const int32_t* br = h->p<int32_t>(h->o_bsrc_r) + (int64_t)r * K;
      const int32_t* act = h->p<int32_t>(h->o_act_all) + (int64_t)r * K;
      if (h->skip_zero)
        rc = group_particles(h, bl, br, act, K, h->p<int32_t>(h->o_order_bwd) + (int64_t)r * K, h->p<int32_t>(h->o_count_bwd) + r, st);
      else
        rc = launch_sort_order(bl, br, act, K, E, h->p<uint64_t>(h->o_keys_in), h->p<uint64_t>(h->o_keys_out), h->p<int32_t>(h->o_vals_in),
                               h->p<int32_t>(h->o_order_bwd) + (int64_t)r * K, h->p<int32_t>(h->o_count_bwd) + r, h->p<char>(h->o_sort_temp), h->sort_temp, st);
      if (rc) return rc;
    }
    cnt_rec = cnt_bwd;
    int32_t max_cnt = 0;
    for (int r = 0; r < N - 1; ++r) max_cnt = cnt_bwd[r] > max_cnt ? cnt_bwd[r] : max_cnt;
    if (max_cnt > 0 && h->world > 1) {
      const int64_t bx = (2 * (int64_t)max_cnt + 63) / 64;
      transition_visited_kernel<<<dim3((unsigned)(bx < 4096 ? bx : 4096), N - 1), 64, 0, st>>>(
          h->Q, h->p<double>(h->o_t2), h->p<int32_t>(h->o_order_bwd), h->p<int32_t>(h->o_count_bwd), K, h->jc, h->p<double>(h->o_P));
      VCSMC_LAUNCH_CHECK("transition_visited_kernel");
    }
    if (max_cnt > 0) {
      const unsigned bx = (unsigned)((max_cnt + 7) / 8 < 2048 ? (max_cnt + 7) / 8 : 2048);
      zero_dP_rows_kernel<<<dim3(bx, N - 1), 256, 0, st>>>(h->p<int32_t>(h->o_order_bwd), h->p<int32_t>(h->o_count_bwd), K, h->p<double>(h->o_dP));
      VCSMC_LAUNCH_CHECK("zero_dP_rows_kernel");
    }
  } else
  if (sorted_b) {
    for (int r = 0; r < N - 1; ++r) {
      bwd_active_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(h->p<double>(h->o_cnew) + (int64_t)r * K, h->p<int32_t>(h->o_consumed) + (int64_t)r * K,
                                                                   K, h->skip_zero, h->skip_below * fabs(grad_elbo), h->p<int32_t>(h->o_act_bwd), h->p<int32_t>(h->o_act_rec));
      VCSMC_LAUNCH_CHECK("bwd_active_kernel");
      const int32_t* bl = h->p<int32_t>(h->o_bsrc_l) + (int64_t)r * K;
      const int32_t* br = h->p<int32_t>(h->o_bsrc_r) + (int64_t)r * K;
      // sparse reverse sweep: grouping by child pair is enough; dense: a full sort keeps runs of the first child together
      // (one flush of its adjoint per run instead of one per pair)
      if (h->skip_zero)
        rc = group_particles(h, bl, br, h->p<int32_t>(h->o_act_bwd), K, h->p<int32_t>(h->o_order_bwd) + (int64_t)r * K, h->p<int32_t>(h->o_count_bwd) + r, st);
      else
        rc = launch_sort_order(bl, br, h->p<int32_t>(h->o_act_bwd), K, E, h->p<uint64_t>(h->o_keys_in), h->p<uint64_t>(h->o_keys_out), h->p<int32_t>(h->o_vals_in),
                               h->p<int32_t>(h->o_order_bwd) + (int64_t)r * K, h->p<int32_t>(h->o_count_bwd) + r, h->p<char>(h->o_sort_temp), h->sort_temp, st);
      if (rc) return rc;
    }
    // one small D2H + sync: the host learns how many particles each rank event really has to visit, so that empty
    // launches are skipped and grids are sized exactly (with ESS ~ 1 almost every reverse event is empty)
    VCSMC_CUDA(cudaMemcpyAsync(cnt_bwd.data(), h->p<int32_t>(h->o_count_bwd), (N - 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VCSMC_CUDA(cudaStreamSynchronize(st));
    cnt_rec = cnt_bwd;
  }
  {
    int64_t visited = 0;
    for (int r = 0; r < N - 1; ++r) visited += cnt_bwd[r];
    int32_t v = (int32_t)(visited > 2147483647LL ? 2147483647LL : visited);
    VCSMC_CUDA(cudaMemcpyAsync(h->p<int32_t>(h->o_status) + 3, &v, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    VCSMC_CUDA(cudaStreamSynchronize(st));
  }

  // ---- per-site pass: reverse pruning, by site chunks
  int Sc = h->chunk_sites;
  double* lpool = h->p<double>(h->o_pool);
  double* gpool;
  if (h->retain) {
    gpool = lpool + E * (int64_t)S * 4;
  } else {
    // only nodes that a later merge consumes are materialised: the chunk is as long as their count allows
    int32_t last_slot = 0, last_flag = 0;  // number of consumed nodes = exclusive scan's last entry + last flag
    VCSMC_CUDA(cudaMemcpyAsync(&last_slot, h->p<int32_t>(h->o_cslot) + (E - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VCSMC_CUDA(cudaMemcpyAsync(&last_flag, h->p<int32_t>(h->o_consumed) + (E - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VCSMC_CUDA(cudaStreamSynchronize(st));
    int64_t n_cons = (int64_t)last_slot + last_flag;
    if (n_cons < 1) n_cons = 1;
    int64_t sc = (h->pool_bytes / 2) / (n_cons * 32);
    if (h->max_chunk_sites > 0 && sc > h->max_chunk_sites) sc = h->max_chunk_sites;
    if (sc >= se - sb) sc = se - sb;
    else sc = sc / 256 * 256;
    if (sc < 256 && sc < se - sb) { set_error("workspace too small for a 256-site backward chunk"); return VCSMC_ERR_ARG; }
    if (sc < 1) sc = 1;
    Sc = (int)sc;
    gpool = lpool + n_cons * (int64_t)Sc * 4;
  }
  // few particles to visit (ESS ~ 1): one site-parallel launch per chunk instead of three launches per rank event
  int64_t visited_total = 0;
  for (int r = 0; r < N - 1; ++r) visited_total += cnt_bwd[r];
  // (cost model: the site-parallel pass walks the visited particle-events one after the other, ~2.5 us each on a short alignment; the
  // per-event path pays three launches, ~6 us each, for every rank event that has something to visit)
  int nonempty = 0;
  for (int r = 0; r < N - 1; ++r) nonempty += cnt_bwd[r] > 0;
  const bool sparse = fused && h->skip_zero && h->sparse_bwd && visited_total <= 1024 && 2.5 * (double)visited_total < 18.0 * nonempty;
  int n_chunks = 0;
  for (int s0 = sb; s0 < se; s0 += Sc, ++n_chunks) {
    const int nc = (se - s0 < Sc) ? se - s0 : Sc;
    const uint8_t* codes_c = h->codes + s0;
    if (sparse) {
      h->prof_begin(2, st);
      rc = launch_bwd_sparse(codes_c, S, lpool, gpool, Sc, nc, N, K, h->retain ? 0 : 1, h->jc, h->skip_zero, h->skip_below * fabs(grad_elbo),
                             h->p<int32_t>(h->o_order_bwd), h->p<int32_t>(h->o_count_bwd), h->p<int32_t>(h->o_bsrc_l), h->p<int32_t>(h->o_bsrc_r),
                             h->p<int32_t>(h->o_bsrc_g), h->p<int32_t>(h->o_bdst), h->p<double>(h->o_P), h->pi, h->p<double>(h->o_cnew),
                             h->p<double>(h->o_dP), dpi, st);
      h->prof_end(st);
      if (rc) return rc;
      continue;
    }
    if (!h->retain) {
      // recompute the forward for this chunk, materialising only nodes that are consumed later
      for (int r = 0; r < N - 1; ++r) {
        if (cnt_rec[r] == 0) continue;
        h->prof_begin(1, st);
        rc = launch_merge_fwd(codes_c, S, lpool, Sc, h->p<int32_t>(h->o_bsrc_l) + (int64_t)r * K, h->p<int32_t>(h->o_bsrc_r) + (int64_t)r * K,
                              h->p<int32_t>(h->o_bdst) + (int64_t)r * K, sorted_b ? h->p<int32_t>(h->o_order_bwd) + (int64_t)r * K : nullptr, sorted_b ? h->p<int32_t>(h->o_count_bwd) + r : nullptr,
                              h->p<double>(h->o_P) + (int64_t)r * K * 32, h->pi, K, cnt_rec[r], nc, h->jc, 1, h->p<double>(h->o_ell_part), nullptr, st);
        h->prof_end(st);
        if (rc) return rc;
      }
    }
    // zero the adjoint slots of consumed nodes
    for (int r = 0; r < N - 1; ++r) {
      if (cnt_rec[r] == 0) continue;
      dim3 grid((nc + 255) / 256, cnt_rec[r] < 32 ? cnt_rec[r] : 32, 1);
      zero_consumed_kernel<<<grid, 256, 0, st>>>(sorted_b ? h->p<int32_t>(h->o_order_bwd) + (int64_t)r * K : nullptr, sorted_b ? h->p<int32_t>(h->o_count_bwd) + r : nullptr, (int)K,
                                                 h->p<int32_t>(h->o_bsrc_g) + (int64_t)r * K, Sc, nc, gpool);
      VCSMC_LAUNCH_CHECK("zero_consumed_kernel");
    }
    for (int r = N - 2; r >= 0; --r) {
      if (h->M > 0 && n_act[r] > 0) {
        // VNCSMC: every (active particle, pair, sub-sample) of this rank event as a virtual merge event
        const int n = N - r;
        const int64_t tot = K * (int64_t)(n * (n - 1) / 2) * h->M;
        const int64_t V = n_act[r];
        rc = launch_nested_keep(r, n, h->M, K, grad_elbo, !h->skip_zero, h->skip_below * fabs(grad_elbo), h->p<double>(h->o_lw) + (int64_t)r * K, h->p<double>(h->o_stats),
                                h->p<double>(h->o_pot) + h->pot_off[r], h->p<int32_t>(h->o_choice) + (int64_t)r * K, h->p<int32_t>(h->o_v_keep), st);
        if (rc) return rc;
        rc = launch_exclusive_scan_i32(h->p<int32_t>(h->o_v_keep), h->p<int32_t>(h->o_v_index), tot, h->p<char>(h->o_v_scan), h->v_scan, st);
        if (rc) return rc;
        const double *lk_bl = nullptr, *lk_br = nullptr;
        if (!h->use_seed) {
          int64_t off = 0;
          for (int q = 0; q < r; ++q) off += (int64_t)((N - q) * (N - q - 1) / 2) * h->M * K;
          lk_bl = h->x_look_bl + off; lk_br = h->x_look_br + off;
        }
        for (int64_t v0 = 0; v0 < V; v0 += h->v_batch) {
          const int64_t Vb = (V - v0 < h->v_batch) ? V - v0 : h->v_batch;
          rc = launch_nested_virtual(r, n, N, h->M, K, grad_elbo, !h->skip_zero, h->skip_below * fabs(grad_elbo), h->p<double>(h->o_lw) + (int64_t)r * K, h->p<double>(h->o_stats),
                                     h->p<double>(h->o_pot) + h->pot_off[r], h->p<int32_t>(h->o_choice) + (int64_t)r * K,
                                     h->p<int32_t>(h->o_v_index), h->p<int32_t>(h->o_rows_all) + (int64_t)r * K * N,
                                     slot_of, lk_bl, lk_br, h->seed, h->lam_l, h->lam_r, v0, v0 + Vb, h->p<int32_t>(h->o_v_lsrc),
                                     h->p<int32_t>(h->o_v_rsrc), h->p<double>(h->o_v_coef), h->p<double>(h->o_v_t2), st);
          if (rc) return rc;
          rc = launch_transition_fwd(h->Q, h->p<double>(h->o_v_t2), 2 * Vb, h->jc, h->p<double>(h->o_v_P), st);
          if (rc) return rc;
          VCSMC_CUDA(cudaMemsetAsync(h->p<char>(h->o_v_dP), 0, 32 * Vb * sizeof(double), st));
          rc = launch_sort_order(h->p<int32_t>(h->o_v_lsrc), h->p<int32_t>(h->o_v_rsrc), nullptr, Vb, E, h->p<uint64_t>(h->o_v_keys_in),
                                 h->p<uint64_t>(h->o_v_keys_out), h->p<int32_t>(h->o_v_vals), h->p<int32_t>(h->o_v_order),
                                 h->p<int32_t>(h->o_v_count), h->p<char>(h->o_v_temp), h->v_temp, st);
          if (rc) return rc;
          rc = launch_merge_bwd(codes_c, S, lpool, gpool, Sc, h->p<int32_t>(h->o_v_lsrc), h->p<int32_t>(h->o_v_rsrc), nullptr,
                                h->p<int32_t>(h->o_v_order), nullptr, h->p<double>(h->o_v_P), h->pi, h->p<double>(h->o_v_coef), Vb, Vb, nc,
                                h->jc, 1, 0.0, h->p<double>(h->o_v_dP), dpi, st);
          if (rc) return rc;
          rc = launch_transition_bwd(h->Q, h->p<double>(h->o_v_t2), h->p<double>(h->o_v_dP), 2 * Vb, h->jc, nullptr, h->p<double>(h->o_v_dt),
                                     h->jc ? nullptr : h->p<double>(h->o_v_dQ), st);
          if (rc) return rc;
          rc = launch_nested_reduce(r, Vb, h->jc, h->p<double>(h->o_v_dt), h->p<double>(h->o_v_dQ), h->p<double>(h->o_v_t2), h->lam_l,
                                    h->lam_r, dlam_l, dlam_r, dQ, st);
          if (rc) return rc;
        }
      }
      if (cnt_bwd[r] == 0) continue;
      h->prof_begin(2, st);
      rc = launch_merge_bwd(codes_c, S, lpool, gpool, Sc, h->p<int32_t>(h->o_bsrc_l) + (int64_t)r * K, h->p<int32_t>(h->o_bsrc_r) + (int64_t)r * K,
                            h->p<int32_t>(h->o_bsrc_g) + (int64_t)r * K, sorted_b ? h->p<int32_t>(h->o_order_bwd) + (int64_t)r * K : nullptr, sorted_b ? h->p<int32_t>(h->o_count_bwd) + r : nullptr,
                            h->p<double>(h->o_P) + (int64_t)r * K * 32, h->pi, h->p<double>(h->o_cnew) + (int64_t)r * K, K, cnt_bwd[r], nc, h->jc, h->skip_zero, h->skip_below * fabs(grad_elbo),
                            h->p<double>(h->o_dP) + (int64_t)r * K * 32, dpi, st);
      h->prof_end(st);
      if (rc) return rc;
    }
  }
  {
    int32_t nch = n_chunks;
    VCSMC_CUDA(cudaMemcpyAsync(h->p<int32_t>(h->o_status) + 2, &nch, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    VCSMC_CUDA(cudaStreamSynchronize(st));
  }

  // ---- dP -> (db, dQ) -> dlam, for the particles the reverse merge visited
  if (fused) {
    int32_t max_cnt = 0;
    for (int r = 0; r < N - 1; ++r) max_cnt = cnt_bwd[r] > max_cnt ? cnt_bwd[r] : max_cnt;
    if (max_cnt > 0) {
      const int64_t bx = (2 * (int64_t)max_cnt + 63) / 64;
      bwd_transition_all_kernel<<<dim3((unsigned)(bx < 4096 ? bx : 4096), N - 1), 64, 0, st>>>(
          h->Q, h->p<double>(h->o_t2), h->p<double>(h->o_dP), h->p<int32_t>(h->o_order_bwd), h->p<int32_t>(h->o_count_bwd), K, h->jc,
          h->p<double>(h->o_b_l), h->p<double>(h->o_b_r), h->lam_l, h->lam_r, h->p<double>(h->o_dQ_acc), dlam_l, dlam_r);
      VCSMC_LAUNCH_CHECK("bwd_transition_all_kernel");
    }
  } else
  for (int r = 0; r < N - 1; ++r) {
    const int64_t cnt = cnt_bwd[r];
    if (cnt == 0) continue;
    const int32_t* list = sorted_b ? h->p<int32_t>(h->o_order_bwd) + (int64_t)r * K : nullptr;
    rc = launch_transition_bwd(h->Q, h->p<double>(h->o_t2) + (int64_t)r * 2 * K, h->p<double>(h->o_dP) + (int64_t)r * K * 32, 2 * cnt, h->jc,
                               list, h->p<double>(h->o_dt), h->jc ? nullptr : h->p<double>(h->o_dQ_each), st);
    if (rc) return rc;
    bwd_branch_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(
        r, cnt, list, h->jc, h->p<double>(h->o_dt), h->p<double>(h->o_dQ_each), h->p<double>(h->o_b_l) + (int64_t)r * K,
        h->p<double>(h->o_b_r) + (int64_t)r * K, h->lam_l, h->lam_r, h->p<double>(h->o_dQ_acc), dlam_l, dlam_r);
    VCSMC_LAUNCH_CHECK("bwd_branch_kernel");
  }
  if (!h->jc) {
    column_sum_kernel<<<16, 256, 0, st>>>(h->p<double>(h->o_dQ_acc), K, 16, 16, dQ);
    VCSMC_LAUNCH_CHECK("column_sum_kernel");
  }
  return VCSMC_OK;
}

int vcsmc_sweep_profile(vcsmc_sweep_t* h, double* out_host) {
  // out_host[8] = {ms, launches} for kind 0 (forward merge / scoring), 1 (recompute merge), 2 (backward merge),
  // 3 (lazy forward: the cooperative event kernel; nested proposal: the look-ahead kernel); resets.
  if (!h || !out_host) return VCSMC_ERR_ARG;
  for (int i = 0; i < 8; ++i) out_host[i] = 0.0;
  for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
    VCSMC_CUDA(cudaEventSynchronize(h->ev[i + 1]));
    float ms = 0.f;
    VCSMC_CUDA(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
    const int kind = h->ev_kind[i / 2];
    out_host[2 * kind] += ms;
    out_host[2 * kind + 1] += 1.0;
  }
  h->ev_used = 0;
  h->ev_kind.clear();
  return VCSMC_OK;
}

void* vcsmc_sweep_output(vcsmc_sweep_t* h, const char* name) {
  if (!h || !name) return nullptr;
  struct { const char* n; int64_t off; } tab[] = {
      {"elbo", h->o_elbo}, {"log_weights", h->o_lw}, {"log_likelihood", h->o_LL}, {"log_likelihood_tilde", h->o_lltilde},
      {"log_likelihood_R", h->o_llR}, {"left_branches", h->o_b_l}, {"right_branches", h->o_b_r}, {"v_minus", h->o_vminus},
      {"ancestors", h->o_anc}, {"left_ref", h->o_lref}, {"right_ref", h->o_rref}, {"leaf_counts", h->o_nleaf},
      {"log_z", h->o_logz}, {"ess", h->o_ess}, {"status", h->o_status}, {"ell_node", h->o_ell_node},
      {"rem_positions", h->o_rempos}, {"event_timing", h->o_ev_timing}};
  if (h->M > 0 && !strcmp(name, "choice")) return h->ws + h->o_choice;
  for (auto& t : tab)
    if (!strcmp(t.n, name)) return h->ws + t.off;
  if (h->keep && !strcmp(name, "node_coef")) return h->ws + h->o_cnew;
  return nullptr;
}

}  // extern "C"
