// The lazy forward sweep, and particle sharding across the GPUs of one NVLink domain.
//
// Same outputs as the eager forward of sweep.cu (sample_phylogenies / body_rank_update, vcsmc.py:332-451), different
// schedule.  Rank event r scores every particle WITHOUT storing its node (merge_score_kernel, score.cu); once the
// weights are known and the next event has drawn its ancestors, only the particles that were drawn at least once
// ("survivors") get their node written.  With ESS ~ 1 that is one or two nodes per rank event instead of K.
//
// Everything of a rank event that is not scoring runs in ONE cooperative kernel (lz_event_kernel) whose phases are
// separated by grid barriers instead of kernel boundaries -- a rank event used to be ~28 dependent launches, which at
// 40 ms per sweep (and K/8 particles per GPU) is where the time went.  Launch r of the kernel finishes event r-1 and
// prepares event r:
//   1. weights of event r-1 for the rank's own particles (vcsmc.py:376-395) [particle sharding: packed into the step
//      record, flag barrier over peer memory, the other ranks' chunks are read straight out of their record buffers];
//   2-4. log-sum-exp, ESS and the categorical CDF over ALL K particles in a fixed tiled order (every rank derives
//      bit-identical ancestors);
//   5. ancestors of event r for all K particles (resample, vcsmc.py:284-285), and the forest ROWS of event r-1 -- but
//      only for "live" particles, those whose normalised weight is not zero in double precision: nobody else can be
//      drawn as an ancestor or carry a gradient.  A row (node ids, leaf counts, kept positions, and the forest's
//      scalars: sum of node log-likelihoods, topology prior, v^-) is rebuilt from the ancestor's row and the particle's
//      own uniforms, so every rank rebuilds every live row locally and no row ever crosses a GPU;
//   6. survivors: the owner lists them for materialisation, rows keep their nodes alive in the garbage-collected pool,
//      nodes a rank's particles descend from but does not hold are listed for a pull; the last CTA to finish runs the
//      slot allocator;
//   7. the survivors' nodes are written (plain merge), and ONE THREAD per own particle proposes event r: the pair is
//      the top-2 of its uniforms (extend_partial_state, vcsmc.py:298-305: no sort is needed for that), two ids and two
//      counts are read from the ancestor's row, the branch lengths are drawn (:351-358), both transition matrices
//      computed (:181-184) and the particle is entered into the child-pair hash table of the scoring kernel
//      [particle sharding: flag barrier, then missing nodes are pulled out of the owners' pools over NVLink];
//   8. grouped visiting order for the scoring kernel.
// The reverse sweep is sharded by SITE on the gathered tables (sweep.cu, options site_begin/site_end).
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "merge_device.cuh"
#include "sweep_state.h"

namespace cg = cooperative_groups;

namespace vcsmc {
namespace {

constexpr int kEvThreads = 256;
constexpr int kEvWarps = kEvThreads / 32;
constexpr int kMatSptEv = 2;

struct EvArgs {
  int r;                 // launch index: finishes rank event r-1 (r > 0) and prepares rank event r (r < N-1)
  int N, S, jc, gc, rank, world, sorted, two_lists, skip_leaf_pairs, row_stride, log2T, tiles, bar_index, n_barriers_total;
  int64_t K, Kl, k0, pool_slots, fetch_cap, slot_sites;
  // model and uniforms (null uniforms: counter-based generator)
  const double* lam_l;
  const double* lam_r;
  const double* Q;
  const double* qpow;         // Q^k / k! and ||Q||_1 (expm_tq_table), general Q only
  const double* pi;
  const double* ldf;
  const float* u_pair_prev;   // [K][N-r+1] of event r-1
  const float* u_pair_cur;    // [K][N-r]   of event r
  const double* u_bl;         // row r, all K
  const double* u_br;
  const double* u_res;        // row r, all K
  const uint64_t* seed_dev;
  const uint8_t* codes;
  // [N-1][K] tables
  int32_t* anc;
  int32_t* lref;
  int32_t* rref;
  int32_t* nleaf;
  double* b_l;
  double* b_r;
  double* t2;
  double* cum_l;
  double* cum_r;
  double* lw;
  double* LL;
  double* P;
  double* ell_node;
  uint8_t* rempos_prev;       // block of event r-1
  double* stats;
  // per-particle work arrays
  const double* ell_part;     // [Kl][tiles] partial sums of event r-1's scoring
  double* pF;                 // [Kl] forest sum without the new node
  double* pT;                 // [Kl] topology prior of the forest after the merge
  int32_t* pV;                // [Kl] v^- of the forest after the merge
  double* pLLt;               // [Kl] log_likelihood_tilde
  double* pEll;               // [Kl] log-likelihood of the new node when the proposing thread scored it (two leaves)
  int32_t* pDirect;           // [Kl] 1: pEll is valid, the scoring kernels skipped the particle
  const int32_t* leaf_tab;    // site-pattern table of every leaf pair (leaf_pair_hist_kernel), or null
  int32_t* vminus;            // [K] output
  int32_t* lsrc[2];           // [Kl] child slots, by event parity
  int32_t* rsrc[2];
  // rows of live particles, by event parity, indexed by GLOBAL particle
  int32_t* row_ids[2];
  int32_t* row_cnt[2];
  double* row_F[2];
  double* row_T[2];
  int32_t* row_V[2];
  double* F0;                 // sum of the leaves' log-likelihoods
  // resampling
  double* cdf;
  double* cdf_scratch;
  long long* lw_max;          // running maximum of the event's log-weights (order-preserving integer image)
  int32_t* live;
  int32_t* surv;
  int32_t* haskid;
  // node pool
  double* pool;
  int32_t* flags;
  int32_t* loc;
  int32_t* slot_id;
  int32_t* pend;
  int32_t* mat_list;          // survivors (GLOBAL particle index) whose node this rank writes in this launch
  int32_t* mat_ls;            // their child slots
  int32_t* mat_rs;
  int32_t* fetch_e;
  int32_t* fetch_src;
  int32_t* counts;            // [0] survivors to materialise, [1] nodes to pull, [2] occupied hash slots, [3..4] last-CTA tickets
  int32_t* status;
  // grouping
  unsigned long long* gtab;
  int32_t* gcnt;
  int32_t* goff;
  int32_t* gslot;
  int32_t* grank;
  int32_t* gocc;
  int32_t* order;
  int32_t* gcount;
  // outputs of the last launch
  double* llR;
  double* ll_tilde_out;
  double* elbo;
  double* logz;
  double* ess;
  double ldf_root;
  // particle sharding
  char* rec;
  int64_t rec_stride;
  const char* peer_rec[kMaxPeers];
  int32_t* peer_sig[kMaxPeers];
  const int32_t* peer_loc[kMaxPeers];
  const double* peer_pool[kMaxPeers];
  int32_t* epoch_base;
  unsigned long long* timing;   // optional [N][16] phase time stamps of CTA 0 (option "event_timing"), else null
};

__device__ __forceinline__ void stamp(const EvArgs& a, int i) {
  if (a.timing && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.timing[a.r * 16 + i] = t;
  }
}

// hash key of a child pair ((min + 256) << 32 | (max + 256)) + 1: exactly one of the two references is a leaf (< 0)
__device__ __forceinline__ bool one_leaf(unsigned long long key) {
  key -= 1ull;
  return (key >> 32) < 256ull && (key & 0xffffffffull) >= 256ull;
}

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// true in every thread of the LAST CTA to get here (all CTAs call it once per ticket); that CTA sees what the others wrote
__device__ __forceinline__ bool last_cta(int32_t* ticket, int* s_flag) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(ticket, 1);
    const int last = (t == (int)gridDim.x - 1);
    if (last) *ticket = 0;
    __threadfence();
    *s_flag = last;
  }
  __syncthreads();
  return *s_flag != 0;
}

// Cross-GPU synchronisation without the host: after a grid barrier CTA 0 stores this rank's epoch into every peer's
// flag array (peer store over NVLink, system scope; the grid barrier ordered every thread's writes before the signalling
// threads' fence, which is cumulative); with `wait` every CTA then polls its own rank's array until all peers have
// arrived.  A timeout (~30 s) turns a lost peer into a reported error instead of a hang.
__device__ __forceinline__ void cross_sync(const EvArgs& a, int index, bool wait, cg::grid_group& grid) {
  grid.sync();
  const int epoch = *a.epoch_base + index;
  const int g = threadIdx.x;
  if (blockIdx.x == 0 && g < a.world) {
    __threadfence_system();
    volatile int32_t* out = a.peer_sig[g] + a.rank;
    *out = epoch;
  }
  if (wait) {
    if (g < a.world) {
      volatile int32_t* in = a.peer_sig[a.rank] + g;
      const long long t0 = clock64();
      while (*in - epoch < 0) {
        if (clock64() - t0 > 60000000000ll) {   // ~30 s at 1.9 GHz
          a.status[0] = VCSMC_ERR_STATE;
          break;
        }
      }
      __threadfence_system();
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// phase 1: weights of event r-1 (forest posterior from the row scalars + branch priors + v^- + weight, vcsmc.py:376-395),
// O(1) per particle; under particle sharding also the rank's chunk of the step record
// ---------------------------------------------------------------------------------------------
// doubles -> int64 with the same order (NaN aside): the running maximum of the log-weights is kept with atomicMax
__device__ __forceinline__ long long ordered_bits(double x) {
  const long long b = __double_as_longlong(x);
  return b >= 0 ? b : (b ^ 0x7fffffffffffffffll);
}
__device__ __forceinline__ double from_ordered_bits(long long b) {
  return __longlong_as_double(b >= 0 ? b : (b ^ 0x7fffffffffffffffll));
}

__device__ __forceinline__ double particle_weight(const EvArgs& a, int64_t k, double ell) {
  const int r = a.r - 1;
  const int64_t kl = k - a.k0;
  const int64_t e = (int64_t)r * a.K + k;
  a.ell_node[a.N + e] = ell;
  const double F = a.pF[kl] + ell;
  const int vm = a.pV[kl];
  const double laml = a.lam_l[r], lamr = a.lam_r[r];
  const double bl = a.b_l[e], br = a.b_r[e];
  const double cl = (r > 0 ? a.cum_l[e - a.K] : 0.0) + bl;   // quirk Q1: slot-wise, un-resampled histories
  const double cr = (r > 0 ? a.cum_r[e - a.K] : 0.0) + br;
  a.cum_l[e] = cl;
  a.cum_r[e] = cr;
  const double llog = log(laml), rlog = log(lamr);
  // quirk Q2: the CURRENT step's rate multiplies ALL earlier branches (vcsmc.py:380-383)
  const double LLr = (F + a.pT[kl]) + (-laml * cl + (double)(r + 1) * llog) + (-lamr * cr + (double)(r + 1) * rlog);
  // quirk Q3: q = 1/C(n,2) is subtracted raw (vcsmc.py:298,392)
  const int n = a.N - r;
  const double q = 1.0 / ((double)n * (double)(n - 1) / 2.0);
  const double lw = LLr - a.pLLt[kl] - (llog - laml * bl + rlog - lamr * br) + log((double)vm) - q;
  a.LL[e] = LLr;
  a.lw[e] = lw;
  a.vminus[k] = vm;
  if (a.world > 1) {   // the per-event record (field-major inside the rank's chunk)
    char* base = a.rec + ((int64_t)(a.r & 1) * a.world + a.rank) * a.rec_stride;   // (two buffers, by launch parity)
    double* d = reinterpret_cast<double*>(base);
    d[0 * a.Kl + kl] = lw;
    d[1 * a.Kl + kl] = LLr;
    d[2 * a.Kl + kl] = ell;
    d[3 * a.Kl + kl] = bl;
    d[4 * a.Kl + kl] = br;
    d[5 * a.Kl + kl] = cl;
    d[6 * a.Kl + kl] = cr;
    int32_t* qi = reinterpret_cast<int32_t*>(base + 56 * a.Kl);
    qi[0 * a.Kl + kl] = a.lref[e];
    qi[1 * a.Kl + kl] = a.rref[e];
    qi[2 * a.Kl + kl] = a.nleaf[e];
    qi[3 * a.Kl + kl] = vm;
  }
  return lw;
}

// a remote particle's record: the log-weight first (the CDF waits for it) ...
__device__ __forceinline__ double particle_unpack_lw(const EvArgs& a, int64_t k) {
  const int r = a.r - 1;
  const int g = (int)(k / a.Kl);
  const int64_t kl = k - (int64_t)g * a.Kl;
  const double* d = reinterpret_cast<const double*>(a.peer_rec[g] + ((int64_t)(a.r & 1) * a.world + g) * a.rec_stride);
  const double lw = d[0 * a.Kl + kl];
  a.lw[(int64_t)r * a.K + k] = lw;
  return lw;
}

// ... and the rest, by the CTAs the CDF's tile stages leave idle (all loads first: they cross the NVLink)
__device__ __forceinline__ void particle_unpack_rest(const EvArgs& a, int64_t k) {
  const int r = a.r - 1;
  const int g = (int)(k / a.Kl);
  const int64_t kl = k - (int64_t)g * a.Kl;
  const int64_t e = (int64_t)r * a.K + k;
  const char* base = a.peer_rec[g] + ((int64_t)(a.r & 1) * a.world + g) * a.rec_stride;
  const double* d = reinterpret_cast<const double*>(base);
  const int32_t* qi = reinterpret_cast<const int32_t*>(base + 56 * a.Kl);
  const double LL = d[1 * a.Kl + kl], ell = d[2 * a.Kl + kl], bl = d[3 * a.Kl + kl], br = d[4 * a.Kl + kl];
  const double cl = d[5 * a.Kl + kl], cr = d[6 * a.Kl + kl];
  const int lref = qi[0 * a.Kl + kl], rref = qi[1 * a.Kl + kl], nleaf = qi[2 * a.Kl + kl], vm = qi[3 * a.Kl + kl];
  a.LL[e] = LL;
  a.ell_node[a.N + e] = ell;
  a.b_l[e] = bl;
  a.b_r[e] = br;
  a.t2[2 * e] = bl;
  a.t2[2 * e + 1] = br;
  a.cum_l[e] = cl;
  a.cum_r[e] = cr;
  a.lref[e] = lref;
  a.rref[e] = rref;
  a.nleaf[e] = nleaf;
  a.vminus[k] = vm;
}

// ---------------------------------------------------------------------------------------------
// phase 5b: the forest row of a live particle of event r-1 (one warp): the ancestor's row, reordered by the ascending
// rank of the particle's pair uniforms (the order tf.nn.top_k(-z) leaves the kept subtrees in, vcsmc.py:305), plus the
// new node; and the row's scalars.  Ties (injected uniforms only) follow tf.nn.top_k's lower-index rule in both
// rankings, including its keep-one-twice / drop-another consequence.
// ---------------------------------------------------------------------------------------------
template <int NQ>
__device__ __forceinline__ void build_row(const EvArgs& a, int64_t k, float* su, int lane) {
  const int r = a.r - 1;          // the event whose row this is
  const int n = a.N - r, N = a.N;  // roots before the merge
  const bool first = (r == 0);
  const int pb = r & 1, pa = pb ^ 1;
  const int64_t anc = first ? k : a.anc[(int64_t)r * a.K + k];
  const int32_t* src_ids = a.row_ids[pa] + anc * N;
  const int32_t* src_cnt = a.row_cnt[pa] + anc * N;
  int32_t* in_ = a.row_ids[pb] + k * N;
  int32_t* cn = a.row_cnt[pb] + k * N;
  uint8_t* rp = a.rempos_prev + k * (int64_t)(n - 2);
  int c0 = 0, c1 = 0;
  if (!a.u_pair_prev) {
    // counter-based uniforms: u_i = ((16 random bits) << 8 | i) 2^-24 (philox_step_kernel), so the integer key orders
    // like u, is tie-free, and carries the element's index: ONE warp bitonic sort of the keys gives the kept order
    // (ascending) and the merged pair (the two largest).  Lane b draws Philox block b (4 uniforms).
    constexpr int NB = NQ >= 4 ? NQ / 4 : 1;
    uint32_t blk[NB][4];
    const uint64_t seed = *a.seed_dev;
#pragma unroll
    for (int t = 0; t < NB; ++t) {
      blk[t][0] = (uint32_t)k;
      blk[t][1] = (uint32_t)((uint64_t)k >> 32) | ((uint32_t)r << 8);
      blk[t][2] = 2u;
      blk[t][3] = (uint32_t)(lane + 32 * t);
      philox4x32_10(blk[t], (uint32_t)seed, (uint32_t)(seed >> 32));
    }
    uint32_t key[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int i = lane + 32 * q;
      const int src = (i >> 2) & 31;
      uint32_t w = 0u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t x = __shfl_sync(0xffffffffu, blk[q >> 2][c], src);   // block (i >> 2) lives in lane src, slot q >> 2
        if (c == (i & 3)) w = x;
      }
      key[q] = i < n ? ((w >> 8) & 0xFFFF00u) | (uint32_t)(i & 0xFF) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int k2 = 2; k2 <= 32 * NQ; k2 <<= 1) {
#pragma unroll
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        if (j >= 32) {
          const int dq = j >> 5;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            if ((q & dq) == 0) {
              const bool asc = (((lane + 32 * q) & k2) == 0);
              const uint32_t lo = min(key[q], key[q | dq]), hi = max(key[q], key[q | dq]);
              key[q] = asc ? lo : hi;
              key[q | dq] = asc ? hi : lo;
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, key[q], j);
            const bool asc = (((lane + 32 * q) & k2) == 0);
            const bool lower = ((lane & j) == 0);
            key[q] = (asc == lower) ? min(key[q], other) : max(key[q], other);
          }
        }
      }
    }
    int i0 = -1, i1 = -1;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int pos = lane + 32 * q;
      if (pos >= n) continue;
      const int i = (int)(key[q] & 0xFFu);
      if (pos < n - 2) {
        rp[pos] = (uint8_t)i;
        in_[pos] = first ? i : src_ids[i];
        cn[pos] = first ? 1 : src_cnt[i];
      } else if (pos == n - 1) {
        i0 = i;
      } else {
        i1 = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      i0 = max(i0, __shfl_xor_sync(0xffffffffu, i0, o));
      i1 = max(i1, __shfl_xor_sync(0xffffffffu, i1, o));
    }
    c0 = i0;
    c1 = i1;
  } else {
    const float* u = a.u_pair_prev + k * n;
    for (int i = lane; i < n; i += 32) su[i] = u[i];
    __syncwarp();
    rank_pairs_warp(su, n, lane, c0, c1, [&](int pos, int i) {
      rp[pos] = (uint8_t)i;
      in_[pos] = first ? i : src_ids[i];
      cn[pos] = first ? 1 : src_cnt[i];
    });
    __syncwarp();
  }
  if (lane == 0) {
    in_[n - 2] = (int32_t)(N + (int64_t)r * a.K + k);
    cn[n - 2] = (first ? 1 : src_cnt[c0]) + (first ? 1 : src_cnt[c1]);
  }
  __syncwarp();
  // the row's scalars: sum of node log-likelihoods (compute_forest_posterior, vcsmc.py:238-242), topology prior (:243),
  // v^- (:247-252)
  double fs = 0.0, ts = 0.0;
  int vs = 0;
  for (int pos = lane; pos < n - 1; pos += 32) {
    const int c = cn[pos];
    fs += a.ell_node[in_[pos]];
    ts -= a.ldf[2 * max(c, 2) - 3];
    vs += c - (c == 1);
  }
  fs = warp_sum(fs);
  ts = warp_sum(ts);
  vs = warp_sum_int(vs);
  if (lane == 0) {
    a.row_F[pb][k] = fs;
    a.row_T[pb][k] = ts;
    a.row_V[pb][k] = vs;
  }
}

// ---------------------------------------------------------------------------------------------
// phase 6: a survivor of event r-1 (one warp): keep its row's nodes alive, list what has to be written or pulled
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void survivor_row(const EvArgs& a, int64_t k, int lane) {
  const int r = a.r - 1;
  const int N = a.N, m = N - r - 1;   // roots after the merge
  const int g = (int)(k / a.Kl);
  const bool own = (g == a.rank);
  const bool kid = a.haskid[k] != 0;
  if (!own && !kid) return;
  const int32_t* ids = a.row_ids[r & 1] + k * N;
  const int64_t newest = (int64_t)r * a.K + k;
  // the survivor's own node: the owner writes it; a rank whose particles descend from a REMOTE survivor writes its own
  // copy when it holds both children (the rule with ESS ~ 1: every rank follows the same one or two lineages), so that
  // nothing has to cross the NVLink and nobody waits for the owner; otherwise the node is pulled from the owner
  int ls = 0, rs = 0;
  bool compute = own;
  if (own) {
    const int64_t kl = k - a.k0;
    ls = a.lsrc[r & 1][kl];
    rs = a.rsrc[r & 1][kl];
  } else {
    const int lid = a.lref[newest], rid = a.rref[newest];
    compute = true;
    if (lid < N) {
      ls = -(lid + 1);
    } else {
      ls = a.loc[lid - N];
      compute = compute && (!a.gc || (ls >= 0 && a.slot_id[ls] == lid - N));
    }
    if (rid < N) {
      rs = -(rid + 1);
    } else {
      rs = a.loc[rid - N];
      compute = compute && (!a.gc || (rs >= 0 && a.slot_id[rs] == rid - N));
    }
  }
  for (int pos = lane; pos < m; pos += 32) {
    const int id = ids[pos];
    if (id < N) continue;
    const int e = id - N;
    bool have;
    if (e == newest) {
      have = compute;                    // gets its slot from the allocator through the survivor list
    } else {
      const int s = a.loc[e];
      have = !a.gc || (s >= 0 && a.slot_id[s] == e);
      if (have && a.gc) a.flags[s] = 1;
    }
    if (!have && kid && atomicExch(a.pend + e, a.r) != a.r) {   // first claim of this node in this rank event
      const int p = atomicAdd(a.counts + 1, 1);
      if (p < a.fetch_cap) {
        a.fetch_e[p] = e;
        a.fetch_src[p] = g;
      } else {
        a.status[0] = VCSMC_ERR_POOL;
      }
    }
  }
  if (compute) {
    if (!own && lane < 2) {
      // the remote survivor's transition matrices, from its gathered branch lengths (vcsmc.py:181-184)
      const double ti = a.t2[2 * newest + lane];
      M4 X;
      if (a.jc) {
        const double o = -0.25 * expm1(-ti);
        const double d = 0.25 + 0.75 * exp(-ti);
#pragma unroll
        for (int q = 0; q < 16; ++q) X.a[q] = (q % 5 == 0) ? d : o;
      } else {
        X = m4_expm_tq(a.qpow, ti);
      }
      double* Pout = a.P + newest * 32 + lane * 16;
#pragma unroll
      for (int q = 0; q < 16; ++q) Pout[q] = X.a[q];
    }
    if (lane == 0) {
      const int j = atomicAdd(a.counts, 1);
      a.mat_list[j] = (int32_t)k;
      a.mat_ls[j] = ls;
      a.mat_rs[j] = rs;
      if (a.gc) {   // the children the node is about to be computed from
        if (ls >= 0) a.flags[ls] = 1;
        if (rs >= 0) a.flags[rs] = 1;
      }
    }
  }
}

// Free slots (flag == 0), lowest first, go to the survivors to materialise and then to the nodes to pull.
// One CTA (the last one to finish phase 6), fixed order; each warp owns a contiguous segment of the flag array.
__device__ void allocate_slots(const EvArgs& a, int64_t* warp_off) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  volatile int32_t* counts = a.counts;
  const int64_t n_mat = counts[0];
  const int64_t n_fetch = counts[1] < a.fetch_cap ? counts[1] : a.fetch_cap;
  const int64_t Kn = n_mat + n_fetch;
  if (Kn == 0) return;
  const int64_t e_base_prev = (int64_t)(a.r - 1) * a.K;   // (the survivor list holds global particle indices)
  // every live slot lies below the running peak, so the Kn lowest free slots are below peak + Kn
  int64_t P = a.pool_slots;
  const int64_t peak = a.status[1];
  if (peak + Kn < P) P = peak + Kn;
  const int64_t seg = ((P + kEvWarps - 1) / kEvWarps + 31) / 32 * 32;
  const int64_t b = min((int64_t)wid * seg, P), e = min(b + seg, P);
  volatile const int32_t* flags = a.flags;
  int cnt = 0;
  for (int64_t i0 = b + lane; i0 < e; i0 += 128) {
    int f[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) f[q] = (i0 + 32 * q < e) ? flags[i0 + 32 * q] : 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) cnt += f[q] == 0;
  }
  cnt = warp_sum_int(cnt);
  if (lane == 0) warp_off[wid + 1] = cnt;
  __syncthreads();
  if (tid == 0) {
    int64_t t = 0;
    warp_off[0] = 0;
    for (int i = 1; i <= kEvWarps; ++i) {
      t += warp_off[i];
      warp_off[i] = t;
    }
    if (t < Kn) {
      a.status[0] = VCSMC_ERR_POOL;
      for (int64_t j = t; j < Kn; ++j) a.loc[j < n_mat ? e_base_prev + a.mat_list[j] : a.fetch_e[j - n_mat]] = -1;
    }
  }
  __syncthreads();
  int64_t j = warp_off[wid];
  int top = 0;
  for (int64_t i0 = b; i0 < e && j < Kn; i0 += 128) {
    int f[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) f[q] = (i0 + 32 * q + lane < e) ? flags[i0 + 32 * q + lane] : 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t i = i0 + 32 * q + lane;
      const bool is_free = f[q] == 0;
      const unsigned mm = __ballot_sync(0xffffffffu, is_free);
      const int64_t mine = j + __popc(mm & ((1u << lane) - 1));
      if (is_free && mine < Kn) {
        const int64_t node = mine < n_mat ? e_base_prev + ((volatile int32_t*)a.mat_list)[mine] : ((volatile int32_t*)a.fetch_e)[mine - n_mat];
        a.loc[node] = (int32_t)i;
        a.slot_id[i] = (int32_t)node;
        top = (int)i + 1;
      }
      j += __popc(mm);
    }
  }
  if (top) atomicMax(a.status + 1, top);
}

// ---------------------------------------------------------------------------------------------
// The part of the proposal of event r that does not depend on the resampling: which two positions of the forest row
// merge (extend_partial_state, vcsmc.py:298-305 -- the top-2 of the particle's uniforms), the branch lengths
// (vcsmc.py:351-358) and both transition matrices (vcsmc.py:181-184).  One thread per own particle; run by the CTAs the
// CDF's tile stages leave idle (the FP64 work of the two matrix exponentials disappears behind the scan).  The pick
// travels to propose_particle in pDirect[kl] (free between the weights of event r-1 and the proposal of event r).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void predraw_pick(const EvArgs& a, int64_t kl) {
  const int r = a.r, n = a.N - r;
  const int64_t k = a.k0 + kl;
  int c0 = 0, c1 = 1;
  bool straddle = false;   // the 2nd and 3rd largest uniforms tie: tf.nn.top_k keeps a merged subtree and drops another
  if (!a.u_pair_cur) {
    uint32_t best = 0u, second = 0u;
    const uint64_t seed = *a.seed_dev;
    for (int j = 0; j < n; j += 4) {
      uint32_t p[4] = {(uint32_t)k, (uint32_t)((uint64_t)k >> 32) | ((uint32_t)r << 8), 2u, (uint32_t)(j >> 2)};
      philox4x32_10(p, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (j + q < n) {
          const uint32_t key = (((p[q] >> 8) & 0xFFFF00u) | (uint32_t)((j + q) & 0xFF)) + 1u;   // > 0, tie-free
          if (key > best) {
            second = best;
            best = key;
          } else if (key > second) {
            second = key;
          }
        }
      }
    }
    c0 = (int)((best - 1u) & 0xFFu);
    c1 = (int)((second - 1u) & 0xFFu);
  } else {
    const float* u = a.u_pair_cur + k * n;
    float b0 = -1.f, b1 = -1.f, b2 = -1.f;   // uniforms are >= 0
    int i0 = 0, i1 = 0;
    for (int i = 0; i < n; ++i) {
      const float v = u[i];
      if (v > b0) {
        b2 = b1;
        b1 = b0; i1 = i0;
        b0 = v; i0 = i;
      } else if (v > b1) {
        b2 = b1;
        b1 = v; i1 = i;
      } else if (v > b2) {
        b2 = v;
      }
    }
    c0 = i0;
    c1 = i1;
    straddle = (n > 2) && (b1 == b2);
  }
  a.pDirect[kl] = c0 | (c1 << 8) | ((int)straddle << 16);
}

// Branch lengths (vcsmc.py:351-358) and both transition matrices (vcsmc.py:181-184) of one own particle.  `qpow` is
// the CTA's shared-memory copy of the table of m4_expm_tq (a thread per particle reading it from global memory spent
// its time in the load/store unit: 448 loads each).
__device__ __forceinline__ void predraw_transitions(const EvArgs& a, int64_t kl, const double* qpow) {
  const int r = a.r;
  const int64_t K = a.K;
  const int64_t k = a.k0 + kl;
  const int64_t e = (int64_t)r * K + k;
  double ubl, ubr;
  if (a.u_bl) {
    ubl = a.u_bl[k];
    ubr = a.u_br[k];
  } else {
    uint32_t c[4] = {(uint32_t)k, (uint32_t)((uint64_t)k >> 32) | ((uint32_t)r << 8), 0u, 0u};
    const uint64_t seed = *a.seed_dev;
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double tiny = 2.2250738585072014e-308;
    ubl = fmax(u64_to_unit_f64(c[0], c[1]), tiny);   // tfp Exponential: U in [tiny, 1)
    ubr = fmax(u64_to_unit_f64(c[2], c[3]), tiny);
  }
  const double bl = -log(ubl) / a.lam_l[r];
  const double br = -log(ubr) / a.lam_r[r];
  a.b_l[e] = bl;
  a.b_r[e] = br;
  a.t2[2 * e] = bl;
  a.t2[2 * e + 1] = br;
  double* Pout = a.P + e * 32;   // (256-byte aligned: whole 32-byte sectors per store)
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const double ti = side ? br : bl;
    M4 X;
    if (a.jc) {
      const double o = -0.25 * expm1(-ti);
      const double d = 0.25 + 0.75 * exp(-ti);
#pragma unroll
      for (int q = 0; q < 16; ++q) X.a[q] = (q % 5 == 0) ? d : o;
    } else {
      X = m4_expm_tq(qpow, ti);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      d4 row;
#pragma unroll
      for (int c = 0; c < 4; ++c) row.v[c] = X.a[q * 4 + c];
      st_site(Pout + side * 16 + q * 4, row);
    }
  }
}

// own particles [Kl * part / parts, Kl * (part + 1) / parts) of this CTA
__device__ __forceinline__ void predraw_share(const EvArgs& a, int part, int parts, int what, const double* qpow) {
  const int64_t lo = a.Kl * part / parts, hi = a.Kl * (part + 1) / parts;
  for (int64_t kl = lo + threadIdx.x; kl < hi; kl += kEvThreads) {
    if (what & 1) predraw_pick(a, kl);
    if (what & 2) predraw_transitions(a, kl, qpow);
  }
}

// ---------------------------------------------------------------------------------------------
// phase 7b: proposal of event r for one own particle (one thread)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void propose_particle(const EvArgs& a, int64_t kl, bool on, int lane) {
  const int r = a.r, N = a.N, n = N - r;
  const int64_t K = a.K;
  const int64_t k = a.k0 + kl;
  const bool first = (r == 0);
  int ls = 0, rs = 0;
  if (on) {
    const int64_t anc = first ? k : a.anc[(int64_t)r * K + k];
    const int pa = (r & 1) ^ 1;
    const int32_t* row_ids = a.row_ids[pa] + anc * N;
    const int32_t* row_cnt = a.row_cnt[pa] + anc * N;
    const int pick = a.pDirect[kl];   // predraw_particle: the two positions that merge; whether the 2nd and 3rd largest uniforms tie
    const int c0 = pick & 0xFF, c1 = (pick >> 8) & 0xFF;
    const bool straddle = (pick >> 16) & 1;
    const int lid = first ? c0 : row_ids[c0];
    const int rid = first ? c1 : row_ids[c1];
    const int cl = first ? 1 : row_cnt[c0];
    const int cr = first ? 1 : row_cnt[c1];
    const int nl = cl + cr;
    const int64_t e = (int64_t)r * K + k;
    a.nleaf[e] = nl;
    a.lref[e] = lid;
    a.rref[e] = rid;
    a.pLLt[kl] = first ? log(1.0 / (double)K) : a.LL[(int64_t)(r - 1) * K + anc];
    if (!straddle) {
      const double F_anc = first ? *a.F0 : a.row_F[pa][anc];
      const double T_anc = first ? 0.0 : a.row_T[pa][anc];
      const int V_anc = first ? 0 : a.row_V[pa][anc];
      a.pF[kl] = F_anc - a.ell_node[lid] - a.ell_node[rid];
      a.pT[kl] = T_anc + a.ldf[2 * max(cl, 2) - 3] + a.ldf[2 * max(cr, 2) - 3] - a.ldf[2 * max(nl, 2) - 3];
      a.pV[kl] = V_anc - (cl - (cl == 1)) - (cr - (cr == 1)) + nl;
    } else {
      // the kept set is "ascending rank < n-2" with tf.nn.top_k's tie rule, not "all but the pair": sum over it
      const float* u = a.u_pair_cur + k * n;
      double fs = 0.0, ts = 0.0;
      int vs = 0;
      for (int i = 0; i < n; ++i) {
        const float ui = u[i];
        int ra = 0;
        for (int j = 0; j < n; ++j) ra += (u[j] < ui) || (u[j] == ui && j < i);
        if (ra < n - 2) {
          const int c = first ? 1 : row_cnt[i];
          fs += a.ell_node[first ? i : row_ids[i]];
          ts -= a.ldf[2 * max(c, 2) - 3];
          vs += c - (c == 1);
        }
      }
      a.pF[kl] = fs;
      a.pT[kl] = ts - a.ldf[2 * max(nl, 2) - 3];
      a.pV[kl] = vs + nl;
    }
    // child slots (the allocator / earlier pulls have placed every node of the ancestor's row)
    ls = lid < N ? -(lid + 1) : a.loc[lid - N];
    rs = rid < N ? -(rid + 1) : a.loc[rid - N];
    a.lsrc[r & 1][kl] = ls;
    a.rsrc[r & 1][kl] = rs;
  }
  // ---- two leaves: the site likelihood depends on the site only through the two state masks, so
  //   sum_s log x[s] = sum over patterns of count(c_a, c_b) log x(c_a, c_b),  x = sum_{j in c_a, m in c_b} M[j][m],
  //   M[j][m] = sum_i pi_i P_a[j][i] P_b[m][i]
  // with the pattern counts of the leaf pair tabulated once per sweep (site-pattern compression, applied per cherry)
  if (on) {
    const bool cherry = a.leaf_tab != nullptr && ls < 0 && rs < 0;
    a.pDirect[kl] = cherry;
    if (cherry) {
      const int la = -ls - 1, lb = -rs - 1;
      const bool sw = la > lb;
      const int lo = sw ? lb : la, hi = sw ? la : lb;
      const int32_t* tab = a.leaf_tab + ((int64_t)lo * (2 * N - lo - 1) / 2 + (hi - lo - 1)) * kLeafPairInts;
      const double* Pa = a.P + ((int64_t)r * K + k) * 32 + (sw ? 16 : 0);   // rows of M follow the LOWER leaf
      const double* Pb = a.P + ((int64_t)r * K + k) * 32 + (sw ? 0 : 16);
      double pi[4], M[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) pi[i] = a.pi[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const d4 ra = ld_site(Pa + 4 * j);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const d4 rb = ld_site(Pb + 4 * m);
          double v = pi[0] * ra.v[0] * rb.v[0];
#pragma unroll
          for (int i = 1; i < 4; ++i) v = fma(pi[i] * ra.v[i], rb.v[i], v);
          M[j * 4 + m] = v;
        }
      }
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int c = tab[q];
        if (c != 0) acc = fma((double)c, log(M[q]), acc);
      }
      const int n_amb = tab[16];
      for (int t = 0; t < n_amb; ++t) {   // gaps and other ambiguity codes: the entries of M the two masks cover
        const int bin = tab[17 + 2 * t], c = tab[18 + 2 * t];
        const int ca = bin >> 4, cb = bin & 15;
        double x = 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int mm = 0; mm < 4; ++mm)
            if ((ca >> jj & 1) && (cb >> mm & 1)) x += M[jj * 4 + mm];
        acc = fma((double)c, log(x), acc);
      }
      a.pEll[kl] = acc;
    }
  }
  if (!a.sorted) return;
  // ---- child-pair hash table of the scoring kernel: particles with the same (unordered) pair become adjacent
  bool ins = on && !(a.skip_leaf_pairs && ls < 0 && rs < 0);   // two leaves: scored from site patterns, not listed
  int slot = -1 - lane;  // lanes that insert nothing never match anybody
  if (ins) {
    const unsigned long long key = (((unsigned long long)(unsigned)(min(ls, rs) + 256)) << 32 | (unsigned)(max(ls, rs) + 256)) + 1ull;
    const unsigned mask = (1u << a.log2T) - 1u;
    unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> (64 - a.log2T));
    while (true) {
      const unsigned long long old = atomicCAS(a.gtab + h, 0ull, key);
      if (old == 0ull) {
        a.gocc[atomicAdd(a.counts + 2, 1)] = (int32_t)h;   // first particle of this pair: the slot is now occupied
        break;
      }
      if (old == key) break;
      h = (h + 1) & mask;
    }
    slot = (int)h;
  }
  const unsigned peers = __match_any_sync(0xffffffffu, slot);
  if (ins) {
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(a.gcnt + slot, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    a.gslot[kl] = slot;
    a.grank[kl] = base + __popc(peers & ((1u << lane) - 1));
  } else if (on) {
    a.gslot[kl] = -1;
  }
}

template <int NQ>
__global__ void __launch_bounds__(kEvThreads, 2) lz_event_kernel(const EvArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float su_all[];
  __shared__ double sm[8];
  __shared__ double wsum[8];
  __shared__ int64_t warp_off[kEvWarps + 1];
  __shared__ int s_flag;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int r = a.r, N = a.N;
  const int64_t K = a.K, Kl = a.Kl, k0 = a.k0;
  const int nb = (int)((K + kCdfTile - 1) / kCdfTile);
  const int64_t gthreads = (int64_t)gridDim.x * kEvThreads, gtid = (int64_t)blockIdx.x * kEvThreads + tid;
  const int64_t gwarps = (int64_t)gridDim.x * kEvWarps, gwid = (int64_t)blockIdx.x * kEvWarps + wid;
  float* su = su_all + wid * a.row_stride;
  __shared__ __align__(16) double s_qpow[kExpmTableDoubles];   // the table of m4_expm_tq (general Q)
  if (!a.jc)
    for (int i = tid; i < kExpmTableDoubles; i += kEvThreads) s_qpow[i] = a.qpow[i];
  __syncthreads();

  if (r > 0) {
    const int64_t e_row = (int64_t)(r - 1) * K;
    double* lw = a.lw + e_row;
    double *pw = a.cdf_scratch + 2 * nb, *pq = pw + nb;
    // ---- phase 1: weights of event r-1 for the own particles (one thread each; the few partial log-likelihoods of a
    // particle are summed in a fixed order), and the maximum log-weight (exact in any order: an atomic max on the
    // order-preserving integer image of the doubles)
    stamp(a, 0);
    {
      double m = -INFINITY;
      for (int64_t kl = gtid; kl < Kl; kl += gthreads) {
        double ell = 0.0;
        if (a.pDirect[kl]) ell = a.pEll[kl];
        else
          for (int t = 0; t < a.tiles; ++t) ell += a.ell_part[kl * a.tiles + t];
        m = fmax(m, particle_weight(a, k0 + kl, ell));
      }
      if (a.world == 1) {
        m = warp_max(m);
        if (lane == 0 && m > -INFINITY) atomicMax(a.lw_max, ordered_bits(m));
      }
    }
    // scratch of the coming phases: survivor / child flags, slot flags below the running peak, the hash slots the
    // previous event occupied
    for (int64_t i = gtid; i < K; i += gthreads) {
      a.surv[i] = 0;
      a.haskid[i] = 0;
    }
    if (a.gc) {
      const int64_t peak = a.status[1];
      for (int64_t i = gtid; i < peak; i += gthreads) a.flags[i] = 0;
    }
    if (a.sorted) {
      const int n_occ = a.counts[2];
      for (int64_t i = gtid; i < n_occ; i += gthreads) {
        const int s = a.gocc[i];
        a.gtab[s] = 0ull;
        a.gcnt[s] = 0;
      }
    }
    if (a.world > 1) {
      // the other ranks' chunks of the step record, straight out of their buffers; the own weights join the maximum here
      stamp(a, 12);
      cross_sync(a, a.bar_index + 1, true, grid);
      stamp(a, 13);
      double m = -INFINITY;
      for (int64_t k = gtid; k < K; k += gthreads) m = fmax(m, (k >= k0 && k < k0 + Kl) ? lw[k] : particle_unpack_lw(a, k));
      m = warp_max(m);
      if (lane == 0 && m > -INFINITY) atomicMax(a.lw_max, ordered_bits(m));
      stamp(a, 14);
    }
    grid.sync();
    stamp(a, 1);
    const double M = from_ordered_bits(*(volatile long long*)a.lw_max);
    stamp(a, 2);
    // ---- phases 2-3: weights exp(lw - max) + live flags, CDF, log-sum-exp (resample, vcsmc.py:284-285)
    if (gtid == 0) {
      a.counts[0] = 0;
      a.counts[1] = 0;
      a.counts[2] = 0;
      a.gcount[0] = 0;
      a.gcount[1] = 0;
    }
    for (int vb = blockIdx.x; vb < nb; vb += gridDim.x) cdf_stage_weights(vb, lw, K, M, a.cdf, pw, pq, a.live, sm);
    {
      // the CTAs that have no CDF tile: the rest of the other ranks' records (needed from phase 5 on), and the part of
      // event r's proposal that does not depend on the resampling
      const int first = nb < (int)gridDim.x ? nb : 0, n_cta = (int)gridDim.x - first;
      if ((int)blockIdx.x >= first) {
        if (a.world > 1)
          for (int64_t k = (int64_t)(blockIdx.x - first) * kEvThreads + tid; k < K; k += (int64_t)n_cta * kEvThreads)
            if (!(k >= k0 && k < k0 + Kl)) particle_unpack_rest(a, k);
        if (r < N - 1) predraw_share(a, (int)blockIdx.x - first, n_cta, 1, s_qpow);
      }
    }
    grid.sync();
    stamp(a, 3);
    stamp(a, 4);
    for (int vb = blockIdx.x; vb < nb; vb += gridDim.x) cdf_stage_scan(vb, K, nb, M, pw, pq, a.cdf, a.stats + (r - 1) * 4, sm, wsum);
    if (r < N - 1) {   // (the CTAs without a CDF tile again: the transition matrices of event r)
      const int first = nb < (int)gridDim.x ? nb : 0;
      if ((int)blockIdx.x >= first) predraw_share(a, (int)blockIdx.x - first, (int)gridDim.x - first, 2, s_qpow);
    }
    grid.sync();
    if (r == N - 1) {
      // ---- last launch: ELBO (vcsmc.py:276), log_likelihood_R (vcsmc.py:254-268, incl. quirk Q4), log_likelihood_tilde
      if (gtid == 0) {
        double s = 0.0;
        const double lk = log((double)K);
        for (int q = 0; q < N - 1; ++q) {
          const double z = a.stats[q * 4] - lk;
          a.logz[q] = z;
          a.ess[q] = a.stats[q * 4 + 2];
          s += z;
        }
        a.elbo[0] = s;
      }
      for (int64_t k = gtid; k < K; k += gthreads) {
        double lp = 0.0, rp = 0.0;
        for (int q = 0; q < N - 1; ++q) {
          const double ll = log(a.lam_l[q]);
          lp += ll - a.b_l[(int64_t)q * K + k] * a.lam_l[q];
          rp += ll - a.b_r[(int64_t)q * K + k] * a.lam_r[q];  // log(LEFT param): vcsmc.py:262
        }
        a.llR[k] = a.LL[(int64_t)(N - 2) * K + k] + a.ldf_root - lp - rp;
        a.ll_tilde_out[k] = N >= 3 ? a.LL[(int64_t)(N - 3) * K + a.anc[(int64_t)(N - 2) * K + k]] : log(1.0 / (double)K);
      }
      if (a.world > 1) {
        // nobody starts the next sweep's record before everybody has read this one; then the epochs move on
        cross_sync(a, a.bar_index + 2, true, grid);
        grid.sync();
        if (gtid == 0) *a.epoch_base += a.n_barriers_total;
      }
      return;
    }
  }

  if (r == 0) predraw_share(a, (int)blockIdx.x, (int)gridDim.x, 3, s_qpow);
  // ---- phase 5: ancestors of event r (all K), rows of the live particles of event r-1 (all K)
  stamp(a, 5);
  if (gtid == 0) *a.lw_max = LLONG_MIN;   // (every CTA has read the previous maximum by now)
  {
    int32_t* anc_row = a.anc + (int64_t)r * K;
    if (r == 0) {
      for (int64_t k = gtid; k < K; k += gthreads) anc_row[k] = (int32_t)k;
      if (blockIdx.x == 0 && wid == 0) {
        double f = 0.0;
        for (int i = lane; i < N; i += 32) f += a.ell_node[i];
        f = warp_sum(f);
        if (lane == 0) *a.F0 = f;
      }
    } else {
      const double total = a.cdf[K - 1];
      // the first levels of the search run on 256 pivots in shared memory (the last entries of 256 equal blocks of the
      // CDF): same result as a binary search over the whole array, a third of the dependent global loads
      const int64_t blk = (K + 255) / 256;
      double* piv = reinterpret_cast<double*>(su_all);
      {
        const int64_t last = min((int64_t)(tid + 1) * blk, K) - 1;
        piv[tid] = (int64_t)tid * blk < K ? a.cdf[last] : INFINITY;
      }
      __syncthreads();
      for (int64_t k = gtid; k < K; k += gthreads) {
        double u;
        if (a.u_res) {
          u = a.u_res[k];
        } else {
          // the same value philox_step_kernel would write: counter (k, r, 1, 0), first two words
          uint32_t d[4] = {(uint32_t)k, (uint32_t)((uint64_t)k >> 32) | ((uint32_t)r << 8), 1u, 0u};
          const uint64_t seed = *a.seed_dev;
          philox4x32_10(d, (uint32_t)seed, (uint32_t)(seed >> 32));
          u = u64_to_unit_f64(d[0], d[1]);
        }
        const double t = u * total;
        int lo = 0, hi = 256;                // first pivot > t (the block that holds the answer)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (piv[mid] > t) hi = mid;
          else lo = mid + 1;
        }
        int idx;
        if (lo >= 256 || (int64_t)lo * blk >= K) {
          idx = (int)(K - 1);                // no entry exceeds t: clamp, like upper_bound_cdf
        } else {
          const int64_t b0 = (int64_t)lo * blk;
          const int64_t n_in = min(blk, K - b0);
          idx = (int)(b0 + upper_bound_cdf(a.cdf + b0, n_in, t));  // resample, vcsmc.py:284-285
        }
        anc_row[k] = idx;
        a.surv[idx] = 1;
        if (k >= k0 && k < k0 + Kl) a.haskid[idx] = 1;
      }
      __syncthreads();   // (the pivots share the rows' staging buffer)
      // (a warp looks at 32 flags at a time)
      for (int64_t kb = gwid * 32; kb < K; kb += gwarps * 32) {
        unsigned todo = __ballot_sync(0xffffffffu, kb + lane < K && a.live[kb + lane] != 0);
        while (todo) {
          const int b = __ffs(todo) - 1;
          todo &= todo - 1;
          build_row<NQ>(a, kb + b, su, lane);
        }
      }
    }
  }
  grid.sync();
  stamp(a, 6);
  // ---- phase 6: survivors of event r-1; the last CTA to finish allocates the slots
  if (r > 0) {
    for (int64_t kb = gwid * 32; kb < K; kb += gwarps * 32) {
      unsigned todo = __ballot_sync(0xffffffffu, kb + lane < K && a.surv[kb + lane] != 0);
      while (todo) {
        const int b = __ffs(todo) - 1;
        todo &= todo - 1;
        survivor_row(a, kb + b, lane);
      }
    }
    if (a.gc && last_cta(a.counts + 3, &s_flag)) allocate_slots(a, warp_off);
  }
  grid.sync();
  stamp(a, 7);
  // ---- phase 7: the survivors' nodes (plain merge, stored); proposal of event r for the own particles
  if (r > 0) {
    const int n_mat = a.counts[0];
    const int tiles = (a.S + kEvThreads * kMatSptEv - 1) / (kEvThreads * kMatSptEv);
    const int64_t e_base = (int64_t)(r - 1) * K;
    for (int64_t w = blockIdx.x; w < (int64_t)n_mat * tiles; w += gridDim.x) {
      const int64_t j = w / tiles;
      const int t = (int)(w - j * tiles);
      const int kg = a.mat_list[j];
      const int ds = a.loc[e_base + kg];
      if (ds < 0) continue;  // pool exhausted (reported through the status word)
      const ChildRef ra = child_ref(a.mat_ls[j], a.codes, a.S, a.pool, a.slot_sites);
      const ChildRef rb = child_ref(a.mat_rs[j], a.codes, a.S, a.pool, a.slot_sites);
      const double* Pk = a.P + (e_base + kg) * 32;
      double* out = a.pool + (int64_t)ds * a.slot_sites * 4;
      if (a.jc) {
        Trans<true> Pl, Pr;
        Pl.load(Pk);
        Pr.load(Pk + 16);
#pragma unroll
        for (int q = 0; q < kMatSptEv; ++q) {
          const int s = t * (kEvThreads * kMatSptEv) + q * kEvThreads + tid;
          if (s < a.S) {
            const d4 lp = Pl.apply(load_child(ra, s)), rp = Pr.apply(load_child(rb, s));
            d4 nw;
#pragma unroll
            for (int i = 0; i < 4; ++i) nw.v[i] = lp.v[i] * rp.v[i];
            st_site(out + (int64_t)s * 4, nw);
          }
        }
      } else {
        Trans<false> Pl, Pr;
        Pl.load(Pk);
        Pr.load(Pk + 16);
#pragma unroll
        for (int q = 0; q < kMatSptEv; ++q) {
          const int s = t * (kEvThreads * kMatSptEv) + q * kEvThreads + tid;
          if (s < a.S) {
            const d4 lp = Pl.apply(load_child(ra, s)), rp = Pr.apply(load_child(rb, s));
            d4 nw;
#pragma unroll
            for (int i = 0; i < 4; ++i) nw.v[i] = lp.v[i] * rp.v[i];
            st_site(out + (int64_t)s * 4, nw);
          }
        }
      }
    }
  }
  stamp(a, 8);
  {
    // every CTA takes an equal share of the particles (two CTAs share an SM's FP64 pipe: an idle CTA next to a full one
    // would make that SM the straggler)
    const int64_t lo = Kl * blockIdx.x / gridDim.x, hi = Kl * (blockIdx.x + 1) / gridDim.x;
    for (int64_t base = lo; base < hi; base += kEvThreads) propose_particle(a, base + tid, base + tid < hi, lane);
  }
  stamp(a, 9);
  if (a.sorted && last_cta(a.counts + 4, &s_flag)) {
    // offsets of the groups, in the order the slots were occupied: every thread takes a contiguous stretch of the
    // occupied list, one CTA-wide exclusive scan of the stretch totals (no chain of dependent atomics).  Two lists
    // share the order array: pairs of a LEAF and an internal node from the front (rows kernel), pairs of two internal
    // nodes from the back (generic kernel).
    const int n_occ = ((volatile int32_t*)a.counts)[2];
    const int per = (n_occ + kEvThreads - 1) / kEvThreads;
    const int i0 = min(tid * per, n_occ), i1 = min(i0 + per, n_occ);
    int mine[2] = {0, 0};
    for (int i = i0; i < i1; ++i) {
      const int sl = ((volatile int32_t*)a.gocc)[i];
      const bool leafy = !a.two_lists || one_leaf(((volatile unsigned long long*)a.gtab)[sl]);   // the smaller reference is a leaf
      mine[leafy ? 0 : 1] += ((volatile int32_t*)a.gcnt)[sl];
    }
    int incl[2] = {mine[0], mine[1]};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t0 = __shfl_up_sync(0xffffffffu, incl[0], o), t1 = __shfl_up_sync(0xffffffffu, incl[1], o);
      if (lane >= o) {
        incl[0] += t0;
        incl[1] += t1;
      }
    }
    int* wtot = reinterpret_cast<int*>(warp_off);
    if (lane == 31) {
      wtot[2 * wid] = incl[0];
      wtot[2 * wid + 1] = incl[1];
    }
    __syncthreads();
    int base[2] = {incl[0] - mine[0], incl[1] - mine[1]};
    for (int w2 = 0; w2 < wid; ++w2) {
      base[0] += wtot[2 * w2];
      base[1] += wtot[2 * w2 + 1];
    }
    for (int i = i0; i < i1; ++i) {
      const int sl = ((volatile int32_t*)a.gocc)[i];
      const bool leafy = !a.two_lists || one_leaf(((volatile unsigned long long*)a.gtab)[sl]);
      const int c = ((volatile int32_t*)a.gcnt)[sl];
      if (leafy) {
        a.goff[sl] = base[0];
        base[0] += c;
      } else {
        base[1] += c;
        a.goff[sl] = (int)Kl - base[1];
      }
    }
    if (tid == kEvThreads - 1) {
      a.gcount[0] = base[0];
      a.gcount[1] = base[1];
    }
  }
  // particle sharding: "my survivors' nodes are written"; only a rank that has nodes to pull waits for the others
  if (a.world > 1) cross_sync(a, a.bar_index + 2, r > 0 && ((volatile int32_t*)a.counts)[1] > 0, grid);
  else grid.sync();
  stamp(a, 10);
  // ---- phase 8: grouped visiting order; particle sharding: pull the missing nodes out of the owners' pools
  if (a.sorted) {
    for (int64_t kl = gtid; kl < Kl; kl += gthreads) {
      const int s = a.gslot[kl];
      if (s >= 0) a.order[a.goff[s] + a.grank[kl]] = (int32_t)kl;
    }
  }
  if (a.world > 1 && r > 0) {
    const int n_fetch = a.counts[1] < a.fetch_cap ? a.counts[1] : (int)a.fetch_cap;
    const int tiles = (a.S + kEvThreads * 4 - 1) / (kEvThreads * 4);
    for (int64_t w = blockIdx.x; w < (int64_t)n_fetch * tiles; w += gridDim.x) {
      const int64_t j = w / tiles;
      const int t = (int)(w - j * tiles);
      const int e = a.fetch_e[j], g = a.fetch_src[j];
      const int ds = a.loc[e];
      const int ss = a.peer_loc[g][e];
      if (ds < 0 || ss < 0) continue;
      const double* src = a.peer_pool[g] + (int64_t)ss * a.slot_sites * 4;
      double* dst = a.pool + (int64_t)ds * a.slot_sites * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int s = t * (kEvThreads * 4) + q * kEvThreads + tid;
        if (s < a.S) st_site(dst + (int64_t)s * 4, ld_site(src + (int64_t)s * 4));
      }
    }
  }
  stamp(a, 11);
}

__global__ void lz_set_seed_kernel(uint64_t* seed_dev, uint64_t seed, const double* Q, double* qpow) {
  *seed_dev = seed;
  if (Q) expm_tq_table(Q, qpow);   // once per sweep: the event kernel's transition matrices are Horner schemes on it
}

__global__ void lz_iota_kernel(int32_t* p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}

template <int NQ>
int launch_event(const EvArgs& a, int blocks, size_t smem, cudaStream_t st) {
  void* kargs[] = {(void*)&a};
  VCSMC_CUDA(cudaLaunchCooperativeKernel((const void*)lz_event_kernel<NQ>, dim3((unsigned)blocks), dim3(kEvThreads), kargs, smem, st));
  count_launch();
  if (debug_sync()) VCSMC_CUDA(cudaDeviceSynchronize());
  return VCSMC_OK;
}

template <int NQ>
int event_blocks(size_t smem, int* out) {
  int per_sm = 0, sms = 0, dev = 0;
  VCSMC_CUDA(cudaGetDevice(&dev));
  VCSMC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  VCSMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lz_event_kernel<NQ>, kEvThreads, smem));
  if (per_sm < 1) { set_error("lz_event_kernel cannot be launched cooperatively"); return VCSMC_ERR_CUDA; }
  static int want_per_sm = 0;
  if (want_per_sm == 0) {
    const char* e = getenv("VCSMC_EVENT_CTAS_PER_SM");   // tuning knob
    want_per_sm = e ? atoi(e) : 2;                      // more CTAs only make the grid barriers slower
    if (want_per_sm < 1) want_per_sm = 1;
  }
  if (per_sm > want_per_sm) per_sm = want_per_sm;
  *out = per_sm * sms;
  return VCSMC_OK;
}

// the launch sequence of one forward sweep: no host synchronisation, no host-dependent argument -- capturable
int lazy_forward_body(vcsmc_sweep* h, const uint8_t* codes, const double* lam_l, const double* lam_r, const double* Q,
                      const double* pi, cudaStream_t st) {
  const int N = h->N, S = h->S, G = h->world;
  const int64_t K = h->K, Kl = h->Kl, k0 = h->k0, E = (int64_t)(N - 1) * K;
  const bool gc = h->fwd_gc;
  int rc;
  if (G > 1 && !gc) { set_error("particle sharding runs on the garbage-collected pool"); return VCSMC_ERR_STATE; }
  if (G > 1 && !h->peer_sync) { set_error("particle sharding synchronises over peer memory (peer_sync = 1)"); return VCSMC_ERR_STATE; }
  if (!h->use_seed && h->x_look_bl != nullptr) { set_error("uniforms were set for the other proposal (nested vs plain)"); return VCSMC_ERR_STATE; }

  int32_t* status = h->p<int32_t>(h->o_status);
  VCSMC_CUDA(cudaMemsetAsync(status, 0, 8 * sizeof(int32_t), st));
  double* ell_node = h->p<double>(h->o_ell_node);
  rc = launch_leaf_ell(codes, S, N, S, pi, ell_node, st);
  if (rc) return rc;
  if (h->allreduce) {  // site sharding: every rank holds a slice of the sites
    if (h->allreduce(h->allreduce_user, ell_node, N, st)) { set_error("allreduce hook failed"); return VCSMC_ERR_CUDA; }
  }
  const int32_t* leaf_hist = nullptr;
  // site patterns of every leaf pair, once per sweep: cherries are scored from these counts by the proposing thread
  // (not under site sharding: a cherry's sum over this rank's sites would bypass the all-reduce of the scoring kernels' sums)
  if (h->leaf_patterns && !h->allreduce) {
    rc = launch_leaf_pair_hist(codes, S, N, S, h->p<int32_t>(h->o_leaf_hist), st);
    if (rc) return rc;
    leaf_hist = h->p<int32_t>(h->o_leaf_hist);
  }
  const bool sorted = h->force_sorted || use_sorted_order(Kl, S);
  const int32_t* leaf_perm = nullptr;
  const uint8_t* leaf_tstate = nullptr;
  if (sorted && h->leaf_rows) {  // every leaf's sites in state order, once per sweep: the rows kernel scores leaf + internal pairs on them
    rc = launch_leaf_sort(codes, S, N, S, h->p<int32_t>(h->o_leaf_perm), h->p<uint8_t>(h->o_leaf_tstate), st);
    if (rc) return rc;
    leaf_perm = h->p<int32_t>(h->o_leaf_perm);
    leaf_tstate = h->p<uint8_t>(h->o_leaf_tstate);
  }
  const int64_t T = group_table_entries(Kl);
  int log2T = 0;
  while (((int64_t)1 << log2T) < T) ++log2T;

  EvArgs a;
  memset(&a, 0, sizeof(a));
  a.N = N; a.S = S; a.jc = h->jc; a.gc = gc; a.rank = h->rank; a.world = G; a.sorted = sorted;
  a.two_lists = leaf_perm != nullptr; a.skip_leaf_pairs = leaf_hist != nullptr; a.row_stride = (N + 3) & ~3; a.log2T = log2T;
  a.K = K; a.Kl = Kl; a.k0 = k0; a.pool_slots = h->pool_slots; a.fetch_cap = h->fetch_cap; a.slot_sites = S;
  a.lam_l = lam_l; a.lam_r = lam_r; a.Q = Q; a.pi = pi; a.ldf = h->p<double>(h->o_ldf);
  a.qpow = h->p<double>(h->o_model) + 2 * (int64_t)N + 24;
  a.seed_dev = h->p<uint64_t>(h->o_seed_dev); a.codes = codes;
  a.anc = h->p<int32_t>(h->o_anc); a.lref = h->p<int32_t>(h->o_lref); a.rref = h->p<int32_t>(h->o_rref); a.nleaf = h->p<int32_t>(h->o_nleaf);
  a.b_l = h->p<double>(h->o_b_l); a.b_r = h->p<double>(h->o_b_r); a.t2 = h->p<double>(h->o_t2);
  a.cum_l = h->p<double>(h->o_cum_l); a.cum_r = h->p<double>(h->o_cum_r); a.lw = h->p<double>(h->o_lw); a.LL = h->p<double>(h->o_LL);
  a.P = h->p<double>(h->o_P); a.ell_node = ell_node; a.stats = h->p<double>(h->o_stats);
  a.pF = h->p<double>(h->o_pF); a.pT = h->p<double>(h->o_pT); a.pV = h->p<int32_t>(h->o_pV); a.pLLt = h->p<double>(h->o_pLLt);
  a.pEll = h->p<double>(h->o_pEll); a.pDirect = h->p<int32_t>(h->o_pDirect); a.leaf_tab = leaf_hist;
  a.vminus = h->p<int32_t>(h->o_vminus);
  a.lsrc[0] = h->p<int32_t>(h->o_lsrc); a.rsrc[0] = h->p<int32_t>(h->o_rsrc);
  a.lsrc[1] = h->p<int32_t>(h->o_lsrc2); a.rsrc[1] = h->p<int32_t>(h->o_rsrc2);
  for (int i = 0; i < 2; ++i) {
    a.row_ids[i] = h->p<int32_t>(h->o_ids[i]); a.row_cnt[i] = h->p<int32_t>(h->o_cnt[i]);
    a.row_F[i] = h->p<double>(h->o_F[i]); a.row_T[i] = h->p<double>(h->o_topo[i]); a.row_V[i] = h->p<int32_t>(h->o_vm[i]);
  }
  a.F0 = h->p<double>(h->o_F0); a.lw_max = reinterpret_cast<long long*>(h->p<double>(h->o_F0) + 1);
  a.cdf = h->p<double>(h->o_cdf); a.cdf_scratch = h->p<double>(h->o_cdf_scratch);
  a.live = h->p<int32_t>(h->o_live); a.surv = h->p<int32_t>(h->o_surv); a.haskid = h->p<int32_t>(h->o_haskid);
  a.pool = h->p<double>(h->o_pool); a.flags = h->p<int32_t>(h->o_flags); a.loc = h->p<int32_t>(h->o_loc);
  a.slot_id = gc ? h->p<int32_t>(h->o_slot_id) : nullptr; a.pend = h->p<int32_t>(h->o_pend);
  a.mat_list = h->p<int32_t>(h->o_mat_list); a.mat_ls = h->p<int32_t>(h->o_mat_ls); a.mat_rs = h->p<int32_t>(h->o_mat_rs); a.fetch_e = h->p<int32_t>(h->o_fetch_e); a.fetch_src = h->p<int32_t>(h->o_fetch_src);
  a.counts = h->p<int32_t>(h->o_counts); a.status = status;
  a.gtab = h->p<unsigned long long>(h->o_gtab); a.gcnt = h->p<int32_t>(h->o_gcnt); a.goff = h->p<int32_t>(h->o_goff);
  a.gslot = h->p<int32_t>(h->o_gslot); a.grank = h->p<int32_t>(h->o_grank); a.gocc = h->p<int32_t>(h->o_gocc);
  a.order = h->p<int32_t>(h->o_order); a.gcount = h->p<int32_t>(h->o_count);
  a.llR = h->p<double>(h->o_llR); a.ll_tilde_out = h->p<double>(h->o_lltilde); a.elbo = h->p<double>(h->o_elbo);
  a.logz = h->p<double>(h->o_logz); a.ess = h->p<double>(h->o_ess); a.ldf_root = log_double_factorial_host(2 * N - 3);
  a.rec = h->p<char>(h->o_rec); a.rec_stride = h->rec_stride; a.epoch_base = h->p<int32_t>(h->o_epoch_dev);
  for (int g = 0; g < kMaxPeers; ++g) {
    const bool on = g < G && G > 1;
    a.peer_rec[g] = on ? h->peer_ws[g] + h->o_rec : nullptr;
    a.peer_sig[g] = on ? reinterpret_cast<int32_t*>(h->peer_ws[g] + h->o_sig) : nullptr;
    a.peer_loc[g] = on ? reinterpret_cast<const int32_t*>(h->peer_ws[g] + h->o_loc) : nullptr;
    a.peer_pool[g] = on ? reinterpret_cast<const double*>(h->peer_ws[g] + h->o_pool) : nullptr;
  }
  a.n_barriers_total = 2 * N;   // every launch reserves two barrier indices
  a.timing = h->event_timing ? h->p<unsigned long long>(h->o_ev_timing) : nullptr;

  // scratch every launch relies on: the node -> slot map, the hash table, counters
  if (gc) {
    VCSMC_CUDA(cudaMemsetAsync(a.loc, 0xFF, E * sizeof(int32_t), st));
    VCSMC_CUDA(cudaMemsetAsync(a.slot_id, 0xFF, (size_t)h->pool_slots * sizeof(int32_t), st));
    VCSMC_CUDA(cudaMemsetAsync(a.flags, 0, (size_t)h->pool_slots * sizeof(int32_t), st));   // (the launches re-zero below the running peak only)
    count_launch(3);
  } else {
    lz_iota_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(a.loc, E);  // direct map: node e lives in slot e
    VCSMC_LAUNCH_CHECK("lz_iota_kernel");
  }
  if (G > 1) {
    VCSMC_CUDA(cudaMemsetAsync(a.pend, 0xFF, E * sizeof(int32_t), st));
    count_launch();
  }
  VCSMC_CUDA(cudaMemsetAsync(a.counts, 0, 8 * sizeof(int32_t), st));
  VCSMC_CUDA(cudaMemsetAsync(a.gcount, 0, 2 * sizeof(int32_t), st));
  if (sorted) {
    VCSMC_CUDA(cudaMemsetAsync(a.gtab, 0, (size_t)T * sizeof(unsigned long long), st));
    VCSMC_CUDA(cudaMemsetAsync(a.gcnt, 0, (size_t)T * sizeof(int32_t), st));
  }
  count_launch(4);

  const int NQ = N <= 32 ? 1 : N <= 64 ? 2 : N <= 128 ? 4 : 8;
  size_t smem = (size_t)kEvWarps * a.row_stride * sizeof(float);   // rows' staging buffer, also the 256 CDF pivots
  if (smem < 256 * sizeof(double)) smem = 256 * sizeof(double);
  int blocks = 0;
  rc = NQ == 1 ? event_blocks<1>(smem, &blocks) : NQ == 2 ? event_blocks<2>(smem, &blocks) : NQ == 4 ? event_blocks<4>(smem, &blocks) : event_blocks<8>(smem, &blocks);
  if (rc) return rc;
  {
    // no more CTAs than the widest phase has work for: grid barriers get cheaper with fewer CTAs
    int64_t want = (K + kEvWarps - 1) / kEvWarps;   // (a warp per particle: the row phase when every particle is live)
    if (want < 8) want = 8;
    if (want < blocks) blocks = (int)want;
  }

  int64_t pair_off = 0;
  for (int r = 0; r <= N - 1; ++r) {
    const int n = N - r;
    a.r = r;
    a.bar_index = 2 * r;
    if (h->use_seed) {
      a.u_pair_prev = nullptr; a.u_pair_cur = nullptr; a.u_bl = nullptr; a.u_br = nullptr; a.u_res = nullptr;
    } else {
      a.u_pair_prev = r > 0 ? h->x_pair + (pair_off - K * (int64_t)(n + 1)) : nullptr;
      a.u_pair_cur = r < N - 1 ? h->x_pair + pair_off : nullptr;
      a.u_bl = r < N - 1 ? h->x_bl + (int64_t)r * K : nullptr;
      a.u_br = r < N - 1 ? h->x_br + (int64_t)r * K : nullptr;
      a.u_res = r < N - 1 ? h->x_res + (int64_t)r * K : nullptr;
      pair_off += K * (int64_t)n;
    }
    a.rempos_prev = r > 0 ? h->p<uint8_t>(h->o_rempos) + h->rem_off[r - 1] : nullptr;
    h->prof_begin(3, st);
    rc = NQ == 1 ? launch_event<1>(a, blocks, smem, st) : NQ == 2 ? launch_event<2>(a, blocks, smem, st)
       : NQ == 4 ? launch_event<4>(a, blocks, smem, st) : launch_event<8>(a, blocks, smem, st);
    h->prof_end(st);
    if (rc) return rc;
    if (r == N - 1) break;

    // ---- scoring of event r: nothing stored
    const int cur = r & 1;
    const double* P = h->p<double>(h->o_P) + ((int64_t)r * K + k0) * 32;
    int tiles = 0;
    // the two scoring kernels are independent: the generic one runs on a side stream beside the rows kernel (fork / join
    // through events; inside a capture these become graph edges)
    const bool fork = leaf_perm != nullptr && h->score_streams && !h->profile;
    if (fork) {
      if (!h->side_stream) {
        VCSMC_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
        VCSMC_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        VCSMC_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
      }
      VCSMC_CUDA(cudaEventRecord(h->ev_fork, st));
      VCSMC_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    }
    h->prof_begin(0, st);
    rc = launch_merge_score(codes, S, h->p<double>(h->o_pool), S, a.lsrc[cur], a.rsrc[cur], sorted ? a.order : nullptr, P, pi, Kl,
                            sorted ? a.gcount : nullptr, S, h->jc, leaf_hist != nullptr, leaf_perm, leaf_tstate, h->p<double>(h->o_ell_part), &tiles, st,
                            fork ? h->side_stream : nullptr);
    h->prof_end(st);
    if (rc) return rc;
    if (fork) {
      VCSMC_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
      VCSMC_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
    }
    a.ell_part = h->p<double>(h->o_ell_part);
    a.tiles = tiles;
    if (h->allreduce) {
      rc = launch_ell_reduce(h->p<double>(h->o_ell_part), tiles, Kl, h->p<double>(h->o_ell_new), st);
      if (rc) return rc;
      if (h->allreduce(h->allreduce_user, h->p<double>(h->o_ell_new), Kl, st)) { set_error("allreduce hook failed"); return VCSMC_ERR_CUDA; }
      a.ell_part = h->p<double>(h->o_ell_new);
      a.tiles = 1;
    }
  }
  return VCSMC_OK;
}
}  // namespace

int sweep_forward_lazy(vcsmc_sweep* h, const uint8_t* codes, const double* lam_l_in, const double* lam_r_in, const double* Q_in,
                       const double* pi_in, cudaStream_t st) {
  const int N = h->N;
  if (!h->ldf_ready) {  // log-double-factorial table: once per sweep object
    std::vector<double> ldf(2 * N + 4, 0.0);
    for (int m = 0; m < 2 * N + 4; ++m) ldf[m] = log_double_factorial_host(m);
    VCSMC_CUDA(cudaMemcpy(h->p<double>(h->o_ldf), ldf.data(), ldf.size() * sizeof(double), cudaMemcpyHostToDevice));
    // kept positions of particles that are never rebuilt (weight exactly zero) must still be valid positions
    VCSMC_CUDA(cudaMemset(h->p<uint8_t>(h->o_rempos), 0, (size_t)(h->rem_off[N - 2] + 16)));
    h->ldf_ready = true;
  }
  // the model lives in the workspace from here on (the reverse sweep reads it too): the caller's tensors may move
  double* m = h->p<double>(h->o_model);
  double *lam_l = m, *lam_r = m + (N - 1), *Q = m + 2 * (N - 1), *pi = m + 2 * (N - 1) + 16;
  VCSMC_CUDA(cudaMemcpyAsync(lam_l, lam_l_in, (N - 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  VCSMC_CUDA(cudaMemcpyAsync(lam_r, lam_r_in, (N - 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (Q_in) VCSMC_CUDA(cudaMemcpyAsync(Q, Q_in, 16 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  VCSMC_CUDA(cudaMemcpyAsync(pi, pi_in, 4 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  lz_set_seed_kernel<<<1, 1, 0, st>>>(h->p<uint64_t>(h->o_seed_dev), h->seed, Q_in ? Q : nullptr, m + 2 * (int64_t)N + 24);
  VCSMC_LAUNCH_CHECK("lz_set_seed_kernel");
  count_launch(4);
  h->lam_l = lam_l; h->lam_r = lam_r; h->Q = Q_in ? Q : nullptr; h->pi = pi;
  const double* Qm = Q_in ? Q : nullptr;
  ++h->forwards;

  // a sequence that involves the host (collective hooks, per-launch timing) cannot be captured
  const bool capturable = h->use_graph && !h->profile && !h->allreduce;
  if (!capturable || h->forwards < 2)   // (the first forward also runs the one-time function-attribute setup)
    return lazy_forward_body(h, codes, lam_l, lam_r, Qm, pi, st);
  const void* key[6] = {codes, h->use_seed ? nullptr : (const void*)h->x_pair, h->use_seed ? nullptr : (const void*)h->x_bl,
                        h->use_seed ? nullptr : (const void*)h->x_br, h->use_seed ? nullptr : (const void*)h->x_res,
                        (const void*)(uintptr_t)(h->use_seed ? 1 : 2)};
  if (h->fwd_graph && memcmp(key, h->graph_key, sizeof(key)) != 0) {
    cudaGraphExecDestroy(h->fwd_graph);
    h->fwd_graph = nullptr;
  }
  if (!h->fwd_graph) {
    // captured on a private stream (the caller's may be the legacy default stream, which cannot capture); replayed on `st`
    if (!h->cap_stream) VCSMC_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    const uint64_t before = vcsmc_launch_count();
    VCSMC_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed));
    const int rc = lazy_forward_body(h, codes, lam_l, lam_r, Qm, pi, h->cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &graph);
    if (rc != VCSMC_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    VCSMC_CUDA(ce);
    h->graph_launches = vcsmc_launch_count() - before;
    const cudaError_t ie = cudaGraphInstantiate(&h->fwd_graph, graph, 0);
    cudaGraphDestroy(graph);
    VCSMC_CUDA(ie);
    memcpy(h->graph_key, key, sizeof(key));
  } else {
    count_launch((int)h->graph_launches);
  }
  VCSMC_CUDA(cudaGraphLaunch(h->fwd_graph, st));
  return VCSMC_OK;
}

}  // namespace vcsmc
