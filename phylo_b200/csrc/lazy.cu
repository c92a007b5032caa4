// The lazy forward sweep, and particle sharding across the GPUs of one NVLink domain.
//
// Same outputs as the eager forward of sweep.cu (sample_phylogenies / body_rank_update, vcsmc.py:332-451), different
// schedule.  Rank event r scores every particle WITHOUT storing its node (merge_score_kernel, score.cu); once the
// weights are known and the next event has drawn its ancestors, only the particles that were drawn at least once
// ("survivors") get their node written.  With ESS ~ 1 that is one or two nodes per rank event instead of K.
//
// Particle sharding (north_star): rank g owns the logical particles [g K/G, (g+1) K/G) and a private node
// pool.  Nodes are named by a global index e = r*K + k; `loc[e]` maps a node to its local pool slot (valid iff
// slot_id[loc[e]] == e), so a pool is a cache of the nodes this rank's particles reference.  Per rank event:
//   1. every rank builds the same CDF from the all-gathered weights and draws ALL K ancestors (same uniforms);
//   2. the owner of each survivor materialises it into its own pool;
//   3. every rank copies the forest ROW (node ids, leaf counts) of each of its particles' ancestors -- a peer read when
//      the ancestor lives on another GPU -- and lists the nodes of those rows it holds no copy of;
//   4. the slot allocator (free = not referenced by any surviving row of this rank) serves both lists;
//   5. barrier; missing nodes are pulled out of the owners' pools over NVLink (pull_kernel);
//   6. pair proposal, branch lengths, transition matrices, scoring, weights for the rank's own particles;
//   7. ONE all-gather of the step record (weights, likelihoods, branch lengths, child references, kept positions):
//      the tables every rank needs for the next CDF, for the outputs, and for the reverse sweep.
// The reverse sweep is sharded by SITE on the gathered tables (sweep.cu, options site_begin/site_end): the gradient is
// a sum over sites and the recompute backward needs nothing but those tables.
#include <stdlib.h>
#include <string.h>

#include "sweep_state.h"

namespace vcsmc {
namespace {

// u_res == null: the resampling uniform of logical particle k at rank event r comes straight from the counter-based
// generator (the same value philox_step_kernel would write: counter (k, r, 1, 0), first two words)
__global__ void lz_ancestors_kernel(int first, int64_t K, const double* __restrict__ cdf, const double* __restrict__ u_res,
                                    const uint64_t* __restrict__ seed_dev, int r, int32_t* __restrict__ anc, int32_t* __restrict__ surv) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  if (first) {
    anc[k] = (int32_t)k;
    return;
  }
  double u;
  if (u_res) {
    u = u_res[k];
  } else {
    uint32_t d[4] = {(uint32_t)k, (uint32_t)((uint64_t)k >> 32) | ((uint32_t)r << 8), 1u, 0u};
    const uint64_t seed = *seed_dev;
    philox4x32_10(d, (uint32_t)seed, (uint32_t)(seed >> 32));
    u = u64_to_unit_f64(d[0], d[1]);
  }
  const int idx = upper_bound_cdf(cdf, K, u * cdf[K - 1]);  // resample, vcsmc.py:284-285
  anc[k] = idx;
  surv[idx] = 1;
}

struct SurvArgs {
  int n, N, gc;
  int64_t Kl, k0;
  const int32_t* surv;
  const int32_t* ids_prev;
  const int32_t* lsrc_prev;
  const int32_t* rsrc_prev;
  const int32_t* loc;
  int32_t* flags;
  int32_t* mat_list;
  int32_t* counts;
};

// survivors of this rank: list them for materialisation; keep their forest rows and their children alive
__global__ void lz_survivors_kernel(const SurvArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = a.n + 2;
  const int64_t kl = i / w;
  const int p = (int)(i - kl * w);
  if (kl >= a.Kl || !a.surv[a.k0 + kl]) return;
  if (p < a.n) {
    if (a.gc) {
      const int id = a.ids_prev[kl * a.N + p];
      if (id >= a.N) {
        const int s = a.loc[id - a.N];
        if (s >= 0) a.flags[s] = 1;  // (the row's newest node has no slot yet)
      }
    }
  } else if (p == a.n) {
    a.mat_list[atomicAdd(a.counts, 1)] = (int32_t)kl;
    if (a.gc && a.lsrc_prev[kl] >= 0) a.flags[a.lsrc_prev[kl]] = 1;
  } else {
    if (a.gc && a.rsrc_prev[kl] >= 0) a.flags[a.rsrc_prev[kl]] = 1;
  }
}

// Free slots (flag == 0), lowest first, go to the survivors to materialise and then to the nodes to pull.
// Single CTA, fixed order; each warp owns a contiguous segment of the flag array.
__global__ void __launch_bounds__(1024) lz_alloc_kernel(const int32_t* __restrict__ flags, int64_t P, const int32_t* __restrict__ counts,
                                                        int64_t fetch_cap, const int32_t* __restrict__ mat_list, int64_t e_base_prev,
                                                        const int32_t* __restrict__ fetch_e, int32_t* __restrict__ loc,
                                                        int32_t* __restrict__ slot_id, int32_t* __restrict__ status) {
  __shared__ int64_t warp_off[33];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t n_mat = counts[0];
  const int64_t n_fetch = counts[1] < fetch_cap ? counts[1] : fetch_cap;
  const int64_t Kn = n_mat + n_fetch;
  if (Kn == 0) return;
  // every live slot lies below the running peak, so the Kn lowest free slots are below peak + Kn
  const int64_t peak = status[1];
  if (peak + Kn < P) P = peak + Kn;
  const int64_t seg = ((P + 31) / 32 + 31) / 32 * 32;
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
    if (t < Kn) {
      status[0] = VCSMC_ERR_POOL;
      for (int64_t j = t; j < Kn; ++j) loc[j < n_mat ? e_base_prev + mat_list[j] : fetch_e[j - n_mat]] = -1;
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
      const unsigned m = __ballot_sync(0xffffffffu, is_free);
      const int64_t mine = j + __popc(m & ((1u << lane) - 1));
      if (is_free && mine < Kn) {
        const int64_t node = mine < n_mat ? e_base_prev + mat_list[mine] : fetch_e[mine - n_mat];
        loc[node] = (int32_t)i;
        slot_id[i] = (int32_t)node;
        top = (int)i + 1;
      }
      j += __popc(m);
    }
  }
  if (top) atomicMax(status + 1, top);
}

// ---------------------------------------------------------------------------------------------
// Inherit + propose in one pass (one warp per own particle): copy the ancestor's forest row (peer read when it lives on
// another GPU), keep cached nodes alive / list missing ones, draw the pair and the two branch lengths, write the new
// row.  The forest's scalars (sum of node log-likelihoods F, topology prior, v^-) travel with the particle and are
// updated incrementally, so the weight step no longer walks the forest:
//     F' = F_anc - ell[left] - ell[right] (+ ell[new], added once the merge has been scored).
// resample vcsmc.py:284-289,318-325; extend_partial_state :298-305; Exponential sample :351-358; state update :361-373;
// compute_forest_posterior :231-245 and overcounting_correct :247-252 as running sums.
// ---------------------------------------------------------------------------------------------
struct ProposeArgs {
  int r, n, N, gc, rank, row_stride;
  int64_t K, Kl, k0, fetch_cap;
  const int32_t* anc;
  const int32_t* peer_ids[kMaxPeers];
  const int32_t* peer_cnt[kMaxPeers];
  const double* peer_F[kMaxPeers];
  const double* peer_topo[kMaxPeers];
  const int32_t* peer_vm[kMaxPeers];
  const int32_t* loc;
  const int32_t* slot_id;
  int32_t* flags;
  int32_t* pend;
  int32_t* fetch_e;
  int32_t* fetch_src;
  int32_t* counts;
  int32_t* status;
  const float* u_pair;   // explicit uniforms of the own particles ([Kl][n], [Kl], [Kl]) or null: counter-based generator
  const double* u_bl;
  const double* u_br;
  const uint64_t* seed_dev;
  const double* lam_l;
  const double* lam_r;
  const double* ell_node;
  const double* ldf;
  const double* LL_prev;
  int32_t* ids_new;   // [Kl][N]
  int32_t* cnt_new;
  double* F_new;      // [Kl] forest scalars after this event (F without the new node's term until the weight step)
  double* topo_new;
  int32_t* vm_new;
  int32_t* lref;      // row r of the [N-1][K] tables
  int32_t* rref;
  int32_t* nleaf;
  uint8_t* rempos;
  double* b_l;
  double* b_r;
  double* t2;
  double* ll_tilde;   // [Kl]
};

constexpr int kProposeWarps = 8;

template <int NQ>
__global__ void __launch_bounds__(kProposeWarps * 32) lz_propose_kernel(const ProposeArgs a) {
  extern __shared__ __align__(16) float su_all[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t kl = (int64_t)blockIdx.x * kProposeWarps + wid;
  if (kl >= a.Kl) return;
  const int n = a.n, N = a.N, r = a.r;
  const int64_t k = a.k0 + kl;
  const bool first = (r == 0);  // initial forest = the N leaves, one each (vcsmc.py:414-415)
  float* su = su_all + wid * a.row_stride;

  // ---- the ancestor's row
  int64_t anc = k;
  int g = a.rank;
  int64_t al = kl;
  if (!first) {
    anc = a.anc[k];
    g = (int)(anc / a.Kl);
    al = anc - (int64_t)g * a.Kl;
  }
  const int32_t* src_ids = a.peer_ids[g] + al * N;
  const int32_t* src_cnt = a.peer_cnt[g] + al * N;
  const int64_t newest = (int64_t)(r - 1) * a.K + anc;  // the ancestor's own node: materialised by its owner
  int my_id[NQ], my_cnt[NQ];
  float my_u[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int i = lane + 32 * q;
    my_id[q] = -1;
    my_cnt[q] = 0;
    my_u[q] = 0.f;
    if (i >= n) continue;
    if (first) {
      my_id[q] = i;
      my_cnt[q] = 1;
    } else {
      const int id = src_ids[i];
      my_id[q] = id;
      my_cnt[q] = src_cnt[i];
      if (id >= N) {
        const int e = id - N;
        bool have;
        if (e == newest) {
          have = (g == a.rank);  // gets its slot from the allocator, via the survivor list
        } else {
          const int s = a.loc[e];
          have = !a.gc || (s >= 0 && a.slot_id[s] == e);
          if (have && a.gc) a.flags[s] = 1;
        }
        if (!have && atomicExch(a.pend + e, r) != r) {  // first claim of this node in this rank event
          const int pos = atomicAdd(a.counts + 1, 1);
          if (pos < a.fetch_cap) {
            a.fetch_e[pos] = e;
            a.fetch_src[pos] = g;
          } else {
            a.status[0] = VCSMC_ERR_POOL;
          }
        }
      }
    }
    if (a.u_pair) {   // injected uniforms: ranked by counting below
      const float u = a.u_pair[kl * n + i];
      my_u[q] = u;
      su[i] = u;
    }
  }
  int32_t* in_ = a.ids_new + kl * N;
  int32_t* cn = a.cnt_new + kl * N;
  uint8_t* rp = a.rempos + k * (int64_t)(n - 2);
  int lid = 0, rid = 0, cl = 0, cr = 0;   // (valid in lane 0 afterwards)
  bool tie_row = false;
  double tie_F = 0.0, tie_topo = 0.0;
  int tie_vm = 0;
  if (!a.u_pair) {
    // ---- counter-based uniforms: u_i = ((16 random bits) << 8 | i) 2^-24 (philox_step_kernel), so the integer key
    // orders like u, is tie-free, and carries the element's index: ONE warp bitonic sort of the keys gives the kept
    // order (ascending) and the merged pair (the two largest).  Lane b draws Philox block b (4 uniforms).
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
    // bitonic sort of the 32*NQ keys, element index = lane + 32 q, ascending
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
    // positions n-1 / n-2 live in lanes (n-1) & 31 / (n-2) & 31 (max over the lane's slots picks the one that was set)
    const int c0 = __shfl_sync(0xffffffffu, i0, (n - 1) & 31), c1 = __shfl_sync(0xffffffffu, i1, (n - 2) & 31);
    lid = first ? c0 : src_ids[c0];
    rid = first ? c1 : src_ids[c1];
    cl = first ? 1 : src_cnt[c0];
    cr = first ? 1 : src_cnt[c1];
  } else {
  for (int i = n + lane; i < ((n + 3) & ~3); i += 32) su[i] = INFINITY;   // padding of the float4 reads below
  __syncwarp();

  // ---- ranks.  z = -log(-log u) is increasing in u, so ranking u reproduces tf.nn.top_k (smc_device.cuh).  Without
  // ties the ascending rank is #{u_j < u_i} and the descending rank its mirror image; rows with a tie (never produced
  // by the generator above, possible with injected uniforms) take the exact all-pairs routine with the reference's
  // tie rule.
  int lt[NQ], eq[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) lt[q] = eq[q] = 0;
  for (int j = 0; j < n; j += 4) {
    const float4 v = *reinterpret_cast<const float4*>(su + j);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      lt[q] += (v.x < my_u[q]) + (v.y < my_u[q]) + (v.z < my_u[q]) + (v.w < my_u[q]);
      eq[q] += (v.x == my_u[q]) + (v.y == my_u[q]) + (v.z == my_u[q]) + (v.w == my_u[q]);
    }
  }
  bool tie = false;
#pragma unroll
  for (int q = 0; q < NQ; ++q) tie = tie || (lane + 32 * q < n && eq[q] != 1);
  if (__any_sync(0xffffffffu, tie)) {
    int c0, c1;
    rank_pairs_warp(su, n, lane, c0, c1, [&](int pos, int i) {
      rp[pos] = (uint8_t)i;
      in_[pos] = first ? i : src_ids[i];
      cn[pos] = first ? 1 : src_cnt[i];
    });
    lid = first ? c0 : src_ids[c0];
    rid = first ? c1 : src_ids[c1];
    cl = first ? 1 : src_cnt[c0];
    cr = first ? 1 : src_cnt[c1];
    // tf.nn.top_k's tie rule can keep a merged subtree and drop another one (duplicate-on-tie quirk, vcsmc.py:304-305),
    // so the kept set is not "all but the pair": the forest scalars of such a row are summed over the row itself
    __syncwarp();
    double fs = 0.0, ts = 0.0;
    int vs = 0;
    for (int pos = lane; pos < n - 2; pos += 32) {
      const int c = cn[pos];
      fs += a.ell_node[in_[pos]];
      ts -= a.ldf[2 * max(c, 2) - 3];
      vs += c - (c == 1);
    }
    tie_F = warp_sum(fs);
    tie_topo = warp_sum(ts);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
    tie_vm = vs;
    tie_row = true;
  } else {
    int id0 = -1, ct0 = 0, id1 = -1, ct1 = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int i = lane + 32 * q;
      if (i >= n) continue;
      const int ra = lt[q];
      if (ra < n - 2) {
        rp[ra] = (uint8_t)i;
        in_[ra] = my_id[q];
        cn[ra] = my_cnt[q];
      } else if (ra == n - 1) {
        id0 = my_id[q];
        ct0 = my_cnt[q];
      } else {
        id1 = my_id[q];
        ct1 = my_cnt[q];
      }
    }
    const int s0 = __ffs(__ballot_sync(0xffffffffu, id0 >= 0)) - 1, s1 = __ffs(__ballot_sync(0xffffffffu, id1 >= 0)) - 1;
    lid = __shfl_sync(0xffffffffu, id0, s0);
    cl = __shfl_sync(0xffffffffu, ct0, s0);
    rid = __shfl_sync(0xffffffffu, id1, s1);
    cr = __shfl_sync(0xffffffffu, ct1, s1);
  }
  }
  // ---- the forest's scalars before the merge
  double F_anc = 0.0;
  if (first) {
    for (int i = lane; i < N; i += 32) F_anc += a.ell_node[i];
    F_anc = warp_sum(F_anc);
  }
  if (lane == 0) {
    double topo_anc = 0.0;   // -sum log (2 max(c,2) - 3)!! over the roots: 0 for N leaves
    int vm_anc = 0;          // sum (c - [c == 1]) over the roots: 0 for N leaves
    if (!first) {
      F_anc = a.peer_F[g][al];
      topo_anc = a.peer_topo[g][al];
      vm_anc = a.peer_vm[g][al];
    }
    in_[n - 2] = (int32_t)(N + (int64_t)r * a.K + k);
    const int nl = cl + cr;
    cn[n - 2] = nl;
    a.nleaf[k] = nl;
    a.lref[k] = lid;
    a.rref[k] = rid;
    double ubl, ubr;
    if (a.u_bl) {
      ubl = a.u_bl[kl];
      ubr = a.u_br[kl];
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
    a.b_l[k] = bl;
    a.b_r[k] = br;
    a.t2[2 * k] = bl;
    a.t2[2 * k + 1] = br;
    a.ll_tilde[kl] = first ? log(1.0 / (double)a.K) : a.LL_prev[anc];
    if (tie_row) {
      a.F_new[kl] = tie_F;
      a.topo_new[kl] = tie_topo - a.ldf[2 * max(nl, 2) - 3];
      a.vm_new[kl] = tie_vm + nl;
    } else {
      a.F_new[kl] = F_anc - a.ell_node[lid] - a.ell_node[rid];
      a.topo_new[kl] = topo_anc + a.ldf[2 * max(cl, 2) - 3] + a.ldf[2 * max(cr, 2) - 3] - a.ldf[2 * max(nl, 2) - 3];
      a.vm_new[kl] = vm_anc - (cl - (cl == 1)) - (cr - (cr == 1)) + nl;
    }
  }
}

// child slots of the own particles, once the allocator / the pulls have placed every node
__global__ void lz_resolve_kernel(int64_t Kl, int N, const int32_t* __restrict__ lref, const int32_t* __restrict__ rref,
                                  const int32_t* __restrict__ loc, int32_t* __restrict__ lsrc, int32_t* __restrict__ rsrc) {
  const int64_t kl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kl >= Kl) return;
  const int l = lref[kl], r = rref[kl];
  lsrc[kl] = l < N ? -(l + 1) : loc[l - N];
  rsrc[kl] = r < N ? -(r + 1) : loc[r - N];
}

struct LzWeightArgs {
  int r, n, tiles;
  int64_t Kl;
  const double* ell_part;
  double* F;            // [Kl] in: forest sum without the new node; out: with it
  const double* topo;   // [Kl]
  const int32_t* vm;    // [Kl]
  const double* lam_l;
  const double* lam_r;
  const double* b_l;    // own columns of row r
  const double* b_r;
  const double* cum_l_prev;
  const double* cum_r_prev;
  double* cum_l;
  double* cum_r;
  const double* ll_tilde;
  double* ell_new;      // ell_node + N + r*K + k0
  double* lw;
  double* LL;
  int32_t* vminus;
  double q;
};

// forest posterior from the running sums + branch priors + v^- + weight (vcsmc.py:376-395), O(1) per particle
__global__ void lz_weights_kernel(const LzWeightArgs a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.Kl) return;
  const int r = a.r;
  double ell = 0.0;
  for (int t = 0; t < a.tiles; ++t) ell += a.ell_part[k * a.tiles + t];
  a.ell_new[k] = ell;
  const double F = a.F[k] + ell;
  a.F[k] = F;
  const int vm = a.vm[k];
  const double laml = a.lam_l[r], lamr = a.lam_r[r];
  const double bl = a.b_l[k], br = a.b_r[k];
  const double cl = (r > 0 ? a.cum_l_prev[k] : 0.0) + bl;   // quirk Q1: slot-wise, un-resampled histories
  const double cr = (r > 0 ? a.cum_r_prev[k] : 0.0) + br;
  a.cum_l[k] = cl;
  a.cum_r[k] = cr;
  const double llog = log(laml), rlog = log(lamr);
  // quirk Q2: the CURRENT step's rate multiplies ALL earlier branches (vcsmc.py:380-383)
  const double LLr = (F + a.topo[k]) + (-laml * cl + (double)(r + 1) * llog) + (-lamr * cr + (double)(r + 1) * rlog);
  // quirk Q3: q = 1/C(n,2) is subtracted raw (vcsmc.py:298,392)
  const double lw = LLr - a.ll_tilde[k] - (llog - laml * bl + rlog - lamr * br) + log((double)vm) - a.q;
  a.LL[k] = LLr;
  a.lw[k] = lw;
  a.vminus[k] = vm;
}

// ---- the per-event record that is all-gathered across ranks (field-major inside a rank's chunk)
struct RecArgs {
  int n, rank;
  int64_t Kl, k0, K, stride;
  char* rec;
  const char* peer_rec[kMaxPeers];  // null: every chunk has been gathered into `rec` (collective hook); else read chunk g from peer g
  double* lw;
  double* LL;
  double* ell;  // ell_node + N + r*K
  double* b_l;
  double* b_r;
  double* cum_l;
  double* cum_r;
  double* t2;
  int32_t* lref;
  int32_t* rref;
  int32_t* nleaf;
  int32_t* vminus;
  uint8_t* rempos;
};

// Cross-GPU barrier without the host: lane g stores this rank's epoch into peer g's flag array (peer store over
// NVLink, system scope) and then spins until peer g's epoch has arrived in this rank's own array.  Every GPU runs
// its own stream, so the store a lane waits for never depends on this kernel.  A timeout (~30 s) turns a lost peer
// into a reported error instead of a hang.
struct SigArgs {
  int32_t* peer_sig[kMaxPeers];  // peer g's flag array (own array at index rank)
  int rank, world, index;        // this is the index-th barrier since the epoch base was last advanced
  const int32_t* epoch_base;     // device-resident, advanced at the end of every forward (a replayed graph stays monotonic)
  int32_t* status;
};

__global__ void lz_barrier_kernel(const SigArgs a) {
  const int g = threadIdx.x;
  if (g >= a.world) return;
  const int epoch = *a.epoch_base + a.index;
  __threadfence_system();
  volatile int32_t* out = a.peer_sig[g] + a.rank;
  *out = epoch;
  __threadfence_system();
  volatile int32_t* in = a.peer_sig[a.rank] + g;
  const long long t0 = clock64();
  while (*in - epoch < 0) {
    if (clock64() - t0 > 60000000000ll) {   // ~30 s at 1.9 GHz
      a.status[0] = VCSMC_ERR_STATE;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

__global__ void lz_advance_kernel(int32_t* epoch_base, int n) { *epoch_base += n; }
__global__ void lz_set_seed_kernel(uint64_t* seed_dev, uint64_t seed) { *seed_dev = seed; }

__global__ void lz_pack_kernel(const RecArgs a) {
  const int64_t kl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kl >= a.Kl) return;
  const int64_t k = a.k0 + kl;
  char* base = a.rec + (int64_t)a.rank * a.stride;
  double* d = reinterpret_cast<double*>(base);
  d[0 * a.Kl + kl] = a.lw[k];
  d[1 * a.Kl + kl] = a.LL[k];
  d[2 * a.Kl + kl] = a.ell[k];
  d[3 * a.Kl + kl] = a.b_l[k];
  d[4 * a.Kl + kl] = a.b_r[k];
  d[5 * a.Kl + kl] = a.cum_l[k];
  d[6 * a.Kl + kl] = a.cum_r[k];
  int32_t* q = reinterpret_cast<int32_t*>(base + 56 * a.Kl);
  q[0 * a.Kl + kl] = a.lref[k];
  q[1 * a.Kl + kl] = a.rref[k];
  q[2 * a.Kl + kl] = a.nleaf[k];
  q[3 * a.Kl + kl] = a.vminus[k];
  uint8_t* b = reinterpret_cast<uint8_t*>(base + 72 * a.Kl);
  const int m = a.n - 2;
  for (int p = 0; p < m; ++p) b[kl * m + p] = a.rempos[k * m + p];
}

__global__ void lz_unpack_kernel(const RecArgs a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.K) return;
  const int g = (int)(k / a.Kl);
  if (g == a.rank) return;
  const int64_t kl = k - (int64_t)g * a.Kl;
  const char* base = (a.peer_rec[g] ? a.peer_rec[g] : a.rec) + (int64_t)g * a.stride;
  const double* d = reinterpret_cast<const double*>(base);
  a.lw[k] = d[0 * a.Kl + kl];
  a.LL[k] = d[1 * a.Kl + kl];
  a.ell[k] = d[2 * a.Kl + kl];
  const double bl = d[3 * a.Kl + kl], br = d[4 * a.Kl + kl];
  a.b_l[k] = bl;
  a.b_r[k] = br;
  a.t2[2 * k] = bl;
  a.t2[2 * k + 1] = br;
  a.cum_l[k] = d[5 * a.Kl + kl];
  a.cum_r[k] = d[6 * a.Kl + kl];
  const int32_t* q = reinterpret_cast<const int32_t*>(base + 56 * a.Kl);
  a.lref[k] = q[0 * a.Kl + kl];
  a.rref[k] = q[1 * a.Kl + kl];
  a.nleaf[k] = q[2 * a.Kl + kl];
  a.vminus[k] = q[3 * a.Kl + kl];
  const uint8_t* b = reinterpret_cast<const uint8_t*>(base + 72 * a.Kl);
  const int m = a.n - 2;
  for (int p = 0; p < m; ++p) a.rempos[k * m + p] = b[kl * m + p];
}

__global__ void lz_lltilde_kernel(int64_t K, int N, const double* __restrict__ LL, const int32_t* __restrict__ anc,
                                  double* __restrict__ ll_tilde) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  ll_tilde[k] = N >= 3 ? LL[(int64_t)(N - 3) * K + anc[(int64_t)(N - 2) * K + k]] : log(1.0 / (double)K);
}

__global__ void lz_iota_kernel(int32_t* p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}

}  // namespace

namespace {
// all ranks' earlier work on `st` is complete (and visible to peers) before any rank's later work starts
int cross_rank_barrier(vcsmc_sweep* h, int* n_barriers, cudaStream_t st) {
  if (!h->peer_sync) {
    if (h->comm(h->comm_user, VCSMC_COMM_BARRIER, nullptr, 0, st)) { set_error("comm hook failed (barrier)"); return VCSMC_ERR_CUDA; }
    return VCSMC_OK;
  }
  SigArgs a;
  for (int g = 0; g < kMaxPeers; ++g) a.peer_sig[g] = g < h->world ? reinterpret_cast<int32_t*>(h->peer_ws[g] + h->o_sig) : nullptr;
  a.rank = h->rank; a.world = h->world; a.index = ++*n_barriers; a.epoch_base = h->p<int32_t>(h->o_epoch_dev);
  a.status = h->p<int32_t>(h->o_status);
  lz_barrier_kernel<<<1, 32, 0, st>>>(a);
  VCSMC_LAUNCH_CHECK("lz_barrier_kernel");
  return VCSMC_OK;
}
}  // namespace

namespace {
// the launch sequence of one forward sweep: no host synchronisation, no host-dependent argument -- capturable
int lazy_forward_body(vcsmc_sweep* h, const uint8_t* codes, const double* lam_l, const double* lam_r, const double* Q,
                      const double* pi, cudaStream_t st) {
  const int N = h->N, S = h->S, G = h->world;
  int n_barriers = 0;
  const int64_t K = h->K, Kl = h->Kl, k0 = h->k0, E = (int64_t)(N - 1) * K;
  const bool gc = h->fwd_gc;
  int rc;
  if (G > 1 && !gc) { set_error("particle sharding runs on the garbage-collected pool"); return VCSMC_ERR_STATE; }
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
  if (h->leaf_patterns) {  // site patterns of every leaf pair, once per sweep: cherries are scored from these counts
    rc = launch_leaf_pair_hist(codes, S, N, S, h->p<int32_t>(h->o_leaf_hist), st);
    if (rc) return rc;
    leaf_hist = h->p<int32_t>(h->o_leaf_hist);
  }
  double* pool = h->p<double>(h->o_pool);
  int32_t* flags = h->p<int32_t>(h->o_flags);
  int32_t* loc = h->p<int32_t>(h->o_loc);
  int32_t* slot_id = gc ? h->p<int32_t>(h->o_slot_id) : nullptr;
  int32_t* pend = h->p<int32_t>(h->o_pend);
  int32_t* surv = h->p<int32_t>(h->o_surv);
  int32_t* counts = h->p<int32_t>(h->o_counts);
  int32_t* mat_list = h->p<int32_t>(h->o_mat_list);
  int32_t* fetch_e = h->p<int32_t>(h->o_fetch_e);
  int32_t* fetch_src = h->p<int32_t>(h->o_fetch_src);
  int32_t* inh_ids = h->p<int32_t>(h->o_lz_ids);
  int32_t* inh_cnt = h->p<int32_t>(h->o_lz_cnt);
  int32_t* lsrc = h->p<int32_t>(h->o_lsrc);
  int32_t* rsrc = h->p<int32_t>(h->o_rsrc);
  double* ll_tilde = h->p<double>(h->o_lltilde);
  if (gc) {
    VCSMC_CUDA(cudaMemsetAsync(loc, 0xFF, E * sizeof(int32_t), st));
    VCSMC_CUDA(cudaMemsetAsync(slot_id, 0xFF, (size_t)h->pool_slots * sizeof(int32_t), st));
    count_launch(2);
  } else {
    lz_iota_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(loc, E);  // direct map: node e lives in slot e
    VCSMC_LAUNCH_CHECK("lz_iota_kernel");
  }
  if (G > 1) {
    VCSMC_CUDA(cudaMemsetAsync(pend, 0xFF, E * sizeof(int32_t), st));
    count_launch();
  }
  const bool sorted = use_sorted_order(Kl, S);
  int64_t pair_off = 0;

  for (int r = 0; r < N - 1; ++r) {
    const int n = N - r;
    const int cur = r & 1, prev = cur ^ 1;
    // ---- uniforms: pair / branch draws of the rank's own particles, resampling draws of ALL particles
    const float* u_pair;
    const double *u_bl, *u_br, *u_res_all;
    if (h->use_seed) {
      u_pair = nullptr; u_bl = nullptr; u_br = nullptr;   // drawn inside lz_propose_kernel / lz_ancestors_kernel
      u_res_all = nullptr;
    } else {
      u_pair = h->x_pair + pair_off + k0 * n; u_bl = h->x_bl + (int64_t)r * K + k0; u_br = h->x_br + (int64_t)r * K + k0;
      u_res_all = h->x_res + (int64_t)r * K;
      pair_off += K * n;
    }
    int32_t* anc_row = h->p<int32_t>(h->o_anc) + (int64_t)r * K;
    int32_t* ids_prev = h->p<int32_t>(h->o_ids[prev]);
    int32_t* ids_cur = h->p<int32_t>(h->o_ids[cur]);
    int32_t* cnt_cur = h->p<int32_t>(h->o_cnt[cur]);
    int32_t* row_lref = h->p<int32_t>(h->o_lref) + (int64_t)r * K;
    int32_t* row_rref = h->p<int32_t>(h->o_rref) + (int64_t)r * K;
    int32_t* row_nleaf = h->p<int32_t>(h->o_nleaf) + (int64_t)r * K;
    uint8_t* row_rempos = h->p<uint8_t>(h->o_rempos) + h->rem_off[r];
    double* row_b_l = h->p<double>(h->o_b_l) + (int64_t)r * K;
    double* row_b_r = h->p<double>(h->o_b_r) + (int64_t)r * K;
    double* row_t2 = h->p<double>(h->o_t2) + (int64_t)r * 2 * K;
    // inherit + propose for the rank's own particles (reads the rows and forest scalars event r-1 left behind)
    auto launch_propose = [&]() -> int {
      ProposeArgs a;
      a.r = r; a.n = n; a.N = N; a.gc = gc; a.rank = h->rank; a.row_stride = (N + 3) & ~3;
      a.K = K; a.Kl = Kl; a.k0 = k0; a.fetch_cap = h->fetch_cap; a.anc = anc_row;
      for (int g = 0; g < kMaxPeers; ++g) {
        const bool on = g < G;
        a.peer_ids[g] = on ? reinterpret_cast<const int32_t*>(h->peer_ws[g] + h->o_ids[prev]) : nullptr;
        a.peer_cnt[g] = on ? reinterpret_cast<const int32_t*>(h->peer_ws[g] + h->o_cnt[prev]) : nullptr;
        a.peer_F[g] = on ? reinterpret_cast<const double*>(h->peer_ws[g] + h->o_F[prev]) : nullptr;
        a.peer_topo[g] = on ? reinterpret_cast<const double*>(h->peer_ws[g] + h->o_topo[prev]) : nullptr;
        a.peer_vm[g] = on ? reinterpret_cast<const int32_t*>(h->peer_ws[g] + h->o_vm[prev]) : nullptr;
      }
      a.loc = loc; a.slot_id = slot_id; a.flags = flags; a.pend = pend; a.fetch_e = fetch_e; a.fetch_src = fetch_src;
      a.counts = counts; a.status = status;
      a.u_pair = u_pair; a.u_bl = u_bl; a.u_br = u_br; a.seed_dev = h->p<uint64_t>(h->o_seed_dev);
      a.lam_l = lam_l; a.lam_r = lam_r; a.ell_node = ell_node; a.ldf = h->p<double>(h->o_ldf);
      a.LL_prev = r > 0 ? h->p<double>(h->o_LL) + (int64_t)(r - 1) * K : nullptr;
      a.ids_new = ids_cur; a.cnt_new = cnt_cur;
      a.F_new = h->p<double>(h->o_F[cur]); a.topo_new = h->p<double>(h->o_topo[cur]); a.vm_new = h->p<int32_t>(h->o_vm[cur]);
      a.lref = row_lref; a.rref = row_rref; a.nleaf = row_nleaf; a.rempos = row_rempos;
      a.b_l = row_b_l; a.b_r = row_b_r; a.t2 = row_t2; a.ll_tilde = ll_tilde;
      const unsigned grid = (unsigned)((Kl + kProposeWarps - 1) / kProposeWarps);
      const size_t smem = (size_t)kProposeWarps * a.row_stride * sizeof(float);
      if (N <= 64) lz_propose_kernel<2><<<grid, kProposeWarps * 32, smem, st>>>(a);
      else lz_propose_kernel<8><<<grid, kProposeWarps * 32, smem, st>>>(a);
      VCSMC_LAUNCH_CHECK("lz_propose_kernel");
      return VCSMC_OK;
    };

    if (r > 0) {
      VCSMC_CUDA(cudaMemsetAsync(surv, 0, K * sizeof(int32_t), st));
      VCSMC_CUDA(cudaMemsetAsync(counts, 0, 8 * sizeof(int32_t), st));
      if (gc) VCSMC_CUDA(cudaMemsetAsync(flags, 0, (size_t)h->pool_slots * sizeof(int32_t), st));
      count_launch(3);
    }
    lz_ancestors_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(r == 0, K, h->p<double>(h->o_cdf), u_res_all, h->p<uint64_t>(h->o_seed_dev), r, anc_row, surv);
    VCSMC_LAUNCH_CHECK("lz_ancestors_kernel");
    if (r > 0) {
      const int64_t e_base_prev = (int64_t)(r - 1) * K + k0;
      SurvArgs sa;
      sa.n = n; sa.N = N; sa.gc = gc; sa.Kl = Kl; sa.k0 = k0; sa.surv = surv; sa.ids_prev = ids_prev; sa.lsrc_prev = lsrc;
      sa.rsrc_prev = rsrc; sa.loc = loc; sa.flags = flags; sa.mat_list = mat_list; sa.counts = counts;
      lz_survivors_kernel<<<(unsigned)((Kl * (n + 2) + 255) / 256), 256, 0, st>>>(sa);
      VCSMC_LAUNCH_CHECK("lz_survivors_kernel");
      rc = launch_propose(); if (rc) return rc;
      if (gc) {
        lz_alloc_kernel<<<1, 1024, 0, st>>>(flags, h->pool_slots, counts, h->fetch_cap, mat_list, e_base_prev, fetch_e, loc, slot_id, status);
        VCSMC_LAUNCH_CHECK("lz_alloc_kernel");
      }
      // the survivors' nodes: children, P and slots of rank event r-1 are still in place
      h->prof_begin(3, st);
      rc = launch_materialise(codes, S, pool, S, lsrc, rsrc, mat_list, counts, Kl, loc, e_base_prev,
                              h->p<double>(h->o_P) + ((int64_t)(r - 1) * K + k0) * 32, S, h->jc, st);
      if (rc) return rc;
      if (G > 1) {
        rc = cross_rank_barrier(h, &n_barriers, st);
        if (rc) return rc;
        const int32_t* peer_loc[kMaxPeers];
        const double* peer_pool[kMaxPeers];
        for (int g = 0; g < G; ++g) {
          peer_loc[g] = reinterpret_cast<const int32_t*>(h->peer_ws[g] + h->o_loc);
          peer_pool[g] = reinterpret_cast<const double*>(h->peer_ws[g] + h->o_pool);
        }
        rc = launch_pull(fetch_e, fetch_src, counts + 1, h->fetch_cap < Kl * n ? h->fetch_cap : Kl * n, loc, pool, S, S, G,
                         peer_loc, peer_pool, st);
        if (rc) return rc;
      }
      h->prof_end(st);
    }

    if (r == 0) {
      rc = launch_propose();
      if (rc) return rc;
    }
    lz_resolve_kernel<<<(unsigned)((Kl + 255) / 256), 256, 0, st>>>(Kl, N, row_lref + k0, row_rref + k0, loc, lsrc, rsrc);
    VCSMC_LAUNCH_CHECK("lz_resolve_kernel");

    double* P = h->p<double>(h->o_P) + ((int64_t)r * K + k0) * 32;
    rc = launch_transition_fwd(Q, row_t2 + 2 * k0, 2 * Kl, h->jc, P, st);
    if (rc) return rc;
    if (sorted) {
      rc = group_particles(h, lsrc, rsrc, nullptr, Kl, h->p<int32_t>(h->o_order), h->p<int32_t>(h->o_count), st, leaf_hist != nullptr);
      if (rc) return rc;
    }
    int tiles = 0;
    h->prof_begin(0, st);
    rc = launch_merge_score(codes, S, pool, S, lsrc, rsrc, sorted ? h->p<int32_t>(h->o_order) : nullptr, P, pi, Kl,
                            sorted ? h->p<int32_t>(h->o_count) : nullptr, S, h->jc,
                            leaf_hist, N, h->p<double>(h->o_ell_part), &tiles, st);
    h->prof_end(st);
    if (rc) return rc;

    LzWeightArgs w;
    w.r = r; w.n = n; w.tiles = tiles; w.Kl = Kl;
    w.ell_part = h->p<double>(h->o_ell_part);
    if (h->allreduce) {
      rc = launch_ell_reduce(h->p<double>(h->o_ell_part), tiles, Kl, h->p<double>(h->o_ell_new), st);
      if (rc) return rc;
      if (h->allreduce(h->allreduce_user, h->p<double>(h->o_ell_new), Kl, st)) { set_error("allreduce hook failed"); return VCSMC_ERR_CUDA; }
      w.ell_part = h->p<double>(h->o_ell_new);
      w.tiles = 1;
    }
    w.F = h->p<double>(h->o_F[cur]); w.topo = h->p<double>(h->o_topo[cur]); w.vm = h->p<int32_t>(h->o_vm[cur]);
    w.lam_l = lam_l; w.lam_r = lam_r; w.b_l = row_b_l + k0; w.b_r = row_b_r + k0;
    w.cum_l_prev = r > 0 ? h->p<double>(h->o_cum_l) + (int64_t)(r - 1) * K + k0 : nullptr;
    w.cum_r_prev = r > 0 ? h->p<double>(h->o_cum_r) + (int64_t)(r - 1) * K + k0 : nullptr;
    w.cum_l = h->p<double>(h->o_cum_l) + (int64_t)r * K + k0;
    w.cum_r = h->p<double>(h->o_cum_r) + (int64_t)r * K + k0;
    w.ll_tilde = ll_tilde;
    w.ell_new = ell_node + N + (int64_t)r * K + k0;
    w.lw = h->p<double>(h->o_lw) + (int64_t)r * K + k0;
    w.LL = h->p<double>(h->o_LL) + (int64_t)r * K + k0;
    w.vminus = h->p<int32_t>(h->o_vminus) + k0;
    w.q = 1.0 / ((double)n * (double)(n - 1) / 2.0);
    lz_weights_kernel<<<(unsigned)((Kl + 127) / 128), 128, 0, st>>>(w);
    VCSMC_LAUNCH_CHECK("lz_weights_kernel");

    if (G > 1) {
      RecArgs ra;
      ra.n = n; ra.rank = h->rank; ra.Kl = Kl; ra.k0 = k0; ra.K = K; ra.stride = h->rec_stride; ra.rec = h->p<char>(h->o_rec);
      for (int g = 0; g < kMaxPeers; ++g) ra.peer_rec[g] = nullptr;
      ra.lw = h->p<double>(h->o_lw) + (int64_t)r * K; ra.LL = h->p<double>(h->o_LL) + (int64_t)r * K;
      ra.ell = ell_node + N + (int64_t)r * K; ra.b_l = row_b_l; ra.b_r = row_b_r;
      ra.cum_l = h->p<double>(h->o_cum_l) + (int64_t)r * K; ra.cum_r = h->p<double>(h->o_cum_r) + (int64_t)r * K;
      ra.t2 = row_t2; ra.lref = row_lref; ra.rref = row_rref; ra.nleaf = row_nleaf; ra.vminus = h->p<int32_t>(h->o_vminus);
      ra.rempos = row_rempos;
      lz_pack_kernel<<<(unsigned)((Kl + 127) / 128), 128, 0, st>>>(ra);
      VCSMC_LAUNCH_CHECK("lz_pack_kernel");
      if (h->peer_sync) {
        // every rank has packed its chunk: read the other chunks straight out of the peers' record buffers
        rc = cross_rank_barrier(h, &n_barriers, st);
        if (rc) return rc;
        for (int g = 0; g < kMaxPeers; ++g) ra.peer_rec[g] = g < G ? h->peer_ws[g] + h->o_rec : nullptr;
      } else {
        if (h->comm(h->comm_user, VCSMC_COMM_ALLGATHER, ra.rec, h->rec_stride, st)) { set_error("comm hook failed (all-gather)"); return VCSMC_ERR_CUDA; }
      }
      lz_unpack_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(ra);
      VCSMC_LAUNCH_CHECK("lz_unpack_kernel");
    }
    // log-sum-exp + CDF of this step's weights over ALL particles (every rank: same input, same fixed order)
    rc = launch_resample_cdf(h->p<double>(h->o_lw) + (int64_t)r * K, K, h->p<double>(h->o_cdf), h->p<double>(h->o_stats) + r * 4,
                             h->p<double>(h->o_cdf_scratch), st);
    if (rc) return rc;
  }
  if (G > 1) {
    lz_lltilde_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(K, N, h->p<double>(h->o_LL), h->p<int32_t>(h->o_anc), ll_tilde);
    VCSMC_LAUNCH_CHECK("lz_lltilde_kernel");
  }
  if (n_barriers > 0) {
    lz_advance_kernel<<<1, 1, 0, st>>>(h->p<int32_t>(h->o_epoch_dev), n_barriers);
    VCSMC_LAUNCH_CHECK("lz_advance_kernel");
  }
  return launch_finalize(N, K, h->p<double>(h->o_stats), h->p<double>(h->o_LL) + (int64_t)(N - 2) * K, h->p<double>(h->o_b_l),
                         h->p<double>(h->o_b_r), lam_l, lam_r, log_double_factorial_host(2 * N - 3), h->p<double>(h->o_llR),
                         h->p<double>(h->o_elbo), h->p<double>(h->o_logz), h->p<double>(h->o_ess), st);
}
}  // namespace

int sweep_forward_lazy(vcsmc_sweep* h, const uint8_t* codes, const double* lam_l_in, const double* lam_r_in, const double* Q_in,
                       const double* pi_in, cudaStream_t st) {
  const int N = h->N;
  if (!h->ldf_ready) {  // log-double-factorial table: once per sweep object
    std::vector<double> ldf(2 * N + 4, 0.0);
    for (int m = 0; m < 2 * N + 4; ++m) ldf[m] = log_double_factorial_host(m);
    VCSMC_CUDA(cudaMemcpy(h->p<double>(h->o_ldf), ldf.data(), ldf.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->ldf_ready = true;
  }
  // the model lives in the workspace from here on (the reverse sweep reads it too): the caller's tensors may move
  double* m = h->p<double>(h->o_model);
  double *lam_l = m, *lam_r = m + (N - 1), *Q = m + 2 * (N - 1), *pi = m + 2 * (N - 1) + 16;
  VCSMC_CUDA(cudaMemcpyAsync(lam_l, lam_l_in, (N - 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  VCSMC_CUDA(cudaMemcpyAsync(lam_r, lam_r_in, (N - 1) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (Q_in) VCSMC_CUDA(cudaMemcpyAsync(Q, Q_in, 16 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  VCSMC_CUDA(cudaMemcpyAsync(pi, pi_in, 4 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  lz_set_seed_kernel<<<1, 1, 0, st>>>(h->p<uint64_t>(h->o_seed_dev), h->seed);
  VCSMC_LAUNCH_CHECK("lz_set_seed_kernel");
  count_launch(4);
  h->lam_l = lam_l; h->lam_r = lam_r; h->Q = Q_in ? Q : nullptr; h->pi = pi;
  const double* Qm = Q_in ? Q : nullptr;
  ++h->forwards;

  // a sequence that involves the host (collective hooks, per-launch timing) cannot be captured
  const bool capturable = h->use_graph && !h->profile && !h->allreduce && (h->world == 1 || h->peer_sync);
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
