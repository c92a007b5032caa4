// Likelihood-only scoring of a rank event, materialisation of the survivors, and peer pulls.
//
// With >= ~1000 sites the weights of a rank event spread over hundreds of nats, the ESS is ~1 and almost every
// node the eager forward writes (32 B per particle.site) is dead one resampling later.  The lazy forward therefore
// splits broadcast_conditional_likelihood_K + compute_forest_posterior (vcsmc.py:180-188, :231-245) in two:
//
//   merge_score_kernel   ell[k] = sum_s log(pi . ((L_l P_l) * (L_r P_r))[s]) for EVERY particle, nothing stored.
//                        The site likelihood is a bilinear form of the two children,
//                            x[s] = sum_{j,m} L_a[s][j] L_b[s][m] M_k[j][m],   M_k[j][m] = sum_i pi_i P_a[j][i] P_b[m][i],
//                        so a run of particles that share a child pair (a, b) shares the 16 site products
//                        O[s] = L_a[s] (x) L_b[s] (registers) and each particle costs 16 DFMA per site (general Q) or
//                        4 DFMA (JC: x = a1 sa sb + a2 sa pb + a3 pa sb + a4 pab) instead of ~41; when child a is a
//                        leaf with one-hot / all-ones masks it is row `state` of M dotted with L_b (4 DFMA).
//                        Children come from L2 and are scaled per site by a power of two, so that a particle's running
//                        product of site likelihoods is only multiplied per step and split into (mantissa, exponent)
//                        at the end of a run; bound by the FP64 pipe and the latency around it.
//   (particles whose children are both leaves are scored from the pair's site-pattern counts by the proposing thread
//   of the event kernel, lazy.cu: leaf_pair_hist_kernel tabulates the patterns of every leaf pair once per sweep)
//   merge_score_rows_kernel  particles whose children are a LEAF and an internal node (most of the scored work on
//                        leaf-rich forests): the sites are visited in the leaf's state order (leaf_sort_kernel, once per
//                        sweep), so the 128 positions a warp visits share ONE row of M per particle and a site costs
//                        4 DFMA + one multiplication.
//   materialise_kernel   after the next resampling, only particles that were drawn as an ancestor get their node
//                        written (the plain merge formula; 32 B per site, streaming).
//   pull_kernel          particle sharding: a rank that drew a remote ancestor copies the nodes it lacks straight out
//                        of the owner's pool over NVLink (peer pointers, 256-bit loads), one CTA per (node, site tile).
#include <stdlib.h>
#include <string.h>

#include "launch.h"
#include "merge_device.cuh"

namespace vcsmc {
namespace {

constexpr int kRScore = 32;  // particles per scoring group
constexpr int kCoef = 20;    // doubles per particle in shared memory: M[4][4] (or the 4 JC coefficients) + the column sums of M

// Work items of a scoring launch: groups of R particles x chunks of tiles.  `count` (how many particles the launch really
// has to score) is only known on the device, so the kernel itself decides into how many chunks a group's tiles are split
// (at most n_parts: a particle has that many partial sums).
//   fixed2 == 0: as many chunks as it takes to reach about 4 work items per resident CTA (`slots` of them);
//   fixed2 > 0:  a work item costs a fixed part (staging the group, the final reduction: about fixed2 / 2 tiles' worth
//                of time) plus its tiles, and the items run in rounds of `slots` resident CTAs: the number of chunks that
//                minimises rounds x (fixed + tiles per item).
// (measured with scripts/score_bench: the second rule suits the rows kernel, the first the generic one)
__host__ __device__ __noinline__ void chunking(int64_t count, int R, int tiles, int slots, int fixed2, int n_parts,
                                                  int* tiles_per_item, int* n_chunks) {
  unsigned groups = (unsigned)((count + R - 1) / R);
  if (groups < 1u) groups = 1u;
  int nc_max = n_parts < tiles ? n_parts : tiles;
  if (nc_max < 1) nc_max = 1;
  int best = 1;
  if (fixed2 == 0) {
    best = (int)((4u * (unsigned)slots + groups - 1u) / groups);
    best = best > nc_max ? nc_max : best;
  } else {
    int64_t best_cost = -1;
    for (int nc = 1; nc <= nc_max; ++nc) {
      const int tpi = (tiles + nc - 1) / nc;
      const int real = (tiles + tpi - 1) / tpi;
      if (real != nc) continue;   // (the same split as fewer chunks)
      const unsigned rounds = (groups * (unsigned)nc + (unsigned)slots - 1u) / (unsigned)slots;
      const int64_t cost = (int64_t)rounds * (fixed2 + 2 * tpi);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best = nc;
      }
    }
  }
  *tiles_per_item = (tiles + best - 1) / best;
  *n_chunks = (tiles + *tiles_per_item - 1) / *tiles_per_item;
}

struct ScoreArgs {
  const uint8_t* codes;
  int64_t codes_stride;
  const double* pool;
  int64_t slot_sites;
  const int32_t* lsrc;
  const int32_t* rsrc;
  const int32_t* order;  // grouped visiting order (null: identity)
  const int32_t* count;  // device: number of leading entries of `order` to score (null: K)
  const double* P;       // [K][32]
  const double* pi;
  int64_t K;
  int n_sites;
  int tiles;
  int items;            // resident CTAs of a launch (the kernel splits a group's tiles into chunks accordingly)
  int fixed2;           // twice the fixed cost of a work item, in tiles
  int R;
  int skip_leaf_pairs;  // particles whose children are both leaves are scored from the pattern histogram instead
  int n_parts;          // partial sums per particle in ell_part (>= n_chunks; the tail is zero-filled)
  int from_end;         // the `*count` entries to score are the LAST ones of order[0, K) (the generic list of the grouped order)
  double* ell_part;  // [K][n_parts]
};

constexpr int kScoreSmemBytes = kRScore * kCoef * 8 + kRScore * kTileThreads * (8 + 4);

constexpr int kRenorm = 4;   // tiles between two looks at the running products of a run (power of two)

struct ChildSpace {
  const uint8_t* codes;
  int64_t codes_stride;
  const double* pool;
  int64_t slot_sites;
  int n_sites;
  int skip_leaf_pairs;
};

// The power of two that brings the largest of a site's four partials into [1, 2); its exponent is added to *e_sum.
// Non-negative doubles order like their high words; zero / subnormal / non-finite / negative partials get scale 1 and
// then poison the particle's product, which is redone with one log per site.
__device__ __forceinline__ double site_scale(const d4& L, int* e_sum) {
  const unsigned eu = max(max((unsigned)__double2hiint(L.v[0]), (unsigned)__double2hiint(L.v[1])),
                          max((unsigned)__double2hiint(L.v[2]), (unsigned)__double2hiint(L.v[3]))) >> 20;
  const int em = eu - 1u < 0x7feu ? (int)eu : 1023;
  *e_sum += em - 1023;
  return __hiloint2double((2046 - em) << 20, 0);
}

// Site products of one child pair for the SPT sites of a thread, both children scaled per site (site_scale) so that the
// site likelihoods of all particles are O(M_k) and the running products need no split per step: general Q -> the 16
// products L_a[i] L_b[m]; JC -> (sa sb, sa (pi.L_b), (pi.L_a) sb, sum_i pi_i L_a[i] L_b[i]).  Sites past the end get
// zeros and x0 = 1 (their "likelihood").  The exponents taken out are common to every particle of the run.
template <bool JC, int SPT, int NC>
__device__ __forceinline__ void site_products(const ChildSpace& a, int ca, int cb, int sbase, const double (&pi)[4],
                                              double (&C)[SPT][NC], double (&x0)[SPT], int* e_sum) {
  const ChildRef ra = child_ref(ca, a.codes, a.codes_stride, a.pool, a.slot_sites);
  const ChildRef rb = child_ref(cb, a.codes, a.codes_stride, a.pool, a.slot_sites);
  d4 La[SPT], Lb[SPT];
#pragma unroll
  for (int q = 0; q < SPT; ++q) {   // (no branch on the site: the loads are in flight together)
    const int s = min(sbase + q * kTileThreads, a.n_sites - 1);
    La[q] = load_child(ra, s);
    Lb[q] = load_child(rb, s);
  }
#pragma unroll
  for (int q = 0; q < SPT; ++q) {
    const bool valid = sbase + q * kTileThreads < a.n_sites;
    int e = 0;
    const double sa_ = site_scale(La[q], &e), sb_ = site_scale(Lb[q], &e);
    *e_sum += valid ? e : 0;
    x0[q] = valid ? 0.0 : 1.0;
    const double za = valid ? sa_ : 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      La[q].v[i] *= za;
      Lb[q].v[i] *= sb_;
    }
    if (JC) {
      const double sa = (La[q].v[0] + La[q].v[1]) + (La[q].v[2] + La[q].v[3]);
      const double sb = (Lb[q].v[0] + Lb[q].v[1]) + (Lb[q].v[2] + Lb[q].v[3]);
      double qa = pi[0] * La[q].v[0], qb = pi[0] * Lb[q].v[0], qab = pi[0] * La[q].v[0] * Lb[q].v[0];
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        qa = fma(pi[i], La[q].v[i], qa);
        qb = fma(pi[i], Lb[q].v[i], qb);
        qab = fma(pi[i] * La[q].v[i], Lb[q].v[i], qab);
      }
      C[q][0] = sa * sb;
      C[q][1] = sa * qb;
      C[q][2] = qa * sb;
      C[q][3] = qab;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) C[q][(i * 4 + m) % NC] = La[q].v[i] * Lb[q].v[m];
    }
  }
}

template <int SPT>
__device__ __forceinline__ double site_product(const double (&x)[SPT]) {
  if (SPT == 4) return (x[0] * x[1]) * (x[2] * x[3]);
  double v = x[0];
#pragma unroll
  for (int q = 1; q < SPT; ++q) v *= x[q];
  return v;
}

// NP particles (that share the current child pair) against the SPT sites of this thread: NP*SPT*2 independent FMA
// chains (each site likelihood is accumulated in an even and an odd half) -- the FP64 pipe has a long dependent-issue
// latency and only 4 warps per scheduler fit, so the instruction-level parallelism has to come from here.  The product of
// a particle's site likelihoods is multiplied into its running product (shared memory; loads first, stores last).
template <bool JC, int SPT, int NP>
__device__ __forceinline__ void score_particles(const double* coef, double* pp, const double (&C)[SPT][JC ? 4 : 16],
                                                const double (&x0)[SPT]) {
  constexpr int NC = JC ? 4 : 16;
  double xe[NP][SPT], xo[NP][SPT], p[NP];
#pragma unroll
  for (int n = 0; n < NP; ++n) {
    p[n] = pp[n * kTileThreads];
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
      xe[n][q] = x0[q];
      xo[n][q] = 0.0;
    }
  }
#pragma unroll
  for (int c = 0; c < NC; c += 2) {
    double2 m[NP];
#pragma unroll
    for (int n = 0; n < NP; ++n) m[n] = *reinterpret_cast<const double2*>(coef + n * kCoef + c);
#pragma unroll
    for (int n = 0; n < NP; ++n)
#pragma unroll
      for (int q = 0; q < SPT; ++q) {
        xe[n][q] = fma(m[n].x, C[q][c], xe[n][q]);
        xo[n][q] = fma(m[n].y, C[q][c + 1], xo[n][q]);
      }
  }
#pragma unroll
  for (int n = 0; n < NP; ++n) {
    double x[SPT];
#pragma unroll
    for (int q = 0; q < SPT; ++q) x[q] = xe[n][q] + xo[n][q];
    pp[n * kTileThreads] = p[n] * site_product<SPT>(x);
  }
}

// The same when child a is a LEAF with a one-hot (or all-ones) state mask at every site of this thread: the site
// likelihood is row `state` of M (or the column sums of M) dotted with L_b -- 4 DFMA per site instead of 16.  The row is
// fetched from shared memory per lane (offset roff = 4 * state, or 16 for a gap).  (The grouped order sends these
// particles to the rows kernel; this path serves the identity order of small runs.)
template <int SPT>
__device__ __forceinline__ void score_particle_leaf(const double* coef, double* pp, const double (&Lb)[SPT][16],
                                                    const int (&roff)[SPT], const double (&x0)[SPT]) {
  double x[SPT];
#pragma unroll
  for (int q = 0; q < SPT; ++q) {
    const double2* row = reinterpret_cast<const double2*>(coef + roff[q]);
    const double2 r0 = row[0], r1 = row[1];
    const double xe = fma(r0.y, Lb[q][1], fma(r0.x, Lb[q][0], x0[q]));
    const double xo = fma(r1.y, Lb[q][3], r1.x * Lb[q][2]);
    x[q] = xe + xo;
  }
  pp[0] *= site_product<SPT>(x);
}

// exact fallback for a particle whose product was poisoned: one log per site
template <bool JC>
__device__ __noinline__ double score_slow(ChildSpace a, int ca, int cb, const double* cj, int s_begin, int s_end, double pi0,
                                          double pi1, double pi2, double pi3) {
  constexpr int NC = JC ? 4 : 16;
  const ChildRef ra = child_ref(ca, a.codes, a.codes_stride, a.pool, a.slot_sites);
  const ChildRef rb = child_ref(cb, a.codes, a.codes_stride, a.pool, a.slot_sites);
  const double pi[4] = {pi0, pi1, pi2, pi3};
  double acc = 0.0;
  double m[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) m[c] = cj[c];
  for (int s = s_begin + threadIdx.x; s < s_end; s += kTileThreads) {
    const d4 La = load_child(ra, s), Lb = load_child(rb, s);
    double x = 0.0;
    if (JC) {
      const double sa = (La.v[0] + La.v[1]) + (La.v[2] + La.v[3]);
      const double sb = (Lb.v[0] + Lb.v[1]) + (Lb.v[2] + Lb.v[3]);
      double qa = pi[0] * La.v[0], qb = pi[0] * Lb.v[0], qab = pi[0] * La.v[0] * Lb.v[0];
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        qa = fma(pi[i], La.v[i], qa);
        qb = fma(pi[i], Lb.v[i], qb);
        qab = fma(pi[i] * La.v[i], Lb.v[i], qab);
      }
      x = fma(m[3], qab, fma(m[2], qa * sb, fma(m[1], sa * qb, m[0] * (sa * sb))));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) x = fma(m[(i * 4 + mm) % NC], La.v[i] * Lb.v[mm], x);
    }
    acc += log(x);
  }
  return acc;
}

// One work item: a group of <= R particles (runs of particles that share their child pair) x a chunk of tiles.  The
// running product of the site likelihoods of (thread, particle) lives in shared memory and is only MULTIPLIED per step
// (the children are scaled per site, see site_products); at the end of a run's tiles -- and every kRenorm tiles in
// between if some product has left [2^-480, 2^480) -- the products go back to [1, 2) and their exponents, with the ones
// taken out of the sites, to integer sums.  A product found outside [2^-959, 2^1024), negative or NaN poisons the
// particle, which is re-evaluated with one log per site afterwards (never on sane inputs).
template <bool JC, int SPT>
__global__ void __launch_bounds__(kTileThreads, (JC || SPT <= 2) ? 2 : 1) merge_score_kernel(const ScoreArgs a) {
  constexpr int NC = JC ? 4 : 16;  // coefficients per particle == site products per site
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sC = reinterpret_cast<double*>(smem_raw);                    // [R][20] per-particle coefficients (general: M and its column sums)
  double* s_prod = sC + kRScore * kCoef;                                 // [R][256] running products
  int* s_exp = reinterpret_cast<int*>(s_prod + kRScore * kTileThreads); // [R][256] running exponent sums
  __shared__ int s_k[kRScore], s_a[kRScore], s_b[kRScore];
  __shared__ int s_run[kRScore + 1];   // first particle of every run of one child pair; s_run[n_runs] = nj
  __shared__ int s_nruns, s_work;
  __shared__ double s_slow[kWarps];
  __shared__ unsigned s_odd;
  __shared__ int s_chunking[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int R = a.R;
  const int64_t count = a.count ? (int64_t)*a.count : a.K;
  if (tid == 0) chunking(count, R, a.tiles, a.items, a.fixed2, a.n_parts, &s_chunking[0], &s_chunking[1]);   // (divisions: one thread)
  __syncthreads();
  const int tiles_per_item = s_chunking[0], n_chunks = s_chunking[1];
  const int64_t total = ((count + R - 1) / R) * n_chunks;
  double pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = __ldg(a.pi + j);
  double* my_prod = s_prod + tid;
  int* my_exp = s_exp + tid;
  const ChildSpace cs = {a.codes, a.codes_stride, a.pool, a.slot_sites, a.n_sites, a.skip_leaf_pairs};
  const int64_t first = a.from_end ? a.K - count : 0;

  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t g = w / n_chunks;
    const int tc = (int)(w - g * n_chunks);
    const int64_t j0 = g * R;
    const int nj = (int)min((int64_t)R, count - j0);
    __syncthreads();
    if (tid == 0) s_odd = 0u;
    if (tid < nj) {
      const int k = a.order ? a.order[first + j0 + tid] : (int)(j0 + tid);
      const int ls = a.lsrc[k], rs = a.rsrc[k];
      const bool sw = ls > rs;  // canonical child order a <= b; the swap flag rides in the sign of k
      s_a[tid] = sw ? rs : ls;
      s_b[tid] = sw ? ls : rs;
      s_k[tid] = sw ? ~k : k;
    }
    __syncthreads();
    if (tid == 0) {
      int nr = 0, work = 0;
      for (int j = 0; j < nj; ++j) {
        if (j == 0 || s_a[j] != s_a[j - 1] || s_b[j] != s_b[j - 1]) s_run[nr++] = j;
        work |= !(a.skip_leaf_pairs && s_b[j] < 0);   // a leaf pair is scored from its site patterns elsewhere
      }
      s_run[nr] = nj;
      s_nruns = nr;
      s_work = work;
    }
    __syncthreads();
    if (!s_work) continue;   // nothing of this group is ours
    if (JC) {
      if (tid < nj) {
        const int kk = s_k[tid];
        const bool sw = kk < 0;
        const int64_t k = sw ? ~kk : kk;
        const double* Pa = a.P + k * 32 + (sw ? 16 : 0);
        const double* Pb = a.P + k * 32 + (sw ? 0 : 16);
        const double oa = __ldg(Pa + 1), da = __ldg(Pa) - oa, ob = __ldg(Pb + 1), db = __ldg(Pb) - ob;
        sC[tid * kCoef + 0] = oa * ob * ((pi[0] + pi[1]) + (pi[2] + pi[3]));
        sC[tid * kCoef + 1] = oa * db;
        sC[tid * kCoef + 2] = da * ob;
        sC[tid * kCoef + 3] = da * db;
      }
    } else {
      for (int e = tid; e < nj * 16; e += kTileThreads) {
        const int j = e >> 4, ai = (e >> 2) & 3, bi = e & 3;
        const int kk = s_k[j];
        const bool sw = kk < 0;
        const int64_t k = sw ? ~kk : kk;
        const double* Pa = a.P + k * 32 + (sw ? 16 : 0) + ai * 4;
        const double* Pb = a.P + k * 32 + (sw ? 0 : 16) + bi * 4;
        double m = pi[0] * __ldg(Pa) * __ldg(Pb);
#pragma unroll
        for (int i = 1; i < 4; ++i) m = fma(pi[i] * __ldg(Pa + i), __ldg(Pb + i), m);
        sC[j * kCoef + ai * 4 + bi] = m;
      }
      __syncthreads();
      for (int e = tid; e < nj * 4; e += kTileThreads) {  // column sums: the likelihood row of a gap in child a
        const int j = e >> 2, m = e & 3;
        sC[j * kCoef + 16 + m] = (sC[j * kCoef + m] + sC[j * kCoef + 4 + m]) + (sC[j * kCoef + 8 + m] + sC[j * kCoef + 12 + m]);
      }
    }
    for (int j = 0; j < nj; ++j) {
      my_prod[j * kTileThreads] = 1.0;
      my_exp[j * kTileThreads] = 0;
    }
    __syncthreads();
    const int n_runs = s_nruns;
    const int t_begin = tc * tiles_per_item;
    const int t_end = min(a.tiles, t_begin + tiles_per_item);

    for (int run = 0; run < n_runs; ++run) {
      const int jb = s_run[run], len = s_run[run + 1] - jb;
      const int ca = s_a[jb], cb = s_b[jb];
      if (a.skip_leaf_pairs && cb < 0) continue;
      double* const pp0 = my_prod + jb * kTileThreads;
      int* const pe0 = my_exp + jb * kTileThreads;
      const double* const Mj = sC + jb * kCoef;
      int erun = 0;   // exponents taken out of this thread's sites: common to every particle of the run
      for (int t = t_begin; t < t_end; ++t) {
        const int sbase = t * (kTileThreads * SPT) + tid;
        double C[SPT][NC], x0[SPT];
        bool leaf_rows = false;   // (leaf with one-hot / gap masks, internal node): the 4-DFMA path
        int roff[SPT];
        if (!JC && ca < 0 && cb >= 0) {
          // every site of this warp must have a one-hot or all-ones mask
          const uint8_t* crow = a.codes + (int64_t)(-ca - 1) * a.codes_stride;
          bool ok = true;
#pragma unroll
          for (int q = 0; q < SPT; ++q) {
            const int s = sbase + q * kTileThreads;
            const int code = s < a.n_sites ? __ldg(crow + s) & 15 : 15;
            roff[q] = code == 15 ? 16 : 4 * (__ffs(code) - 1);
            ok = ok && (code == 15 || (code & (code - 1)) == 0);
          }
          leaf_rows = __all_sync(0xffffffffu, ok);
        }
        if (leaf_rows) {
          const double* node = a.pool + (int64_t)cb * a.slot_sites * 4;
          d4 L[SPT];
#pragma unroll
          for (int q = 0; q < SPT; ++q) L[q] = ld_site(node + (int64_t)min(sbase + q * kTileThreads, a.n_sites - 1) * 4);
#pragma unroll
          for (int q = 0; q < SPT; ++q) {
            const bool valid = sbase + q * kTileThreads < a.n_sites;
            int e = 0;
            const double sc = site_scale(L[q], &e);
            erun += valid ? e : 0;
            x0[q] = valid ? 0.0 : 1.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) C[q][m % NC] = valid ? L[q].v[m] * sc : 0.0;
          }
          for (int i = 0; i < len; ++i)
            score_particle_leaf<SPT>(Mj + i * kCoef, pp0 + i * kTileThreads, reinterpret_cast<const double(&)[SPT][16]>(C), roff, x0);
        } else {
          site_products<JC, SPT, NC>(cs, ca, cb, sbase, pi, C, x0, &erun);
          int i = 0;
          for (; i + 1 < len; i += 2) score_particles<JC, SPT, 2>(Mj + i * kCoef, pp0 + i * kTileThreads, C, x0);
          if (i < len) score_particles<JC, SPT, 1>(Mj + i * kCoef, pp0 + i * kTileThreads, C, x0);
        }
        bool renorm = t == t_end - 1;
        if (!renorm && ((t - t_begin) & (kRenorm - 1)) == kRenorm - 1) {
          int hmin = 0x7fffffff, hmax = 0;
          for (int i = 0; i < len; ++i) {
            const int h = reinterpret_cast<const int*>(pp0 + i * kTileThreads)[1];
            hmin = min(hmin, h);
            hmax = max(hmax, h);
          }
          renorm = hmin < ((1023 - 480) << 20) || hmax >= ((1023 + 480) << 20);
        }
        if (renorm) {
          for (int i = 0; i < len; ++i) {
            const double pr = pp0[i * kTileThreads];
            const int h2 = __double2hiint(pr);
            const unsigned e2 = (unsigned)h2 >> 20;   // (sign included)
            const bool sane = e2 - 64u < 0x7ffu - 64u;
            pe0[i * kTileThreads] += sane ? (int)e2 - 1023 + erun : 0;
            pp0[i * kTileThreads] = sane ? __hiloint2double((h2 & 0x000fffff) | 0x3ff00000, __double2loint(pr))
                                         : __longlong_as_double(0x7ff8000000000000ll);
          }
          erun = 0;
        }
      }
    }
    // sum_s log x_s = log(prod of the products) + ln2 * (exponents taken out).  A warp combines the 256 per-thread
    // products of a particle (8 entries per lane, each in [1, 2)) before the log: 32 logs per particle instead of 256.
    __syncthreads();
    for (int j = wid; j < nj; j += kWarps) {
      if (a.skip_leaf_pairs && s_b[j] < 0) continue;
      double p = 1.0;
      int e = 0;
#pragma unroll
      for (int i = 0; i < kTileThreads / 32; ++i) {
        p *= s_prod[j * kTileThreads + lane + 32 * i];
        e += s_exp[j * kTileThreads + lane + 32 * i];
      }
      const double ef = (double)e;
      double acc = fma(ef, 6.93147180369123816490e-01, fma(ef, 1.90821492927058770002e-10, log(p)));  // ln2 hi + lo
      acc = warp_sum(acc);
      if (__any_sync(0xffffffffu, p != p)) {
        if (lane == 0) atomicOr(&s_odd, 1u << j);
      } else {
        const int kk = s_k[j];
        const int64_t k = kk < 0 ? ~kk : kk;
        if (lane == 0) a.ell_part[k * a.n_parts + tc] = acc;
        if (tc == 0)   // the entries no chunk of this kernel writes
          for (int t = n_chunks + lane; t < a.n_parts; t += 32) a.ell_part[k * a.n_parts + t] = 0.0;
      }
    }
    __syncthreads();
    const unsigned odd_all = s_odd;
    for (int j = 0; odd_all && j < nj; ++j) {
      if (!(odd_all >> j & 1u)) continue;
      // some likelihood of this particle is 0 / subnormal / not finite: one log per site, like the reference
      double acc = score_slow<JC>(cs, s_a[j], s_b[j], sC + j * kCoef, t_begin * (kTileThreads * SPT),
                                  min(a.n_sites, t_end * (kTileThreads * SPT)), pi[0], pi[1], pi[2], pi[3]);
      acc = warp_sum(acc);
      __syncthreads();
      if (lane == 0) s_slow[wid] = acc;
      __syncthreads();
      if (tid == 0) {
        const int kk = s_k[j];
        const int64_t k = kk < 0 ? ~kk : kk;
        double t = s_slow[0];
        for (int w2 = 1; w2 < kWarps; ++w2) t += s_slow[w2];
        a.ell_part[k * a.n_parts + tc] = t;
        if (tc == 0)
          for (int q = n_chunks; q < a.n_parts; ++q) a.ell_part[k * a.n_parts + q] = 0.0;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// leaf + internal node: sites in the leaf's state order
// ---------------------------------------------------------------------------------------------
// When child a is a LEAF, L_a[s] is a 0/1 mask and the site likelihood is row `state` of M (the column sums of M for a
// gap) dotted with L_b[s].  leaf_sort_kernel lists every leaf's sites by state class once per sweep (A, C, G, T, gap,
// other ambiguity codes; every class padded to whole 128-position units with -1).  A warp visits one unit per step (a
// lane: four consecutive positions), so all its sites use the SAME row: it is fetched once per particle (a broadcast
// read), the four sites of a lane cost 16 DFMA, and their product is folded into the running (mantissa, exponent) with
// ONE split -- against a per-lane row fetch, 4 + 1 FP64 ops and a split per site in the generic kernel's leaf path, which
// was bound by instruction issue, not by the FP64 pipe.  Class boundaries fall between warps, never inside one, so
// every unit takes the same path; a warp whose unit is empty (class tails, the end of the list) skips it.  The
// internal child is read through the permutation (32 B per site: whole sectors, so the gather costs nothing extra).
constexpr int kRowTile = 1024;            // sorted positions per tile: 8 warps x one unit
constexpr int kRowSub = 128;              // positions per unit: 32 lanes x 4 consecutive positions, ONE state class
constexpr int kLeafClasses = 6;

__device__ __forceinline__ int leaf_class(int code) {
  code &= 15;
  return code == 1 ? 0 : code == 2 ? 1 : code == 4 ? 2 : code == 8 ? 3 : code == 15 ? 4 : 5;
}

// one CTA per leaf; stable (thread t owns a contiguous stretch of sites, classes are placed in thread order), so the
// order -- and with it every floating-point sum over sites -- is reproducible
__global__ void __launch_bounds__(256) leaf_sort_kernel(const uint8_t* __restrict__ codes, int64_t stride, int S, int Sp,
                                                        int32_t* __restrict__ perm, uint8_t* __restrict__ tstate) {
  __shared__ int cnt[256][kLeafClasses];
  __shared__ int start[kLeafClasses + 1];
  const int leaf = blockIdx.x, tid = threadIdx.x;
  const uint8_t* row = codes + (int64_t)leaf * stride;
  int32_t* out = perm + (int64_t)leaf * Sp;
  const int per = (S + 255) / 256, s0 = min(tid * per, S), s1 = min(s0 + per, S);
  int mine[kLeafClasses] = {0, 0, 0, 0, 0, 0};
  for (int s = s0; s < s1; ++s) ++mine[leaf_class(row[s])];
#pragma unroll
  for (int c = 0; c < kLeafClasses; ++c) cnt[tid][c] = mine[c];
  for (int p = tid; p < Sp; p += 256) out[p] = -1;
  __syncthreads();
  if (tid == 0) {
    int at = 0;
    for (int c = 0; c < kLeafClasses; ++c) {
      int tot = 0;
      for (int t = 0; t < 256; ++t) tot += cnt[t][c];
      start[c] = at;
      at += (tot + kRowSub - 1) / kRowSub * kRowSub;
    }
    start[kLeafClasses] = at;
  }
  __syncthreads();
  int pos[kLeafClasses];
#pragma unroll
  for (int c = 0; c < kLeafClasses; ++c) {
    int before = 0;
    for (int t = 0; t < tid; ++t) before += cnt[t][c];
    pos[c] = start[c] + before;
  }
  for (int s = s0; s < s1; ++s) {
    const int c = leaf_class(row[s]);
    int p = 0;
#pragma unroll
    for (int q = 0; q < kLeafClasses; ++q)
      if (q == c) p = pos[q]++;
    out[p] = s;
  }
  const int n_sub = Sp / kRowSub;
  for (int t = tid; t < n_sub; t += 256) {
    const int p = t * kRowSub;
    int c = 255;   // no sites
    for (int q = 0; q < kLeafClasses; ++q)
      if (p >= start[q] && p < start[q + 1]) c = q;
    tstate[(int64_t)leaf * n_sub + t] = (uint8_t)c;
  }
}

struct RowArgs {
  const uint8_t* codes;
  int64_t codes_stride;
  const double* pool;
  int64_t slot_sites;
  const int32_t* lsrc;
  const int32_t* rsrc;
  const int32_t* order;   // the first *count entries: particles with one leaf child and one internal child, grouped by pair
  const int32_t* count;
  const double* P;
  const double* pi;
  const int32_t* perm;    // [N][Sp] sites of every leaf in state order (-1: padding)
  const uint8_t* tstate;  // [N][Sp / 128] state class of every unit (255: empty)
  int Sp, tiles, items, fixed2, R, n_parts;
  double* ell_part;
};

// exact fallback for a particle whose product was poisoned: one log per site
__device__ __noinline__ double rows_slow(const RowArgs a, int leaf, int cb, const double* M, int p_begin, int p_end) {
  const uint8_t* crow = a.codes + (int64_t)leaf * a.codes_stride;
  const double* node = a.pool + (int64_t)cb * a.slot_sites * 4;
  const int32_t* pm = a.perm + (int64_t)leaf * a.Sp;
  double acc = 0.0;
  for (int p = p_begin + threadIdx.x; p < p_end; p += kTileThreads) {
    const int s = pm[p];
    if (s < 0) continue;
    const d4 La = leaf_site(__ldg(crow + s)), Lb = ld_site(node + (int64_t)s * 4);
    double x = 0.0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      double d = M[jj * 4] * Lb.v[0];
#pragma unroll
      for (int m = 1; m < 4; ++m) d = fma(M[jj * 4 + m], Lb.v[m], d);
      x = fma(La.v[jj], d, x);
    }
    acc += log(x);
  }
  return acc;
}

// Two particles of a run against the four sites of a lane: their rows of M from shared memory (broadcast reads), 32 DFMA
// in eight independent chains, and the product of each particle's four site likelihoods multiplied into its running
// product (shared memory; distinct words per particle, which the compiler cannot know: loads first, stores last).
__device__ __forceinline__ void rows_trip(const double* rowp, double* pp, const double (&Lb)[4][4], const double (&x0)[4]) {
  const double2 a0 = *reinterpret_cast<const double2*>(rowp), a1 = *reinterpret_cast<const double2*>(rowp + 2);
  const double2 b0 = *reinterpret_cast<const double2*>(rowp + kCoef), b1 = *reinterpret_cast<const double2*>(rowp + kCoef + 2);
  const double pa = pp[0], pb = pp[kTileThreads];
  double xa[4], xb[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    xa[q] = fma(a0.x, Lb[q][0], x0[q]);
    xb[q] = fma(b0.x, Lb[q][0], x0[q]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    xa[q] = fma(a0.y, Lb[q][1], xa[q]);
    xb[q] = fma(b0.y, Lb[q][1], xb[q]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    xa[q] = fma(a1.x, Lb[q][2], xa[q]);
    xb[q] = fma(b1.x, Lb[q][2], xb[q]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    xa[q] = fma(a1.y, Lb[q][3], xa[q]);
    xb[q] = fma(b1.y, Lb[q][3], xb[q]);
  }
  pp[0] = pa * ((xa[0] * xa[1]) * (xa[2] * xa[3]));
  pp[kTileThreads] = pb * ((xb[0] * xb[1]) * (xb[2] * xb[3]));
}

__global__ void __launch_bounds__(kTileThreads, 2) merge_score_rows_kernel(const RowArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sC = reinterpret_cast<double*>(smem_raw);                    // [R][20]: M[4][4] and its column sums
  double* s_prod = sC + kRScore * kCoef;                                 // [R][256] running mantissa products
  int* s_exp = reinterpret_cast<int*>(s_prod + kRScore * kTileThreads); // [R][256] running (biased) exponent sums
  __shared__ int s_k[kRScore], s_a[kRScore], s_b[kRScore];
  __shared__ int s_run[kRScore + 1];   // first particle of every run of one (leaf, node) pair; s_run[n_runs] = nj
  __shared__ int s_nruns;
  __shared__ double s_slow[kWarps];
  __shared__ unsigned s_odd;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int R = a.R;
  const int64_t count = (int64_t)*a.count;
  // chunks of tiles per group of particles, for the particles there really are
  __shared__ int s_chunking[2];
  if (tid == 0) chunking(count, R, a.tiles, a.items, a.fixed2, a.n_parts, &s_chunking[0], &s_chunking[1]);   // (divisions: one thread)
  __syncthreads();
  const int tiles_per_item = s_chunking[0], n_chunks = s_chunking[1];
  const int64_t total = ((count + R - 1) / R) * n_chunks;
  const int n_sub = a.Sp / kRowSub;
  double pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = __ldg(a.pi + j);
  double* my_prod = s_prod + tid;
  int* my_exp = s_exp + tid;

  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t g = w / n_chunks;
    const int tc = (int)(w - g * n_chunks);
    const int64_t j0 = g * R;
    const int nj = (int)min((int64_t)R, count - j0);
    __syncthreads();
    if (tid == 0) s_odd = 0u;
    if (tid < nj) {
      const int k = a.order[j0 + tid];
      const int ls = a.lsrc[k], rs = a.rsrc[k];
      const bool sw = ls > rs;          // the leaf (negative reference) comes first; the swap flag rides in the sign of k
      s_a[tid] = -(sw ? rs : ls) - 1;   // leaf index
      s_b[tid] = sw ? ls : rs;          // slot of the internal node
      s_k[tid] = sw ? ~k : k;
    }
    __syncthreads();
    if (tid == 0) {
      int nr = 0;
      for (int j = 0; j < nj; ++j)
        if (j == 0 || s_a[j] != s_a[j - 1] || s_b[j] != s_b[j - 1]) s_run[nr++] = j;
      s_run[nr] = nj;
      s_nruns = nr;
    }
    for (int e = tid; e < nj * 16; e += kTileThreads) {   // M[j][m] = sum_i pi_i P_leaf[j][i] P_node[m][i]
      const int j = e >> 4, ai = (e >> 2) & 3, bi = e & 3;
      const int kk = s_k[j];
      const bool sw = kk < 0;
      const int64_t k = sw ? ~kk : kk;
      const double* Pa = a.P + k * 32 + (sw ? 16 : 0) + ai * 4;
      const double* Pb = a.P + k * 32 + (sw ? 0 : 16) + bi * 4;
      double m = pi[0] * __ldg(Pa) * __ldg(Pb);
#pragma unroll
      for (int i = 1; i < 4; ++i) m = fma(pi[i] * __ldg(Pa + i), __ldg(Pb + i), m);
      sC[j * kCoef + ai * 4 + bi] = m;
    }
    __syncthreads();
    for (int e = tid; e < nj * 4; e += kTileThreads) {  // column sums: the row of a gap
      const int j = e >> 2, m = e & 3;
      sC[j * kCoef + 16 + m] = (sC[j * kCoef + m] + sC[j * kCoef + 4 + m]) + (sC[j * kCoef + 8 + m] + sC[j * kCoef + 12 + m]);
    }
    for (int j = 0; j < nj; ++j) {
      my_prod[j * kTileThreads] = 1.0;
      my_exp[j * kTileThreads] = 0;
    }
    __syncthreads();
    const int n_runs = s_nruns;

    const int t_begin = tc * tiles_per_item;
    const int t_end = min(a.tiles, t_begin + tiles_per_item);
    for (int run = 0; run < n_runs; ++run) {
      const int jb = s_run[run], len = s_run[run + 1] - jb;
      const int ca = s_a[jb], cb = s_b[jb];
      const double* node = a.pool + (int64_t)cb * a.slot_sites * 4;
      const int4* pm = reinterpret_cast<const int4*>(a.perm + (int64_t)ca * a.Sp) + tid;   // + t * 256: the lane's four positions of tile t
      const uint8_t* ts = a.tstate + (int64_t)ca * n_sub + wid;                            // + t * 8: the class of the warp's unit of tile t
      double* const pp0 = my_prod + jb * kTileThreads;
      int* const pe0 = my_exp + jb * kTileThreads;
      const double* const Mj = sC + jb * kCoef;
      int erun = 0;   // exponents taken out of this lane's sites: common to every particle of the run
      int cls_n = ts[t_begin * kWarps];
      int4 s4_n = __ldg(pm + t_begin * kTileThreads);
      for (int t = t_begin; t < t_end; ++t) {
        const int cls = cls_n;
        const int4 s4 = s4_n;
        if (t + 1 < t_end) {   // the next unit's class and positions are fetched while this one is scored
          cls_n = ts[(t + 1) * kWarps];
          s4_n = __ldg(pm + (t + 1) * kTileThreads);
        }
        if (cls != 255) {   // (warp-uniform: 255 = nothing of this leaf in the unit)
          const int sq[4] = {s4.x, s4.y, s4.z, s4.w};
          double Lb[4][4], x0[4];
          int code[4] = {0, 0, 0, 0};
          d4 Lraw[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) Lraw[q] = ld_site(node + (int64_t)max(sq[q], 0) * 4);   // (no branch: the four loads are in flight together)
          if (cls == 5) {
#pragma unroll
            for (int q = 0; q < 4; ++q) code[q] = sq[q] >= 0 ? __ldg(a.codes + (int64_t)ca * a.codes_stride + sq[q]) & 15 : 0;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            // the site's partials scaled by a power of two so that the largest lies in [1, 2): the site likelihoods of
            // all particles are then O(M), the running products need no split per step, and the exponent taken out
            // is the same for every particle of the run
            const d4& L = Lraw[q];
            // (non-negative doubles order like their high words; a negative or NaN partial gives an exponent out of range)
            const unsigned eu = max(max((unsigned)__double2hiint(L.v[0]), (unsigned)__double2hiint(L.v[1])),
                                    max((unsigned)__double2hiint(L.v[2]), (unsigned)__double2hiint(L.v[3]))) >> 20;
            const bool valid = sq[q] >= 0;                                // padding: no site, likelihood 1
            const int em = eu - 1u < 0x7feu ? (int)eu : 1023;             // zero / subnormal / non-finite / negative partials stay as they are (and poison the product)
            const double sc = valid ? __hiloint2double((2046 - em) << 20, 0) : 0.0;
            erun += valid ? em - 1023 : 0;
#pragma unroll
            for (int m = 0; m < 4; ++m) Lb[q][m] = L.v[m] * sc;
            x0[q] = valid ? 0.0 : 1.0;
          }
          double* pp = pp0;
          if (cls != 5) {
            // one row of M (or its column sums) for the whole unit: a broadcast read per particle
            const double* rowp = Mj + (cls == 4 ? 16 : 4 * cls);
            // two particles per trip (rows_trip), two trips unrolled so that the shared-memory addresses are immediates
            const int pairs = len >> 1;
#pragma unroll 2
            for (int i = 0; i < pairs; ++i) rows_trip(rowp + 2 * kCoef * i, pp + 2 * kTileThreads * i, Lb, x0);
            if (len & 1) {
              rowp += 2 * kCoef * pairs;
              pp += 2 * kTileThreads * pairs;
              const double2 r0 = *reinterpret_cast<const double2*>(rowp), r1 = *reinterpret_cast<const double2*>(rowp + 2);
              double x[4];
#pragma unroll
              for (int q = 0; q < 4; ++q)
                x[q] = fma(r1.y, Lb[q][3], fma(r1.x, Lb[q][2], fma(r0.y, Lb[q][1], fma(r0.x, Lb[q][0], x0[q]))));
              pp[0] *= (x[0] * x[1]) * (x[2] * x[3]);
            }
          } else {
            // other ambiguity codes in the unit: the rows each site's mask covers, summed per site
            const double* Mi = Mj;
            for (int i = 0; i < len; ++i, pp += kTileThreads, Mi += kCoef) {
              double x[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                double acc = x0[q];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  if (code[q] >> jj & 1) {
                    double d = Mi[jj * 4] * Lb[q][0];
#pragma unroll
                    for (int m = 1; m < 4; ++m) d = fma(Mi[jj * 4 + m], Lb[q][m], d);
                    acc += d;
                  }
                }
                x[q] = acc;
              }
              pp[0] *= (x[0] * x[1]) * (x[2] * x[3]);
            }
          }
        }
        // At the end of the run's tiles the running products go back to [1, 2) and their exponents, with the ones taken
        // out of the sites, to the integer sums; every kRenorm tiles in between only if some product has left
        // [2^-480, 2^480) (a look at the high words).  kRenorm further steps keep a product a normal number unless some
        // M_k or partial is below ~1e-9 (each step multiplies by four likelihoods of O(M_k)): a product found outside
        // [2^-959, 2^1024) -- or negative, or NaN -- poisons the particle, which is then redone with one log per site.
        bool renorm = t == t_end - 1;
        if (!renorm && ((t - t_begin) & (kRenorm - 1)) == kRenorm - 1) {
          int hmin = 0x7fffffff, hmax = 0;
          for (int i = 0; i < len; ++i) {
            const int h = reinterpret_cast<const int*>(pp0 + i * kTileThreads)[1];
            hmin = min(hmin, h);
            hmax = max(hmax, h);
          }
          renorm = hmin < ((1023 - 480) << 20) || hmax >= ((1023 + 480) << 20);
        }
        if (renorm) {
          for (int i = 0; i < len; ++i) {
            const double pr = pp0[i * kTileThreads];
            const int h2 = __double2hiint(pr);
            const unsigned e2 = (unsigned)h2 >> 20;   // (sign included)
            const bool sane = e2 - 64u < 0x7ffu - 64u;
            pe0[i * kTileThreads] += sane ? (int)e2 - 1023 + erun : 0;
            pp0[i * kTileThreads] = sane ? __hiloint2double((h2 & 0x000fffff) | 0x3ff00000, __double2loint(pr))
                                         : __longlong_as_double(0x7ff8000000000000ll);
          }
          erun = 0;
        }
      }
    }
    // sum_s log x_s = log(prod mantissas) + ln2 * sum of the exponents.  A warp combines the 256 per-thread products of a
    // particle (8 entries per lane) before the log: 32 logs per particle instead of 256.  Entries are products of at most
    // (tiles of the item) mantissas in [1, 2): with few tiles eight of them multiply without renormalisation.
    const bool few = t_end - t_begin <= 100;
    __syncthreads();
    for (int j = wid; j < nj; j += kWarps) {
      double p = 1.0;
      int e = 0;
      if (few) {
#pragma unroll
        for (int i = 0; i < kTileThreads / 32; ++i) {
          p *= s_prod[j * kTileThreads + lane + 32 * i];
          e += s_exp[j * kTileThreads + lane + 32 * i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < kTileThreads / 32; ++i) {
          p *= s_prod[j * kTileThreads + lane + 32 * i];
          e += s_exp[j * kTileThreads + lane + 32 * i];
          const int hi = __double2hiint(p);
          const int ee = (hi >> 20) & 0x7ff;
          if (ee != 0x7ff) {   // (a poisoned product stays NaN)
            p = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(p));
            e += ee - 1023;
          }
        }
      }
      const double ef = (double)e;
      double acc = fma(ef, 6.93147180369123816490e-01, fma(ef, 1.90821492927058770002e-10, log(p)));  // ln2 hi + lo
      acc = warp_sum(acc);
      if (__any_sync(0xffffffffu, p != p)) {
        if (lane == 0) atomicOr(&s_odd, 1u << j);
      } else {
        const int kk = s_k[j];
        const int64_t k = kk < 0 ? ~kk : kk;
        if (lane == 0) a.ell_part[k * a.n_parts + tc] = acc;
        if (tc == 0)
          for (int t = n_chunks + lane; t < a.n_parts; t += 32) a.ell_part[k * a.n_parts + t] = 0.0;
      }
    }
    __syncthreads();
    const unsigned odd_all = s_odd;
    for (int j = 0; odd_all && j < nj; ++j) {
      if (!(odd_all >> j & 1u)) continue;
      double acc = rows_slow(a, s_a[j], s_b[j], sC + j * kCoef, t_begin * kRowTile, t_end * kRowTile);
      acc = warp_sum(acc);
      __syncthreads();
      if (lane == 0) s_slow[wid] = acc;
      __syncthreads();
      if (tid == 0) {
        const int kk = s_k[j];
        const int64_t k = kk < 0 ? ~kk : kk;
        double t = s_slow[0];
        for (int w2 = 1; w2 < kWarps; ++w2) t += s_slow[w2];
        a.ell_part[k * a.n_parts + tc] = t;
        if (tc == 0)
          for (int q = n_chunks; q < a.n_parts; ++q) a.ell_part[k * a.n_parts + q] = 0.0;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// survivors: the plain merge, stored
// ---------------------------------------------------------------------------------------------
struct MatArgs {
  const uint8_t* codes;
  int64_t codes_stride;
  double* pool;
  int64_t slot_sites;
  const int32_t* lsrc;   // [Kl] child slots of the event being materialised
  const int32_t* rsrc;
  const int32_t* list;   // local particle indices to materialise
  const int32_t* count;  // device: number of list entries
  const int32_t* loc;    // node -> local slot
  int64_t e_base;        // node index of local particle 0 of that event
  const double* P;       // [Kl][32] of that event
  int n_sites;
  int tiles;
};

constexpr int kMatSpt = 2;

template <bool JC>
__global__ void __launch_bounds__(kTileThreads) materialise_kernel(const MatArgs a) {
  const int64_t total = (int64_t)(*a.count) * a.tiles;
  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t j = w / a.tiles;
    const int t = (int)(w - j * a.tiles);
    const int kl = a.list[j];
    const int ds = a.loc[a.e_base + kl];
    if (ds < 0) continue;  // pool exhausted (reported through the status word)
    const ChildRef ra = child_ref(a.lsrc[kl], a.codes, a.codes_stride, a.pool, a.slot_sites);
    const ChildRef rb = child_ref(a.rsrc[kl], a.codes, a.codes_stride, a.pool, a.slot_sites);
    Trans<JC> Pl, Pr;
    Pl.load(a.P + (int64_t)kl * 32);
    Pr.load(a.P + (int64_t)kl * 32 + 16);
    double* out = a.pool + (int64_t)ds * a.slot_sites * 4;
#pragma unroll
    for (int q = 0; q < kMatSpt; ++q) {
      const int s = t * (kTileThreads * kMatSpt) + q * kTileThreads + threadIdx.x;
      if (s < a.n_sites) {
        const d4 lp = Pl.apply(load_child(ra, s)), rp = Pr.apply(load_child(rb, s));
        d4 nw;
#pragma unroll
        for (int i = 0; i < 4; ++i) nw.v[i] = lp.v[i] * rp.v[i];
        st_site(out + (int64_t)s * 4, nw);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// particle sharding: copy missing nodes out of the owner's pool (peer memory over NVLink)
// ---------------------------------------------------------------------------------------------
struct PullArgs {
  const int32_t* fetch_e;    // node indices to fetch
  const int32_t* fetch_src;  // rank that holds a copy
  const int32_t* count;      // device: number of entries
  const int32_t* loc;        // local node -> slot (destination, assigned by the allocator)
  double* pool;
  int64_t slot_sites;
  int n_sites;
  int tiles;
  const int32_t* peer_loc[kMaxPeers];
  const double* peer_pool[kMaxPeers];
};

__global__ void __launch_bounds__(kTileThreads) pull_kernel(const PullArgs a) {
  const int64_t total = (int64_t)(*a.count) * a.tiles;
  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t j = w / a.tiles;
    const int t = (int)(w - j * a.tiles);
    const int e = a.fetch_e[j], g = a.fetch_src[j];
    const int ds = a.loc[e];
    const int ss = a.peer_loc[g][e];
    if (ds < 0 || ss < 0) continue;
    const double* src = a.peer_pool[g] + (int64_t)ss * a.slot_sites * 4;
    double* dst = a.pool + (int64_t)ds * a.slot_sites * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int s = t * (kTileThreads * 4) + q * kTileThreads + threadIdx.x;
      if (s < a.n_sites) st_site(dst + (int64_t)s * 4, ld_site(src + (int64_t)s * 4));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// leaf-leaf merges: site patterns instead of sites
// ---------------------------------------------------------------------------------------------
// When both children are leaves the site likelihood depends on the site only through the two 4-bit state masks:
//   x(ca, cb) = sum_i pi_i (sum_{j in ca} P_a[j][i]) (sum_{m in cb} P_b[m][i]),
// so sum_s log x[s] = sum over the <= 256 patterns of count(ca, cb) * log x(ca, cb).  The counts of every leaf pair are
// tabulated once per sweep (site-pattern compression, the standard trick of likelihood codes, applied per cherry); a
// cherry then costs a few dozen logs per particle instead of S site evaluations.  About a third of all merges of a
// sweep join two leaves.

// Pattern table of one leaf pair (kLeafPairInts ints): [0,16) counts of the one-hot x one-hot patterns (state j of the
// lower leaf, state m of the higher one, at j*4+m); [16] number of other patterns with a nonzero count; then
// (mask_a * 16 + mask_b, count) pairs for those.  The event kernel scores a cherry from it in the proposing thread.
__global__ void __launch_bounds__(256) leaf_pair_hist_kernel(const uint8_t* __restrict__ codes, int64_t stride, int N, int S,
                                                             int32_t* __restrict__ table) {
  __shared__ int sh[256];
  // blockIdx.x enumerates the pairs a < b in leaf_pair_index order
  int a = 0;
  int64_t rem = blockIdx.x;
  while (rem >= N - 1 - a) {
    rem -= N - 1 - a;
    ++a;
  }
  const int b = a + 1 + (int)rem;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint8_t* ra = codes + (int64_t)a * stride;
  const uint8_t* rb = codes + (int64_t)b * stride;
  for (int s = threadIdx.x; s < S; s += 256) atomicAdd(&sh[(ra[s] & 15) * 16 + (rb[s] & 15)], 1);
  __syncthreads();
  int32_t* out = table + (int64_t)blockIdx.x * kLeafPairInts;
  if (threadIdx.x < 16) out[threadIdx.x] = sh[(1 << (threadIdx.x >> 2)) * 16 + (1 << (threadIdx.x & 3))];
  if (threadIdx.x == 0) {
    int n = 0;
    for (int bin = 0; bin < 256; ++bin) {
      const int ca = bin >> 4, cb = bin & 15;
      const bool onehot = ca != 0 && cb != 0 && (ca & (ca - 1)) == 0 && (cb & (cb - 1)) == 0;
      if (!onehot && sh[bin] != 0) {
        out[17 + 2 * n] = bin;
        out[18 + 2 * n] = sh[bin];
        ++n;
      }
    }
    out[16] = n;
  }
}

constexpr int64_t kScoreItems = 148 * 2;  // resident CTAs of a scoring launch (two per SM): the rounds of the chunking cost model

}  // namespace

int64_t leaf_pair_hist_ints(int N) { return (int64_t)N * (N - 1) / 2 * kLeafPairInts; }

int launch_leaf_pair_hist(const uint8_t* codes, int64_t stride, int N, int S, int32_t* hist, cudaStream_t st) {
  if (N < 2 || S <= 0) return VCSMC_OK;
  leaf_pair_hist_kernel<<<(unsigned)((int64_t)N * (N - 1) / 2), 256, 0, st>>>(codes, stride, N, S, hist);
  VCSMC_LAUNCH_CHECK("leaf_pair_hist_kernel");
  return VCSMC_OK;
}

int leaf_sort_stride(int n_sites) {   // padded length of a leaf's sorted site list: every class ends on a unit boundary
  return (n_sites + kLeafClasses * kRowSub + kRowTile - 1) / kRowTile * kRowTile;
}

int launch_leaf_sort(const uint8_t* codes, int64_t stride, int N, int S, int32_t* perm, uint8_t* tstate, cudaStream_t st) {
  if (N < 1 || S <= 0) return VCSMC_OK;
  leaf_sort_kernel<<<N, 256, 0, st>>>(codes, stride, S, leaf_sort_stride(S), perm, tstate);
  VCSMC_LAUNCH_CHECK("leaf_sort_kernel");
  return VCSMC_OK;
}

constexpr int kScoreParts = 8;   // partial sums per particle (upper bound of the chunks a group's tiles are split into)

// order == null: every particle in identity order through the generic kernel (its leaf path included).
// order != null: the grouped order of the event kernel -- count[0] leaf + internal particles at the front (rows kernel,
// needs leaf_perm / leaf_tstate), count[1] internal + internal particles at the end (generic kernel).
// skip_leaf_pairs: particles with two leaf children have been scored from the pair's site patterns by the event kernel.
int launch_merge_score(const uint8_t* codes, int64_t codes_stride, const double* pool, int64_t slot_sites,
                       const int32_t* lsrc, const int32_t* rsrc, const int32_t* order, const double* P, const double* pi,
                       int64_t K, const int32_t* count, int n_sites, int jc, int skip_leaf_pairs,
                       const int32_t* leaf_perm, const uint8_t* leaf_tstate, double* ell_part, int* n_parts, cudaStream_t st,
                       cudaStream_t st_generic) {
  // st_generic (optional): the generic kernel runs there, beside the rows kernel on `st` (the caller joins the streams)
  if (n_parts) *n_parts = 0;
  if (K <= 0 || n_sites <= 0) return VCSMC_OK;
  static int spt_general = 0;
  if (spt_general == 0) {
    const char* e = getenv("VCSMC_SCORE_SPT");  // tuning knob: sites per thread of the general-Q scoring kernel
    spt_general = e ? atoi(e) : 2;
    if (spt_general != 2 && spt_general != 4) spt_general = 2;
    VCSMC_CUDA(cudaFuncSetAttribute(merge_score_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScoreSmemBytes));
    VCSMC_CUDA(cudaFuncSetAttribute(merge_score_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScoreSmemBytes));
    VCSMC_CUDA(cudaFuncSetAttribute(merge_score_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScoreSmemBytes));
    VCSMC_CUDA(cudaFuncSetAttribute(merge_score_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kScoreSmemBytes));
  }
  static int64_t items = 0;
  if (items == 0) {
    const char* e = getenv("VCSMC_SCORE_ITEMS");   // tuning knob: resident CTAs assumed by the chunking
    items = e ? atoll(e) : kScoreItems;
    if (items < 1) items = kScoreItems;
  }
  const int spt = jc ? 4 : spt_general;
  const bool rows = order != nullptr && leaf_perm != nullptr;
  ScoreArgs a;
  a.codes = codes; a.codes_stride = codes_stride; a.pool = pool; a.slot_sites = slot_sites; a.lsrc = lsrc; a.rsrc = rsrc;
  a.order = order; a.count = order ? (rows ? count + 1 : count) : nullptr; a.P = P; a.pi = pi; a.K = K; a.n_sites = n_sites; a.ell_part = ell_part;
  a.skip_leaf_pairs = skip_leaf_pairs;
  a.from_end = rows ? 1 : 0;
  a.tiles = (n_sites + kTileThreads * spt - 1) / (kTileThreads * spt);
  a.items = (int)items;
  a.fixed2 = 0;
  // groups as large as the machine fill allows (site data is amortised over the group); K bounds the particles of a launch
  {
    int64_t R = (K * a.tiles) / (4 * items);
    a.R = (int)(R < 1 ? 1 : R > kRScore ? kRScore : R);
  }
  RowArgs b;
  memset(&b, 0, sizeof(b));
  if (rows) {
    b.codes = codes; b.codes_stride = codes_stride; b.pool = pool; b.slot_sites = slot_sites; b.lsrc = lsrc; b.rsrc = rsrc;
    b.order = order; b.count = count; b.P = P; b.pi = pi; b.perm = leaf_perm; b.tstate = leaf_tstate; b.ell_part = ell_part;
    b.Sp = leaf_sort_stride(n_sites);
    b.tiles = b.Sp / kRowTile;
    b.items = (int)items;
    b.fixed2 = 3;
    if (const char* e = getenv("VCSMC_ROWS_FIXED2")) b.fixed2 = atoi(e);   // tuning knob
    b.R = kRScore;
  }
  a.n_parts = kScoreParts;
  b.n_parts = kScoreParts;
  if (n_parts) *n_parts = kScoreParts;
  static int64_t cap = 0;
  if (cap == 0) {
    const char* e = getenv("VCSMC_SCORE_GRID");   // tuning knob: CTAs per scoring launch (each loops over the work items)
    cap = e ? atoll(e) : 148 * 4;   // two rounds of resident CTAs: few enough that a CTA's start-up does not show, enough to even out the items
  }
  {
    int tpi, nc;
    chunking(K, a.R, a.tiles, a.items, a.fixed2, a.n_parts, &tpi, &nc);
    const int64_t total = ((K + a.R - 1) / a.R) * (order ? (int64_t)kScoreParts : (int64_t)nc);   // upper bound of the work items
    const unsigned grid = (unsigned)(total < cap ? total : cap);
    cudaStream_t sg = (rows && st_generic) ? st_generic : st;
    if (jc) merge_score_kernel<true, 4><<<grid, kTileThreads, kScoreSmemBytes, sg>>>(a);
    else if (spt == 4) merge_score_kernel<false, 4><<<grid, kTileThreads, kScoreSmemBytes, sg>>>(a);
    else merge_score_kernel<false, 2><<<grid, kTileThreads, kScoreSmemBytes, sg>>>(a);
    VCSMC_LAUNCH_CHECK("merge_score_kernel");
  }
  if (rows) {
    const int64_t total = ((K + b.R - 1) / b.R) * kScoreParts;
    const unsigned grid = (unsigned)(total < cap ? total : cap);
    merge_score_rows_kernel<<<grid, kTileThreads, kScoreSmemBytes, st>>>(b);
    VCSMC_LAUNCH_CHECK("merge_score_rows_kernel");
  }
  return VCSMC_OK;
}

int launch_materialise(const uint8_t* codes, int64_t codes_stride, double* pool, int64_t slot_sites, const int32_t* lsrc,
                       const int32_t* rsrc, const int32_t* list, const int32_t* count, int64_t max_count, const int32_t* loc,
                       int64_t e_base, const double* P, int n_sites, int jc, cudaStream_t st) {
  if (max_count <= 0 || n_sites <= 0) return VCSMC_OK;
  MatArgs a;
  a.codes = codes; a.codes_stride = codes_stride; a.pool = pool; a.slot_sites = slot_sites; a.lsrc = lsrc; a.rsrc = rsrc;
  a.list = list; a.count = count; a.loc = loc; a.e_base = e_base; a.P = P; a.n_sites = n_sites;
  a.tiles = (n_sites + kTileThreads * kMatSpt - 1) / (kTileThreads * kMatSpt);
  const int64_t total = max_count * a.tiles, cap = 148 * 8;
  const unsigned grid = (unsigned)(total < cap ? total : cap);
  if (jc) materialise_kernel<true><<<grid, kTileThreads, 0, st>>>(a);
  else materialise_kernel<false><<<grid, kTileThreads, 0, st>>>(a);
  VCSMC_LAUNCH_CHECK("materialise_kernel");
  return VCSMC_OK;
}

int launch_pull(const int32_t* fetch_e, const int32_t* fetch_src, const int32_t* count, int64_t max_count, const int32_t* loc,
                double* pool, int64_t slot_sites, int n_sites, int world, const int32_t* const* peer_loc,
                const double* const* peer_pool, cudaStream_t st) {
  if (max_count <= 0 || n_sites <= 0) return VCSMC_OK;
  PullArgs a;
  a.fetch_e = fetch_e; a.fetch_src = fetch_src; a.count = count; a.loc = loc; a.pool = pool; a.slot_sites = slot_sites;
  a.n_sites = n_sites; a.tiles = (n_sites + kTileThreads * 4 - 1) / (kTileThreads * 4);
  for (int g = 0; g < kMaxPeers; ++g) {
    a.peer_loc[g] = g < world ? peer_loc[g] : nullptr;
    a.peer_pool[g] = g < world ? peer_pool[g] : nullptr;
  }
  const int64_t total = max_count * a.tiles, cap = 148 * 8;
  const unsigned grid = (unsigned)(total < cap ? total : cap);
  pull_kernel<<<grid, kTileThreads, 0, st>>>(a);
  VCSMC_LAUNCH_CHECK("pull_kernel");
  return VCSMC_OK;
}

}  // namespace vcsmc
