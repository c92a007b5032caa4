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
//                        Children come from L2; bound by instruction issue around the FP64 pipe.
//   score_leaf_pairs_kernel  particles whose children are both leaves: site patterns instead of sites.
//   merge_score_mma_kernel   experiment: the same GEMM on the FP64 tensor cores (off by default, see the launcher).
//   materialise_kernel   after the next resampling, only particles that were drawn as an ancestor get their node
//                        written (the plain merge formula; 32 B per site, streaming).
//   pull_kernel          particle sharding: a rank that drew a remote ancestor copies the nodes it lacks straight out
//                        of the owner's pool over NVLink (peer pointers, 256-bit loads), one CTA per (node, site tile).
#include <stdlib.h>

#include "launch.h"
#include "merge_device.cuh"

namespace vcsmc {
namespace {

constexpr int kRScore = 32;  // particles per scoring group
constexpr int kCoef = 20;    // doubles per particle in shared memory: M[4][4] (or the 4 JC coefficients) + the column sums of M

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
  int tiles_per_item;
  int n_chunks;
  int R;
  int skip_leaf_pairs;  // particles whose children are both leaves are scored from the pattern histogram instead
  int skip_octets;      // octets of particles that share one child pair are scored by merge_score_mma_kernel instead
  int part_stride;      // partial sums per (particle, chunk): 1, or kWarps when the tensor-core experiment shares the array
  double* ell_part;  // [K][n_chunks][part_stride]
};

constexpr int kScoreSmemBytes = kRScore * kCoef * 8 + kRScore * kTileThreads * (8 + 4);

// site products of one child pair for the SPT sites of a thread: general Q -> the 16 products L_a[i] L_b[m];
// JC -> (sa sb, sa (pi.L_b), (pi.L_a) sb, sum_i pi_i L_a[i] L_b[i]).  Sites past the end get zeros.
struct ChildSpace {
  const uint8_t* codes;
  int64_t codes_stride;
  const double* pool;
  int64_t slot_sites;
  int n_sites;
  int skip_leaf_pairs;
  int skip_octets;
};

template <bool JC, int SPT, int NC>
__device__ __forceinline__ void site_products(const ChildSpace& a, int ca, int cb, int sbase, const double (&pi)[4],
                                              double (&C)[SPT][NC]) {
  const ChildRef ra = child_ref(ca, a.codes, a.codes_stride, a.pool, a.slot_sites);
  const ChildRef rb = child_ref(cb, a.codes, a.codes_stride, a.pool, a.slot_sites);
#pragma unroll
  for (int q = 0; q < SPT; ++q) {
    const int s = sbase + q * kTileThreads;
    if (s < a.n_sites) {
      const d4 La = load_child(ra, s), Lb = load_child(rb, s);
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
        C[q][0] = sa * sb;
        C[q][1] = sa * qb;
        C[q][2] = qa * sb;
        C[q][3] = qab;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int m = 0; m < 4; ++m) C[q][(i * 4 + m) % NC] = La.v[i] * Lb.v[m];
      }
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) C[q][c] = 0.0;
    }
  }
}

// fold the SPT site likelihoods x[] of one particle into its running (mantissa product, biased exponent sum)
template <int SPT>
__device__ __forceinline__ void fold_sites(const double (&x)[SPT], bool renorm, double* prod, int* expo) {
  double pr = *prod;
  int ex = *expo;
  bool odd = false;
#pragma unroll
  for (int q = 0; q < SPT; ++q) {
    const int hi = __double2hiint(x[q]);
    const unsigned e = (unsigned)hi >> 20;  // biased exponent (sign bit included: negative values are "odd")
    odd |= (e - 1u) >= 0x7feu;
    ex += (int)e;
    pr *= __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x[q]));
  }
  if (odd) pr = __longlong_as_double(0x7ff8000000000000ll);
  if (renorm) {  // keep the mantissa product far from 2^1024 (NaN stays NaN)
    const int hi = __double2hiint(pr);
    ex += (int)(((unsigned)hi >> 20) & 0x7ffu) - 1023;
    if ((((unsigned)hi >> 20) & 0x7ffu) != 0x7ffu) pr = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(pr));
  }
  *prod = pr;
  *expo = ex;
}

// NP particles (that share the current child pair) against the SPT sites of this thread: NP*SPT*2 independent FMA
// chains (each site likelihood is accumulated in an even and an odd half) -- the FP64 pipe has a long dependent-issue
// latency and only 4 warps per scheduler fit, so the instruction-level parallelism has to come from here.
template <bool JC, int SPT, int NP>
__device__ __forceinline__ void score_particles(int j, const double (&C)[SPT][JC ? 4 : 16], const double (&x0)[SPT], bool renorm,
                                                const double* sC, double* my_prod, int* my_exp) {
  constexpr int NC = JC ? 4 : 16;
  double xe[NP][SPT], xo[NP][SPT];
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
      xe[p][q] = x0[q];
      xo[p][q] = 0.0;
    }
#pragma unroll
  for (int c = 0; c < NC; c += 2) {
    double2 m[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) m[p] = *reinterpret_cast<const double2*>(sC + (j + p) * kCoef + c);
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
      for (int q = 0; q < SPT; ++q) {
        xe[p][q] = fma(m[p].x, C[q][c], xe[p][q]);
        xo[p][q] = fma(m[p].y, C[q][c + 1], xo[p][q]);
      }
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    double x[SPT];
#pragma unroll
    for (int q = 0; q < SPT; ++q) x[q] = xe[p][q] + xo[p][q];
    fold_sites<SPT>(x, renorm, my_prod + (j + p) * kTileThreads, my_exp + (j + p) * kTileThreads);
  }
}

// The same when child a is a LEAF with a one-hot (or all-ones) state mask at every site of this thread: the site
// likelihood is row `state` of M (or the column sums of M) dotted with L_b -- 4 DFMA per site instead of 16.  The row is
// fetched from shared memory per lane (offset roff = 4 * state, or 16 for a gap).
template <int SPT, int NP>
__device__ __forceinline__ void score_particles_leaf(int j, const double (&Lb)[SPT][16], const int (&roff)[SPT],
                                                     const double (&x0)[SPT], bool renorm, const double* sC, double* my_prod,
                                                     int* my_exp) {
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    double x[SPT];
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
      const double2* row = reinterpret_cast<const double2*>(sC + (j + p) * kCoef + roff[q]);
      const double2 r0 = row[0], r1 = row[1];
      const double xe = fma(r0.y, Lb[q][1], fma(r0.x, Lb[q][0], x0[q]));
      const double xo = fma(r1.y, Lb[q][3], r1.x * Lb[q][2]);
      x[q] = xe + xo;
    }
    fold_sites<SPT>(x, renorm, my_prod + (j + p) * kTileThreads, my_exp + (j + p) * kTileThreads);
  }
}

// particles j .. j+7 of the group exist and share one child pair (and are not a leaf-leaf pair when those are skipped)
__device__ __forceinline__ bool octet_uniform(const int* s_a, const int* s_b, int j, int nj) {
  if (j + 8 > nj) return false;
  const int ca = s_a[j], cb = s_b[j];
  bool u = true;
#pragma unroll
  for (int i = 1; i < 8; ++i) u = u && s_a[j + i] == ca && s_b[j + i] == cb;
  return u;
}

// One tile of SPT*256 sites for the nj particles of a group.  The running product of the site likelihoods of
// (thread, particle) is kept as (mantissa product, BIASED exponent sum) in shared memory.  The split is three integer
// ops per site; a likelihood that is not a positive normal number (0, subnormal, inf, NaN, negative) poisons the
// mantissa product with NaN and the particle is re-evaluated with one log per site afterwards (never on sane inputs).
// Sites past the end of the alignment have zero site products and x0 = 1, i.e. x = 1.
template <bool JC, int SPT>
__device__ __forceinline__ void score_tile(const ChildSpace& a, int nj, unsigned skip, int sbase, bool renorm, const double (&pi)[4],
                                           const int* s_a, const int* s_b, const double* sC, double* my_prod, int* my_exp) {
  constexpr int NC = JC ? 4 : 16;
  int pa = kNone, pb = kNone;
  double C[SPT][NC];
  double x0[SPT];
  int roff[SPT];
  bool leaf_rows = false;  // current pair = (leaf with one-hot / gap masks, internal node): use the 4-DFMA path
#pragma unroll
  for (int q = 0; q < SPT; ++q) {
    x0[q] = (sbase + q * kTileThreads < a.n_sites) ? 0.0 : 1.0;
    roff[q] = 0;
  }
  int j = 0;
  while (j < nj) {
    const int ca = s_a[j], cb = s_b[j];
    if (skip >> j & 1u) {  // a leaf pair (scored from site patterns) or part of a uniform octet (tensor-core kernel)
      ++j;
      continue;
    }
    if (ca != pa || cb != pb) {
      leaf_rows = false;
      if (!JC && ca < 0 && cb >= 0) {
        // leaf + internal node: try the row path (every site of this warp must have a one-hot or all-ones mask)
        const uint8_t* crow = a.codes + (int64_t)(-ca - 1) * a.codes_stride;
        const double* node = a.pool + (int64_t)cb * a.slot_sites * 4;
        bool ok = true;
#pragma unroll
        for (int q = 0; q < SPT; ++q) {
          const int s = sbase + q * kTileThreads;
          if (s < a.n_sites) {
            const int code = __ldg(crow + s) & 15;
            const d4 L = ld_site(node + (int64_t)s * 4);
#pragma unroll
            for (int m = 0; m < 4; ++m) C[q][m % NC] = L.v[m];
            roff[q] = code == 15 ? 16 : 4 * (__ffs(code) - 1);
            ok = ok && (code == 15 || (code & (code - 1)) == 0);
          } else {
#pragma unroll
            for (int m = 0; m < 4; ++m) C[q][m % NC] = 0.0;
            roff[q] = 0;
          }
        }
        leaf_rows = __all_sync(0xffffffffu, ok);
      }
      if (!leaf_rows) site_products<JC, SPT, NC>(a, ca, cb, sbase, pi, C);
      pa = ca;
      pb = cb;
    }
    const bool two = j + 1 < nj && !(skip >> (j + 1) & 1u) && s_a[j + 1] == ca && s_b[j + 1] == cb;
    if (!JC && leaf_rows) {
      if (two) score_particles_leaf<SPT, 2>(j, reinterpret_cast<const double(&)[SPT][16]>(C), roff, x0, renorm, sC, my_prod, my_exp);
      else score_particles_leaf<SPT, 1>(j, reinterpret_cast<const double(&)[SPT][16]>(C), roff, x0, renorm, sC, my_prod, my_exp);
    } else {
      if (two) score_particles<JC, SPT, 2>(j, C, x0, renorm, sC, my_prod, my_exp);
      else score_particles<JC, SPT, 1>(j, C, x0, renorm, sC, my_prod, my_exp);
    }
    j += two ? 2 : 1;
  }
}

// exact fallback for a particle whose product was poisoned: one log per site
template <bool JC>
__device__ __noinline__ double score_slow(ChildSpace a, int ca, int cb, const double* cj, int s_begin, int s_end, double pi0,
                                          double pi1, double pi2, double pi3) {
  constexpr int NC = JC ? 4 : 16;
  const double pi[4] = {pi0, pi1, pi2, pi3};
  double acc = 0.0;
  double C[1][NC], m[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) m[c] = cj[c];
  for (int s = s_begin + threadIdx.x; s < s_end; s += kTileThreads) {
    site_products<JC, 1, NC>(a, ca, cb, s, pi, C);
    double x = m[0] * C[0][0];
#pragma unroll
    for (int c = 1; c < NC; ++c) x = fma(m[c], C[0][c], x);
    acc += log(x);
  }
  return acc;
}

template <bool JC, int SPT>
__global__ void __launch_bounds__(kTileThreads, (JC || SPT <= 2) ? 2 : 1) merge_score_kernel(const ScoreArgs a) {
  constexpr int NC = JC ? 4 : 16;  // coefficients per particle == site products per site
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sC = reinterpret_cast<double*>(smem_raw);                    // [R][16] per-particle coefficients
  double* s_prod = sC + kRScore * kCoef;                                 // [R][256] running mantissa products
  int* s_exp = reinterpret_cast<int*>(s_prod + kRScore * kTileThreads); // [R][256] running (biased) exponent sums
  __shared__ int s_k[kRScore], s_a[kRScore], s_b[kRScore];
  __shared__ double s_slow[kWarps];
  __shared__ unsigned s_odd, s_skip;   // s_skip: particles of the group some other kernel scores (leaf pairs, uniform octets)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int R = a.R;
  const int64_t count = a.count ? (int64_t)*a.count : a.K;
  const int64_t total = ((count + R - 1) / R) * a.n_chunks;
  double pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = __ldg(a.pi + j);
  double* my_prod = s_prod + tid;
  int* my_exp = s_exp + tid;
  const ChildSpace cs = {a.codes, a.codes_stride, a.pool, a.slot_sites, a.n_sites, a.skip_leaf_pairs, a.skip_octets};

  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t g = w / a.n_chunks;
    const int tc = (int)(w - g * a.n_chunks);
    const int64_t j0 = g * R;
    const int nj = (int)min((int64_t)R, count - j0);
    __syncthreads();
    if (tid == 0) s_odd = 0u;
    if (tid < nj) {
      const int k = a.order ? a.order[j0 + tid] : (int)(j0 + tid);
      const int ls = a.lsrc[k], rs = a.rsrc[k];
      const bool sw = ls > rs;  // canonical child order a <= b; the swap flag rides in the sign of k
      s_a[tid] = sw ? rs : ls;
      s_b[tid] = sw ? ls : rs;
      s_k[tid] = sw ? ~k : k;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned skip = 0u;
      for (int j = 0; j < nj; ++j)
        if (a.skip_leaf_pairs && s_b[j] < 0) skip |= 1u << j;
      if (a.skip_octets)
        for (int o = 0; o < 4; ++o)
          if (octet_uniform(s_a, s_b, 8 * o, nj)) skip |= 0xffu << (8 * o);
      s_skip = skip;
    }
    __syncthreads();
    const unsigned skip = s_skip;
    if ((skip | (nj < 32 ? ~0u << nj : 0u)) == ~0u) continue;   // nothing of this group is ours
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
      if (skip >> j & 1u) continue;
      my_prod[j * kTileThreads] = 1.0;
      my_exp[j * kTileThreads] = 0;
    }
    __syncthreads();

    const int t_begin = tc * a.tiles_per_item;
    const int t_end = min(a.tiles, t_begin + a.tiles_per_item);
    for (int t = t_begin; t < t_end; ++t) {
      const int sbase = t * (kTileThreads * SPT) + tid;
      const bool renorm = ((t - t_begin) & 127) == 127;
      score_tile<JC, SPT>(cs, nj, skip, sbase, renorm, pi, s_a, s_b, sC, my_prod, my_exp);
    }
    // sum_s log x_s = log(prod mantissas) + ln2 * sum (exponents - bias).  The 256 per-thread products of a particle are
    // combined by ONE warp (8 entries per lane, renormalised after every multiply) before the log: 32 logs per particle
    // instead of 256.
    const int bias = 1023 * SPT * (t_end - t_begin);
    __syncthreads();
    for (int j = wid; j < nj; j += kWarps) {
      if (skip >> j & 1u) continue;  // written by score_leaf_pairs_kernel / merge_score_mma_kernel
      double p = 1.0;
      int e = 0;
#pragma unroll
      for (int i = 0; i < kTileThreads / 32; ++i) {
        p *= s_prod[j * kTileThreads + lane + 32 * i];
        e += s_exp[j * kTileThreads + lane + 32 * i] - bias;
        const int hi = __double2hiint(p);
        const int ee = (hi >> 20) & 0x7ff;
        if (ee != 0x7ff) {   // (a poisoned product stays NaN)
          p = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(p));
          e += ee - 1023;
        }
      }
      const double ef = (double)e;
      double acc = fma(ef, 6.93147180369123816490e-01, fma(ef, 1.90821492927058770002e-10, log(p)));  // ln2 hi + lo
      acc = warp_sum(acc);
      if (__any_sync(0xffffffffu, p != p)) {
        if (lane == 0) atomicOr(&s_odd, 1u << j);
      } else if (lane < a.part_stride) {
        const int kk = s_k[j];
        const int64_t k = kk < 0 ? ~kk : kk;
        a.ell_part[(k * a.n_chunks + tc) * a.part_stride + lane] = lane == 0 ? acc : 0.0;
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
      if (tid < a.part_stride) {
        const int kk = s_k[j];
        const int64_t k = kk < 0 ? ~kk : kk;
        double t = s_slow[tid];
        if (a.part_stride == 1)
          for (int w2 = 1; w2 < kWarps; ++w2) t += s_slow[w2];
        a.ell_part[(k * a.n_chunks + tc) * a.part_stride + tid] = t;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The same scoring on the FP64 tensor cores, for octets of particles that share one child pair.
//
// x[particle][site] = sum_c M_particle[c] * O_site[c] is a GEMM with inner dimension 16.  One mma.sync.m8n8k4.f64
// multiplies 8 particles (A: 8x4 coefficients, registers, loaded once per work item) by 8 sites (B: 4x8 site products,
// 4 DMUL per lane per site block) -- 256 FMA per warp instruction instead of 32.  B200's DMMA rate equals its DFMA rate
// (18.5 T FMA/s, scripts/microbench_dmma.cu): the gain is instruction issue, which is what bounds the vector kernel.
// A warp owns 64 sites of every 512-site tile of the work item; a lane ends up with the likelihoods of ONE particle
// (row lane/4) at two sites per block, folds them into a running (mantissa product, exponent sum) in registers, and the
// four lanes of a row are combined at the end.  Same per-(particle, chunk, warp) partial sums as the vector kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kTileThreads, 2) merge_score_mma_kernel(const ScoreArgs a) {
  __shared__ __align__(16) double sC[kRScore * 16];
  __shared__ int s_k[kRScore], s_a[kRScore], s_b[kRScore];
  __shared__ unsigned s_odd, s_mine;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int row = lane >> 2, q4 = lane & 3;
  const int R = a.R;
  const int64_t count = a.count ? (int64_t)*a.count : a.K;
  const int64_t total = ((count + R - 1) / R) * a.n_chunks;
  double pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = __ldg(a.pi + j);
  const ChildSpace cs = {a.codes, a.codes_stride, a.pool, a.slot_sites, a.n_sites, a.skip_leaf_pairs, 0};
  constexpr int kTile = kTileThreads * 2;  // sites per tile (the vector kernel's tile with 2 sites per thread)

  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t g = w / a.n_chunks;
    const int tc = (int)(w - g * a.n_chunks);
    const int64_t j0 = g * R;
    const int nj = (int)min((int64_t)R, count - j0);
    __syncthreads();
    if (tid == 0) s_odd = 0u;
    if (tid < nj) {
      const int k = a.order ? a.order[j0 + tid] : (int)(j0 + tid);
      const int ls = a.lsrc[k], rs = a.rsrc[k];
      const bool sw = ls > rs;
      s_a[tid] = sw ? rs : ls;
      s_b[tid] = sw ? ls : rs;
      s_k[tid] = sw ? ~k : k;
    }
    __syncthreads();
    // which octets are mine (uniform pair, not a leaf pair)
    if (tid == 0) {
      unsigned m = 0u;
      for (int o = 0; o < 4; ++o)
        if (octet_uniform(s_a, s_b, 8 * o, nj) && !(a.skip_leaf_pairs && s_b[8 * o] < 0)) m |= 1u << o;
      s_mine = m;
    }
    __syncthreads();
    const unsigned mine = s_mine;
    if (mine == 0u) continue;   // (uniform across the CTA)
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
      sC[j * 16 + ai * 4 + bi] = m;
    }
    __syncthreads();
    // A fragments: lane (row, q4) holds M_{particle 8 o + row}[4 ks + q4] for k-step ks
    double A[4][4];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) A[o][ks] = (mine >> o & 1u) ? sC[(8 * o + row) * 16 + 4 * ks + q4] : 0.0;
    double pr[4] = {1.0, 1.0, 1.0, 1.0};
    int ex[4] = {0, 0, 0, 0};

    const int t_begin = tc * a.tiles_per_item;
    const int t_end = min(a.tiles, t_begin + a.tiles_per_item);
    int folded = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const bool renorm = ((t - t_begin) & 31) == 31;
#pragma unroll 2
      for (int blk = 0; blk < 8; ++blk) {
        const int s0 = t * kTile + wid * 64 + blk * 8;   // the 8 sites of this block
        const int sB = s0 + row;                          // B fragment: lane (q4, site row)
        int pa = kNone, pb = kNone;
        double B[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          if (!(mine >> o & 1u)) continue;
          const int ca = s_a[8 * o], cb = s_b[8 * o];
          if (ca != pa || cb != pb) {
            // site products L_a[site][ks] * L_b[site][q4] for the four k-steps
            if (sB < a.n_sites) {
              const ChildRef ra = child_ref(ca, cs.codes, cs.codes_stride, cs.pool, cs.slot_sites);
              const d4 La = load_child(ra, sB);
              double lb;
              if (cb < 0) lb = (__ldg(cs.codes + (int64_t)(-cb - 1) * cs.codes_stride + sB) >> q4 & 1) ? 1.0 : 0.0;
              else lb = __ldg(cs.pool + ((int64_t)cb * cs.slot_sites + sB) * 4 + q4);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) B[ks] = La.v[ks] * lb;
            } else {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) B[ks] = 0.0;
            }
            pa = ca;
            pb = cb;
          }
          double c0 = 0.0, c1 = 0.0;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) dmma8x8x4(c0, c1, A[o][ks], B[ks]);
          // this lane: particle 8 o + row at sites s0 + 2 q4, s0 + 2 q4 + 1 (past the end: x = 1)
          double x[2] = {s0 + 2 * q4 < a.n_sites ? c0 : 1.0, s0 + 2 * q4 + 1 < a.n_sites ? c1 : 1.0};
          fold_sites<2>(x, renorm && blk == 7, &pr[o], &ex[o]);
        }
      }
      folded += 16;   // site values folded per lane and octet in this tile
    }
    // poisoned products: the particle is re-evaluated with one log per site (never on sane inputs)
    unsigned odd_mask = 0u;
#pragma unroll
    for (int o = 0; o < 4; ++o)
      if ((mine >> o & 1u) && pr[o] != pr[o]) odd_mask |= 1u << (8 * o + row);
    if (odd_mask) atomicOr(&s_odd, odd_mask);
    __syncthreads();
    const unsigned odd_all = s_odd;
    const int bias = 1023 * folded;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      if (!(mine >> o & 1u)) continue;
      const double e = (double)(ex[o] - bias);
      double acc = fma(e, 6.93147180369123816490e-01, fma(e, 1.90821492927058770002e-10, log(pr[o])));
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      const int j = 8 * o + row;
      if (q4 == 0 && !(odd_all >> j & 1u)) {
        const int kk = s_k[j];
        const int64_t k = kk < 0 ? ~kk : kk;
        a.ell_part[(k * a.n_chunks + tc) * kWarps + wid] = acc;
      }
    }
    if (odd_all) {
      for (int j = 0; j < nj; ++j) {
        if (!(odd_all >> j & 1u)) continue;
        double acc = score_slow<false>(cs, s_a[j], s_b[j], sC + j * 16, t_begin * kTile, min(a.n_sites, t_end * kTile),
                                       pi[0], pi[1], pi[2], pi[3]);
        acc = warp_sum(acc);
        if (lane == 0) {
          const int kk = s_k[j];
          const int64_t k = kk < 0 ? ~kk : kk;
          a.ell_part[(k * a.n_chunks + tc) * kWarps + wid] = acc;
        }
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
__device__ __forceinline__ int64_t leaf_pair_index(int a, int b, int N) {  // a < b
  return (int64_t)a * (2 * N - a - 1) / 2 + (b - a - 1);
}

__global__ void __launch_bounds__(256) leaf_pair_hist_kernel(const uint8_t* __restrict__ codes, int64_t stride, int N, int S,
                                                             int32_t* __restrict__ hist) {
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
  hist[(int64_t)blockIdx.x * 256 + threadIdx.x] = sh[threadIdx.x];
}

// one warp per particle: 16 lanes build M[j][m] = sum_i pi_i P_l[j][i] P_r[m][i] (the likelihood of the one-hot pattern
// (j, m)), then the lanes share the 256 patterns; a pattern with ambiguity masks sums the M entries its masks cover
__global__ void __launch_bounds__(256) score_leaf_pairs_kernel(const int32_t* __restrict__ lsrc, const int32_t* __restrict__ rsrc,
                                                               const double* __restrict__ P, const double* __restrict__ pi_,
                                                               int64_t K, int N, const int32_t* __restrict__ hist,
                                                               int n_parts, double* __restrict__ ell_part) {
  __shared__ double sM[8][16];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t k = (int64_t)blockIdx.x * 8 + wid;
  if (k >= K) return;
  const int ls = lsrc[k], rs = rsrc[k];
  if (ls >= 0 || rs >= 0) return;  // (whole warp)
  const int la = -ls - 1, lb = -rs - 1;  // leaf indices of the left / right child
  const bool sw = la > lb;
  const int32_t* h = hist + leaf_pair_index(sw ? lb : la, sw ? la : lb, N) * 256;
  int c[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) c[q] = __ldg(h + lane + 32 * q);
  if (lane < 16) {
    // rows of M follow the LOWER leaf (the histogram's first code), columns the higher one
    const double* Pa = P + k * 32 + (sw ? 16 : 0) + (lane >> 2) * 4;
    const double* Pb = P + k * 32 + (sw ? 0 : 16) + (lane & 3) * 4;
    double m = __ldg(pi_) * __ldg(Pa) * __ldg(Pb);
#pragma unroll
    for (int i = 1; i < 4; ++i) m = fma(__ldg(pi_ + i) * __ldg(Pa + i), __ldg(Pb + i), m);
    sM[wid][lane] = m;
  }
  __syncwarp();
  double acc = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (c[q] == 0) continue;
    const int bin = lane + 32 * q, ca = bin >> 4, cb = bin & 15;
    double x = 0.0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int mm = 0; mm < 4; ++mm)
        if ((ca >> jj & 1) && (cb >> mm & 1)) x += sM[wid][jj * 4 + mm];
    acc = fma((double)c[q], log(x), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) ell_part[k * n_parts] = acc;
  for (int t = 1 + lane; t < n_parts; t += 32) ell_part[k * n_parts + t] = 0.0;
}

constexpr int64_t kScoreItems = 148 * 32;  // work items wanted per launch (measured: 592 .. 18,944; work per item is uneven)

}  // namespace

int64_t leaf_pair_hist_ints(int N) { return (int64_t)N * (N - 1) / 2 * 256; }

int launch_leaf_pair_hist(const uint8_t* codes, int64_t stride, int N, int S, int32_t* hist, cudaStream_t st) {
  if (N < 2 || S <= 0) return VCSMC_OK;
  leaf_pair_hist_kernel<<<(unsigned)((int64_t)N * (N - 1) / 2), 256, 0, st>>>(codes, stride, N, S, hist);
  VCSMC_LAUNCH_CHECK("leaf_pair_hist_kernel");
  return VCSMC_OK;
}

int launch_merge_score(const uint8_t* codes, int64_t codes_stride, const double* pool, int64_t slot_sites,
                       const int32_t* lsrc, const int32_t* rsrc, const int32_t* order, const double* P, const double* pi,
                       int64_t K, const int32_t* count, int n_sites, int jc, const int32_t* leaf_hist, int n_taxa,
                       double* ell_part, int* n_parts, cudaStream_t st) {
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
  }
  const int spt = jc ? 4 : spt_general;
  ScoreArgs a;
  a.codes = codes; a.codes_stride = codes_stride; a.pool = pool; a.slot_sites = slot_sites; a.lsrc = lsrc; a.rsrc = rsrc;
  a.order = order; a.count = order ? count : nullptr; a.P = P; a.pi = pi; a.K = K; a.n_sites = n_sites; a.ell_part = ell_part;
  a.skip_leaf_pairs = leaf_hist != nullptr;
  static int use_mma = -1;
  if (use_mma < 0) {
    // VCSMC_SCORE_MMA=1 routes octets of particles that share a child pair to merge_score_mma_kernel (FP64 tensor cores).
    // Measured at 64 x 10k x 65,536: 29.6 ms per sweep against 24.5 ms for the vector kernel alone -- DMMA and DFMA share
    // one pipe on B200 (18.5 vs 18.2 T FMA/s) and the fold after each 8x8 block is the same work, so it stays off.
    const char* e = getenv("VCSMC_SCORE_MMA");
    use_mma = e ? atoi(e) != 0 : 0;
  }
  a.skip_octets = (!jc && spt == 2 && use_mma) ? 1 : 0;
  a.tiles = (n_sites + kTileThreads * spt - 1) / (kTileThreads * spt);
  static int64_t items = 0;
  if (items == 0) {
    const char* e = getenv("VCSMC_SCORE_ITEMS");   // tuning knob: work items wanted per launch
    items = e ? atoll(e) : kScoreItems;
    if (items < 1) items = kScoreItems;
  }
  // groups as large as the machine fill allows (shared children and site products are amortised over the group)
  int64_t R = (K * a.tiles) / items;
  if (R < 1) R = 1;
  if (R > kRScore) R = kRScore;
  a.R = (int)R;
  const int64_t groups = (K + R - 1) / R;
  int64_t nc = (items + groups - 1) / groups;  // split a group's tiles only when the groups cannot fill the SMs
  if (nc < 1) nc = 1;
  if (nc > a.tiles) nc = a.tiles;
  a.tiles_per_item = (int)((a.tiles + nc - 1) / nc);
  a.n_chunks = (a.tiles + a.tiles_per_item - 1) / a.tiles_per_item;
  a.part_stride = a.skip_octets ? kWarps : 1;
  if (n_parts) *n_parts = a.n_chunks * a.part_stride;
  const int64_t total = groups * a.n_chunks;
  const int64_t cap = 148 * 2 * 8;
  const unsigned grid = (unsigned)(total < cap ? total : cap);
  if (jc) merge_score_kernel<true, 4><<<grid, kTileThreads, kScoreSmemBytes, st>>>(a);
  else if (spt == 4) merge_score_kernel<false, 4><<<grid, kTileThreads, kScoreSmemBytes, st>>>(a);
  else merge_score_kernel<false, 2><<<grid, kTileThreads, kScoreSmemBytes, st>>>(a);
  VCSMC_LAUNCH_CHECK("merge_score_kernel");
  if (a.skip_octets && a.R >= 8) {
    merge_score_mma_kernel<<<grid, kTileThreads, 0, st>>>(a);
    VCSMC_LAUNCH_CHECK("merge_score_mma_kernel");
  }
  if (leaf_hist) {
    score_leaf_pairs_kernel<<<(unsigned)((K + 7) / 8), 256, 0, st>>>(lsrc, rsrc, P, pi, K, n_taxa, leaf_hist, a.n_chunks * a.part_stride, ell_part);
    VCSMC_LAUNCH_CHECK("score_leaf_pairs_kernel");
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
