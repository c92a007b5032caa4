// (b) batched per-particle 4x4 transition matrices P = expm(t Q) and their adjoint.
//
// Replaces tf.linalg.expm on [K,4,4] (vcsmc.py:181-184).  JC (vcsmc.py:126-129) has the closed form
// P_ii = 1/4 + 3/4 e^-t, P_ij = 1/4 - 1/4 e^-t.  The reference's "GTR" Q (vcsmc.py:138-148) is a general
// non-reversible rate matrix (complex eigenvalues are common, the initial Q has a triple eigenvalue), so the
// general path is a scaling-and-squaring Taylor series: Q is the same for every matrix of a launch, so the powers
// Q^k / k! are tabulated once and a matrix is a Horner scheme in its scalar t (common.cuh::m4_expm_tq).  The adjoint uses the Frechet derivative identity
// L*(A, G) = L(A^T, G) evaluated with the same series on the block matrix [[A^T, G], [0, A^T]].
#include "common.cuh"
#include "launch.h"

namespace vcsmc {
namespace {

__global__ void transition_fwd_kernel(const double* __restrict__ Q, const double* __restrict__ t, int64_t n, int jc,
                                      double* __restrict__ P) {
  __shared__ __align__(16) double s_tab[kExpmTableDoubles];   // Q^k / k!: the table of m4_expm_tq (one Q, many t)
  if (!jc) {
    if (threadIdx.x == 0) expm_tq_table(Q, s_tab);
    __syncthreads();
  }
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ti = t[i];
  double* out = P + i * 16;
  if (jc) {
    const double o = -0.25 * expm1(-ti);
    const double d = 0.25 + 0.75 * exp(-ti);
#pragma unroll
    for (int e = 0; e < 16; ++e) out[e] = (e % 5 == 0) ? d : o;
    return;
  }
  const M4 X = m4_expm_tq(s_tab, ti);   // (the same bits as the event kernel's transition matrices)
#pragma unroll
  for (int e = 0; e < 16; ++e) out[e] = X.a[e];
}

// `list` (optional): matrix i of the launch is matrix 2*list[i/2] + (i&1) of t / dP (the two matrices of a listed
// particle); results are written compactly at i.
__global__ void transition_bwd_kernel(const double* __restrict__ Q, const double* __restrict__ t,
                                      const double* __restrict__ dP, int64_t n, int jc, const int32_t* __restrict__ list,
                                      double* __restrict__ dt, double* __restrict__ dQ_each) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t src = list ? 2 * (int64_t)list[i >> 1] + (i & 1) : i;
  const double ti = t[src];
  const double* G = dP + src * 16;
  if (jc) {
    // compressed adjoint: G[0] = sum_i dP_ii, G[1] = sum_{i!=j} dP_ij;  d' = -3/4 e^-t, o' = 1/4 e^-t
    const double e = exp(-ti);
    dt[i] = e * (0.25 * G[1] - 0.75 * G[0]);
    return;
  }
  {
    // a zero adjoint (the rule with ESS ~ 1: almost no particle of a rank event carries weight) needs no series
    double any = 0.0;
#pragma unroll
    for (int e = 0; e < 16; ++e) any += fabs(G[e]);
    if (any == 0.0) {
      dt[i] = 0.0;
      if (dQ_each) {
#pragma unroll
        for (int e = 0; e < 16; ++e) dQ_each[i * 16 + e] = 0.0;
      }
      return;
    }
  }
  M4 At, E, X, Y;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) At.a[r * 4 + c] = __ldg(Q + c * 4 + r) * ti;
#pragma unroll
  for (int e = 0; e < 16; ++e) E.a[e] = G[e];
  m4_expm_frechet(At, E, X, Y);
  double s = 0.0;
#pragma unroll
  for (int e = 0; e < 16; ++e) s = fma(Y.a[e], __ldg(Q + e), s);
  dt[i] = s;
  if (dQ_each) {
#pragma unroll
    for (int e = 0; e < 16; ++e) dQ_each[i * 16 + e] = ti * Y.a[e];
  }
}

}  // namespace

int launch_transition_fwd(const double* Q, const double* t, int64_t n, int jc, double* P, cudaStream_t st) {
  if (n <= 0) return VCSMC_OK;
  transition_fwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(Q, t, n, jc, P);
  VCSMC_LAUNCH_CHECK("transition_fwd_kernel");
  return VCSMC_OK;
}

int launch_transition_bwd(const double* Q, const double* t, const double* dP, int64_t n, int jc, const int32_t* list,
                          double* dt, double* dQ_each, cudaStream_t st) {
  if (n <= 0) return VCSMC_OK;
  transition_bwd_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(Q, t, dP, n, jc, list, dt, dQ_each);
  VCSMC_LAUNCH_CHECK("transition_bwd_kernel");
  return VCSMC_OK;
}

}  // namespace vcsmc
