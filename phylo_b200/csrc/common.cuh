// Shared device helpers for libvcsmc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vcsmc_b200.h"

namespace vcsmc {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_cuda(cudaError_t e, const char* what);
#define VCSMC_CUDA(call)                                  \
  do {                                                    \
    int _rc = ::vcsmc::check_cuda((call), #call);         \
    if (_rc != 0) return _rc;                             \
  } while (0)
bool debug_sync();  // VCSMC_SYNC_CHECK=1: synchronise after every launch so that a faulting kernel is named (debugging aid)
#define VCSMC_LAUNCH_CHECK(name)                          \
  do {                                                    \
    ::vcsmc::count_launch();                              \
    int _rc = ::vcsmc::check_cuda(cudaGetLastError(), name); \
    if (_rc != 0) return _rc;                             \
    if (::vcsmc::debug_sync()) {                          \
      _rc = ::vcsmc::check_cuda(cudaDeviceSynchronize(), name); \
      if (_rc != 0) return _rc;                           \
    }                                                     \
  } while (0)

constexpr int kMaxPeers = 8;         // ranks of one NVLink domain a sweep can address (particle sharding)
constexpr int kTileThreads = 256;    // threads per merge CTA
constexpr int kSitesPerThread = 4;   // sites per thread per tile
constexpr int kTileSites = kTileThreads * kSitesPerThread;  // 1024 sites per (particle, tile) work item

// ---------------------------------------------------------------------------------------------
// 256-bit global access (LDG.E.256 / STG.E.256 on sm_100a): one site = 4 doubles = one access
// ---------------------------------------------------------------------------------------------
struct __align__(32) d4 {
  double v[4];
};

__device__ __forceinline__ d4 ld_site(const double* p) {
  d4 r;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_site(double* p, const d4& r) {
  asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]), "l"(p)
               : "memory");
}
// leaf state mask (bit a set <=> state a compatible) -> 0/1 partials
__device__ __forceinline__ d4 leaf_site(uint8_t code) {
  d4 r;
#pragma unroll
  for (int a = 0; a < 4; ++a) r.v[a] = (code >> a) & 1 ? 1.0 : 0.0;
  return r;
}

// ---------------------------------------------------------------------------------------------
// reductions (fixed order => deterministic)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_down_sync(0xffffffffu, x, o));
  return x;
}
// sum over a CTA of NT threads; result valid in thread 0. `sm` needs NT/32 doubles.
template <int NT>
__device__ __forceinline__ double block_sum(double x, double* sm) {
  x = warp_sum(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = x;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) t += sm[i];
  }
  return t;
}

// ---------------------------------------------------------------------------------------------
// 4x4 matrix helpers (registers)
// ---------------------------------------------------------------------------------------------
struct M4 {
  double a[16];
};
__host__ __device__ __forceinline__ M4 m4_eye() {
  M4 r;
#pragma unroll
  for (int i = 0; i < 16; ++i) r.a[i] = (i % 5 == 0) ? 1.0 : 0.0;
  return r;
}
__host__ __device__ __forceinline__ M4 m4_mul(const M4& x, const M4& y) {
  M4 r;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = x.a[i * 4] * y.a[j];
#pragma unroll
      for (int k = 1; k < 4; ++k) s = fma(x.a[i * 4 + k], y.a[k * 4 + j], s);
      r.a[i * 4 + j] = s;
    }
  return r;
}
__device__ __forceinline__ double m4_norm1(const M4& x) {
  double m = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += fabs(x.a[i * 4 + j]);
    m = fmax(m, s);
  }
  return m;
}

constexpr int kTaylorDegree = 18;  // ||A||_1 <= 1/2 after scaling: remainder 0.5^19/19! ~ 1.6e-23

__host__ __device__ __forceinline__ int expm_scale(double norm) {
  // smallest s >= 0 with norm / 2^s <= 1/2
  int s = 0;
  if (norm > 0.5) {
    int e;
    frexp(norm, &e);   // norm = m * 2^e, m in [0.5,1)
    s = e + 1;         // norm / 2^(e+1) = m/2 in [0.25, 0.5)
    if (s > 1000) s = 1000;
  }
  return s;
}

// expm(A) by scaling-and-squaring on the degree-14 Taylor polynomial (||B||_1 <= 1/2: remainder 0.5^15/15! ~ 2e-17),
// evaluated in Paterson-Stockmeyer form on I, B, B^2, B^3 and powers of B^4: 6 matrix products instead of the 14 of a
// Horner scheme.  Stands in for tf.linalg.expm (vcsmc.py:183-184).
__device__ __forceinline__ M4 m4_expm(const M4& A) {
  const int s = expm_scale(m4_norm1(A));
  const double sc = ldexp(1.0, -s);
  M4 B;
#pragma unroll
  for (int i = 0; i < 16; ++i) B.a[i] = A.a[i] * sc;
  const M4 B2 = m4_mul(B, B), B3 = m4_mul(B2, B), B4 = m4_mul(B2, B2);
  // 1/k!, k = 0..14
  constexpr double c[15] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
                            1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0, 1.0 / 87178291200.0};
  M4 X;   // c12 I + c13 B + c14 B^2
#pragma unroll
  for (int i = 0; i < 16; ++i) X.a[i] = fma(c[14], B2.a[i], fma(c[13], B.a[i], (i % 5 == 0) ? c[12] : 0.0));
#pragma unroll
  for (int blk = 2; blk >= 0; --blk) {   // X <- (c_{4b} I + c_{4b+1} B + c_{4b+2} B^2 + c_{4b+3} B^3) + B^4 X
    const M4 T = m4_mul(B4, X);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      X.a[i] = fma(c[4 * blk + 3], B3.a[i], fma(c[4 * blk + 2], B2.a[i], fma(c[4 * blk + 1], B.a[i], T.a[i] + ((i % 5 == 0) ? c[4 * blk] : 0.0))));
  }
  for (int j = 0; j < s; ++j) X = m4_mul(X, X);
  return X;
}

// expm(t Q) for a FIXED Q and many t: with the matrices Q^k / k! (k = 0..14) tabulated once (expm_tq_table), the scaled
// Taylor polynomial is a Horner scheme in the scalar t / 2^s on 16 independent entries -- 14 x 16 FMA and no matrix
// product -- followed by the s squarings.  Same polynomial and the same scaling rule as m4_expm (||t Q||_1 / 2^s <= 1/2).
// table: [15][16] matrices, then ||Q||_1.
constexpr int kExpmTableDoubles = 15 * 16 + 4;
__host__ __device__ inline void expm_tq_table(const double* Q, double* table) {
  double cur[16];
  for (int i = 0; i < 16; ++i) cur[i] = table[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int k = 1; k <= 14; ++k) {
    double nxt[16];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double v = 0.0;
        for (int m = 0; m < 4; ++m) v = fma(cur[i * 4 + m], Q[m * 4 + j], v);
        nxt[i * 4 + j] = v / (double)k;
      }
    for (int i = 0; i < 16; ++i) cur[i] = table[k * 16 + i] = nxt[i];
  }
  double nrm = 0.0;
  for (int j = 0; j < 4; ++j) {
    double c = 0.0;
    for (int i = 0; i < 4; ++i) c += fabs(Q[i * 4 + j]);
    nrm = fmax(nrm, c);
  }
  table[15 * 16] = nrm;
}
__host__ __device__ __forceinline__ M4 m4_expm_tq(const double* __restrict__ table, double t) {
  const int s = expm_scale(fabs(t) * table[15 * 16]);
  const double ts = ldexp(t, -s);
  M4 X;
#pragma unroll
  for (int i = 0; i < 16; ++i) X.a[i] = table[14 * 16 + i];
#pragma unroll 1
  for (int k = 13; k >= 0; --k) {
#pragma unroll
    for (int i = 0; i < 16; ++i) X.a[i] = fma(X.a[i], ts, table[k * 16 + i]);
  }
  for (int j = 0; j < s; ++j) X = m4_mul(X, X);
  return X;
}

// Frechet derivative L(A, E) = d/de expm(A + eE) at e=0, by the same series on the block matrix
// [[A, E], [0, A]]: pairs (X, Y) with (X1,Y1)(X2,Y2) = (X1X2, X1Y2 + Y1X2).
__device__ __forceinline__ void m4_expm_frechet(const M4& A, const M4& E, M4& X, M4& Y) {
  const int s = expm_scale(m4_norm1(A));
  const double sc = ldexp(1.0, -s);
  M4 B, F;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    B.a[i] = A.a[i] * sc;
    F.a[i] = E.a[i] * sc;
  }
  X = m4_eye();
#pragma unroll
  for (int i = 0; i < 16; ++i) Y.a[i] = 0.0;
  for (int k = kTaylorDegree; k >= 1; --k) {
    // (X, Y) <- (I, 0) + (B, F)(X, Y)/k
    M4 BX = m4_mul(B, X), BY = m4_mul(B, Y), FX = m4_mul(F, X);
    const double ik = 1.0 / (double)k;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      X.a[i] = ((i % 5 == 0) ? 1.0 : 0.0) + BX.a[i] * ik;
      Y.a[i] = (BY.a[i] + FX.a[i]) * ik;
    }
  }
  for (int j = 0; j < s; ++j) {
    M4 XY = m4_mul(X, Y), YX = m4_mul(Y, X), XX = m4_mul(X, X);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      Y.a[i] = XY.a[i] + YX.a[i];
      X.a[i] = XX.a[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (c0,c1,c2,c3), key = (k0,k1)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}
__host__ __device__ __forceinline__ float u32_to_unit_f32(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }
__host__ __device__ __forceinline__ double u64_to_unit_f64(uint32_t hi, uint32_t lo) {
  const uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)x * 1.1102230246251565e-16;  // 2^-53
}

}  // namespace vcsmc
