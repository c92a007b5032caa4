// (a) alignment loader: pack the one-hot / all-ones [N,S,4] float64 genome of runner.py:107-115 into one
// 4-bit state mask per (taxon, site), once, instead of the reference's K-fold host replication (vcsmc.py:479).
#include "common.cuh"
#include "launch.h"

namespace vcsmc {
namespace {

__global__ void pack_alignment_kernel(const double* __restrict__ genome, int64_t n, uint8_t* __restrict__ codes,
                                      int* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const d4 v = ld_site(genome + i * 4);
  unsigned code = 0;
  bool bad = false;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (v.v[a] == 1.0) code |= 1u << a;
    else if (v.v[a] != 0.0) bad = true;
  }
  if (code == 0) bad = true;  // all-zero row: log(0) in the reference (spikeGP.p)
  if (bad) atomicExch(status, VCSMC_ERR_DATA);
  codes[i] = (uint8_t)code;
}

__global__ void gather_sites_kernel(const uint8_t* __restrict__ codes, int N, int S, const int32_t* __restrict__ idx,
                                    int n_sel, uint8_t* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (j >= n_sel || t >= N) return;
  out[(int64_t)t * n_sel + j] = codes[(int64_t)t * S + idx[j]];
}

}  // namespace

int launch_pack_alignment(const double* genome, int N, int S, uint8_t* codes, int* status, cudaStream_t st) {
  const int64_t n = (int64_t)N * S;
  if (n <= 0) return VCSMC_OK;
  pack_alignment_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(genome, n, codes, status);
  VCSMC_LAUNCH_CHECK("pack_alignment_kernel");
  return VCSMC_OK;
}

int launch_gather_sites(const uint8_t* codes, int N, int S, const int32_t* site_idx, int n_sel, uint8_t* out,
                        cudaStream_t st) {
  if (n_sel <= 0 || N <= 0) return VCSMC_OK;
  dim3 grid((n_sel + 255) / 256, N);
  gather_sites_kernel<<<grid, 256, 0, st>>>(codes, N, S, site_idx, n_sel, out);
  VCSMC_LAUNCH_CHECK("gather_sites_kernel");
  return VCSMC_OK;
}

}  // namespace vcsmc
