// extern "C" surface of libvcsmc_b200 (kernel-level entry points) + error plumbing.  See include/vcsmc_b200.h.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "launch.h"

namespace vcsmc {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool debug_sync() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VCSMC_SYNC_CHECK");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return VCSMC_OK;
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return VCSMC_ERR_CUDA;
}

}  // namespace vcsmc

using namespace vcsmc;

extern "C" {

int vcsmc_abi_version(void) { return VCSMC_ABI_VERSION; }
const char* vcsmc_last_error(void) { return g_err; }
uint64_t vcsmc_launch_count(void) { return g_launches.load(); }

int vcsmc_pack_alignment(const double* genome, int n_taxa, int n_sites, uint8_t* codes, int* status, void* stream) {
  if (!genome || !codes || !status || n_taxa < 1 || n_sites < 1) { set_error("pack_alignment: bad argument"); return VCSMC_ERR_ARG; }
  return launch_pack_alignment(genome, n_taxa, n_sites, codes, status, (cudaStream_t)stream);
}

int vcsmc_gather_sites(const uint8_t* codes, int n_taxa, int n_sites, const int32_t* site_idx, int n_sel, uint8_t* out, void* stream) {
  if (!codes || !site_idx || !out || n_taxa < 1 || n_sites < 1 || n_sel < 0) { set_error("gather_sites: bad argument"); return VCSMC_ERR_ARG; }
  return launch_gather_sites(codes, n_taxa, n_sites, site_idx, n_sel, out, (cudaStream_t)stream);
}

int vcsmc_transition_host(const double* Q, const double* t, int64_t n, double* P) {
  if (!Q || !t || !P || n < 0) { set_error("transition_host: bad argument"); return VCSMC_ERR_ARG; }
  double table[kExpmTableDoubles];
  expm_tq_table(Q, table);
  for (int64_t i = 0; i < n; ++i) {
    const M4 X = m4_expm_tq(table, t[i]);
    for (int e = 0; e < 16; ++e) P[i * 16 + e] = X.a[e];
  }
  return VCSMC_OK;
}

int vcsmc_transition_fwd(const double* Q, const double* t, int64_t n, int jc, double* P, void* stream) {
  if ((!jc && !Q) || !t || !P || n < 0) { set_error("transition_fwd: bad argument"); return VCSMC_ERR_ARG; }
  return launch_transition_fwd(Q, t, n, jc, P, (cudaStream_t)stream);
}

int vcsmc_transition_bwd(const double* Q, const double* t, const double* dP, int64_t n, int jc, double* dt, double* dQ_each, void* stream) {
  if ((!jc && !Q) || !t || !dP || !dt || n < 0) { set_error("transition_bwd: bad argument"); return VCSMC_ERR_ARG; }
  return launch_transition_bwd(Q, t, dP, n, jc, nullptr, dt, dQ_each, (cudaStream_t)stream);
}

int vcsmc_merge_tiles(int n_sites) { return merge_ell_parts(n_sites); }
int64_t vcsmc_resample_work_doubles(int64_t K) { return K + 4 + resample_scratch_doubles(K); }

int vcsmc_merge_fwd(const uint8_t* codes, int64_t codes_stride, double* pool, int64_t slot_sites, const int32_t* lsrc,
                    const int32_t* rsrc, const int32_t* dst, const double* P, const double* pi, int64_t K, int n_sites,
                    int jc, double* ell_part, double* ell, void* stream) {
  if (!lsrc || !rsrc || !P || !pi || !ell_part || K < 0 || n_sites < 0 || slot_sites < n_sites) { set_error("merge_fwd: bad argument"); return VCSMC_ERR_ARG; }
  int n_parts = 0;
  int rc = launch_merge_fwd(codes, codes_stride, pool, slot_sites, lsrc, rsrc, dst, nullptr, nullptr, P, pi, K, -1, n_sites, jc, 0, ell_part, &n_parts, (cudaStream_t)stream);
  if (rc || !ell || K == 0 || n_sites == 0) return rc;
  return launch_ell_reduce(ell_part, n_parts, K, ell, (cudaStream_t)stream);
}

int vcsmc_merge_bwd(const uint8_t* codes, int64_t codes_stride, const double* pool, double* gpool, int64_t slot_sites,
                    const int32_t* lsrc, const int32_t* rsrc, const int32_t* gsrc, const double* P, const double* pi,
                    const double* coef, int64_t K, int n_sites, int jc, double* dP, double* dpi, void* stream) {
  if (!lsrc || !rsrc || !P || !pi || !coef || !dP || K < 0 || n_sites < 0 || slot_sites < n_sites) { set_error("merge_bwd: bad argument"); return VCSMC_ERR_ARG; }
  return launch_merge_bwd(codes, codes_stride, pool, gpool, slot_sites, lsrc, rsrc, gsrc, nullptr, nullptr, P, pi, coef, K, -1, n_sites, jc, 0, 0.0, dP, dpi, (cudaStream_t)stream);
}

int vcsmc_propose_pairs(const float* u, int64_t K, int n, int32_t* coal, int32_t* rem, void* stream) {
  if (!u || !coal || (!rem && n > 2) || K < 0) { set_error("propose_pairs: bad argument"); return VCSMC_ERR_ARG; }
  return launch_propose_pairs(u, K, n, coal, rem, (cudaStream_t)stream);
}

int vcsmc_resample(const double* lw, const double* u, int64_t K, int32_t* idx, double* lse, double* ess, double* work, void* stream) {
  if (!lw || !u || !idx || !work || K < 1) { set_error("resample: bad argument"); return VCSMC_ERR_ARG; }
  // work: K doubles of CDF, 4 doubles of statistics, then the per-tile partials of the multi-CTA scan
  double* stats = work + K;
  int rc = launch_resample_cdf(lw, K, work, stats, stats + 4, (cudaStream_t)stream);
  if (rc) return rc;
  rc = launch_resample_search(work, stats, u, K, idx, (cudaStream_t)stream);
  if (rc) return rc;
  if (lse) VCSMC_CUDA(cudaMemcpyAsync(lse, stats, sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  if (ess) VCSMC_CUDA(cudaMemcpyAsync(ess, stats + 2, sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return VCSMC_OK;
}

int vcsmc_ipc_export(void* dev_ptr, void* handle_host, int64_t* offset_host) {
  if (!dev_ptr || !handle_host || !offset_host) { set_error("ipc_export: null argument"); return VCSMC_ERR_ARG; }
  cudaIpcMemHandle_t hd;
  VCSMC_CUDA(cudaIpcGetMemHandle(&hd, dev_ptr));
  memcpy(handle_host, &hd, sizeof(hd));
  // the handle names the whole allocation dev_ptr lies in: report where inside it dev_ptr is
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  VCSMC_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
  if (!fn || qr != cudaDriverEntryPointSuccess) { set_error("ipc_export: cuMemGetAddressRange unavailable"); return VCSMC_ERR_CUDA; }
  unsigned long long base = 0;
  size_t size = 0;
  if (((range_fn)fn)(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0) { set_error("ipc_export: cuMemGetAddressRange failed"); return VCSMC_ERR_CUDA; }
  *offset_host = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
  return VCSMC_OK;
}

int vcsmc_ipc_open(const void* handle_host, void** base_out) {
  if (!handle_host || !base_out) { set_error("ipc_open: null argument"); return VCSMC_ERR_ARG; }
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle_host, sizeof(hd));
  VCSMC_CUDA(cudaIpcOpenMemHandle(base_out, hd, cudaIpcMemLazyEnablePeerAccess));
  return VCSMC_OK;
}

int vcsmc_ipc_close(void* base) {
  if (!base) return VCSMC_OK;
  VCSMC_CUDA(cudaIpcCloseMemHandle(base));
  return VCSMC_OK;
}

int vcsmc_philox_step_uniforms(uint64_t seed, int r, int64_t k0, int64_t K, int n, float* u_pair, double* u_bl, double* u_br, double* u_res, void* stream) {
  if (K < 0 || n < 0 || r < 0) { set_error("philox: bad argument"); return VCSMC_ERR_ARG; }
  return launch_philox_step(seed, r, k0, K, n, u_pair, u_bl, u_br, u_res, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
