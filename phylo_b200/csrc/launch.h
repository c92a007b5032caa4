// Internal host-side launch prototypes shared by the translation units of libvcsmc_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vcsmc {

// merge.cu
int merge_fwd_tiles(int n_sites);
int merge_ell_parts(int n_sites);  // upper bound of the partial sums per particle written by launch_merge_fwd
int launch_merge_fwd(const uint8_t* codes, int64_t codes_stride, double* pool, int64_t slot_sites, const int32_t* lsrc,
                     const int32_t* rsrc, const int32_t* dst, const int32_t* order, const int32_t* count,
                     const double* P, const double* pi, int64_t K, int64_t n_active, int n_sites, int jc,
                     int skip_unstored, double* ell_part, int* n_parts, cudaStream_t st);
int launch_ell_reduce(const double* ell_part, int n_part, int64_t K, double* ell, cudaStream_t st);
int launch_merge_bwd(const uint8_t* codes, int64_t codes_stride, const double* pool, double* gpool, int64_t slot_sites,
                     const int32_t* lsrc, const int32_t* rsrc, const int32_t* gsrc, const int32_t* order,
                     const int32_t* count, const double* P, const double* pi, const double* coef, int64_t K,
                     int64_t n_active, int n_sites, int jc, int skip_zero, double skip_below, double* dP, double* dpi_acc, cudaStream_t st);
int launch_bwd_sparse(const uint8_t* codes, int64_t codes_stride, double* pool, double* gpool, int64_t slot_sites, int n_sites,
                      int N, int64_t K, int recompute, int jc, int skip_zero, double skip_below, const int32_t* order,
                      const int32_t* count, const int32_t* lsrc, const int32_t* rsrc, const int32_t* gsrc, const int32_t* dst,
                      const double* P, const double* pi, const double* coef, double* dP, double* dpi_acc, cudaStream_t st);

// score.cu (lazy forward: likelihood-only scoring, survivor materialisation, peer pulls)
constexpr int kLeafPairInts = 512;   // ints per leaf pair in the site-pattern table (see leaf_pair_hist_kernel)
int64_t leaf_pair_hist_ints(int N);
int launch_leaf_pair_hist(const uint8_t* codes, int64_t stride, int N, int S, int32_t* hist, cudaStream_t st);
int leaf_sort_stride(int n_sites);
int launch_leaf_sort(const uint8_t* codes, int64_t stride, int N, int S, int32_t* perm, uint8_t* tstate, cudaStream_t st);
int launch_merge_score(const uint8_t* codes, int64_t codes_stride, const double* pool, int64_t slot_sites,
                       const int32_t* lsrc, const int32_t* rsrc, const int32_t* order, const double* P, const double* pi,
                       int64_t K, const int32_t* count, int n_sites, int jc, int skip_leaf_pairs,
                       const int32_t* leaf_perm, const uint8_t* leaf_tstate, double* ell_part, int* n_parts, cudaStream_t st,
                       cudaStream_t st_generic = nullptr);
int launch_materialise(const uint8_t* codes, int64_t codes_stride, double* pool, int64_t slot_sites, const int32_t* lsrc,
                       const int32_t* rsrc, const int32_t* list, const int32_t* count, int64_t max_count, const int32_t* loc,
                       int64_t e_base, const double* P, int n_sites, int jc, cudaStream_t st);
int launch_pull(const int32_t* fetch_e, const int32_t* fetch_src, const int32_t* count, int64_t max_count, const int32_t* loc,
                double* pool, int64_t slot_sites, int n_sites, int world, const int32_t* const* peer_loc,
                const double* const* peer_pool, cudaStream_t st);

// sort.cu
size_t sort_temp_bytes(int64_t K);
size_t scan_temp_bytes(int64_t n);
int launch_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* temp, size_t temp_bytes, cudaStream_t st);
int launch_sort_order(const int32_t* lsrc, const int32_t* rsrc, const int32_t* active, int64_t K, int64_t max_slot, uint64_t* keys_in,
                      uint64_t* keys_out, int32_t* vals_in, int32_t* order_out, int32_t* count_out, void* temp,
                      size_t temp_bytes, cudaStream_t st);

int64_t group_table_entries(int64_t K);
int launch_group_order(const int32_t* lsrc, const int32_t* rsrc, const int32_t* active, int skip_leaf_pairs, int64_t K, unsigned long long* tab,
                       int32_t* cnt, int32_t* off, int32_t* gslot, int32_t* grank, int32_t* order_out, int32_t* count_out,
                       void* temp, size_t temp_bytes, cudaStream_t st);

// transition.cu
int launch_transition_fwd(const double* Q, const double* t, int64_t n, int jc, double* P, cudaStream_t st);
int launch_transition_bwd(const double* Q, const double* t, const double* dP, int64_t n, int jc, const int32_t* list,
                          double* dt, double* dQ_each, cudaStream_t st);

// loader.cu
int launch_pack_alignment(const double* genome, int N, int S, uint8_t* codes, int* status, cudaStream_t st);
int launch_gather_sites(const uint8_t* codes, int N, int S, const int32_t* site_idx, int n_sel, uint8_t* out,
                        cudaStream_t st);

// smc.cu
int launch_propose_pairs(const float* u, int64_t K, int n, int32_t* coal, int32_t* rem, cudaStream_t st);
int64_t resample_scratch_doubles(int64_t K);
int launch_resample_cdf(const double* lw, int64_t K, double* cdf, double* stats /*[4]: lse,total,ess,max*/, double* scratch, cudaStream_t st);
int launch_resample_search(const double* cdf, const double* stats, const double* u, int64_t K, int32_t* idx, cudaStream_t st);
int launch_philox_step(uint64_t seed, int r, int64_t k0, int64_t K, int n, float* u_pair, double* u_bl, double* u_br,
                       double* u_res, double* u_cat, cudaStream_t st, const uint64_t* seed_dev = nullptr);

// nested.cu (VNCSMC look-ahead)
int nested_max_roots();
int launch_nested_inherit(int r, int n, int N, int64_t K, const double* cdf, const double* u_res, const int32_t* ids_prev,
                          const int32_t* cnt_prev, const int32_t* slot_prev, int32_t* ids, int32_t* cnt, int32_t* slot,
                          int32_t* rows_all, const double* LL_prev, int32_t* anc, double* ll_tilde, cudaStream_t st);
int launch_lookahead(int r, int n, int N, int M, int jc, int gc, int S, int64_t K, const int32_t* ids, const int32_t* cnt,
                     const int32_t* slot, const uint8_t* codes, int64_t codes_stride, const double* pool, int64_t slot_sites,
                     const double* ell_node, const double* ldf, const double* Q, const double* pi, const double* lam_l,
                     const double* lam_r, const double* u_bl, const double* u_br, uint64_t seed, double* pot,
                     double share, cudaStream_t st);
int launch_nested_choose(int r, int n, int N, int M, int gc, int64_t K, double* pot, const double* u_cat, const double* u_bl,
                         const double* u_br, uint64_t seed, const double* lam_l, const double* lam_r, const int32_t* ids,
                         const int32_t* cnt, const int32_t* slot, int32_t* ids_new, int32_t* cnt_new, int32_t* slot_new,
                         int32_t* lref, int32_t* rref, int32_t* nleaf, uint8_t* rempos, int32_t* choice, double* b_l,
                         double* b_r, double* t2, double* qlog, int32_t* lsrc, int32_t* rsrc, int32_t* dst, cudaStream_t st);
int launch_nested_active(int r, int64_t K, int skip_zero, double rel, const double* lw, const double* stats, int32_t* active, cudaStream_t st);
int launch_nested_mark_roots(int n, int N, int64_t K, const int32_t* active, const int32_t* rows, int32_t* consumed, cudaStream_t st);
int launch_nested_coef(int r, int n, int N, int M, int64_t K, double grad, const double* lw, const double* stats,
                       const double* pot, const int32_t* choice, const int32_t* anc, double* Dacc_next, cudaStream_t st);
int launch_nested_keep(int r, int n, int M, int64_t K, double grad, int dense, double thresh, const double* lw, const double* stats,
                       const double* pot, const int32_t* choice, int32_t* keep, cudaStream_t st);
int launch_nested_virtual(int r, int n, int N, int M, int64_t K, double grad, int dense, double thresh, const double* lw, const double* stats,
                          const double* pot, const int32_t* choice, const int32_t* index, const int32_t* rows,
                          const int32_t* slot_of, const double* u_bl, const double* u_br, uint64_t seed, const double* lam_l,
                          const double* lam_r, int64_t v0, int64_t v1, int32_t* v_lsrc, int32_t* v_rsrc, double* v_coef,
                          double* v_t2, cudaStream_t st);
int launch_nested_reduce(int r, int64_t V, int jc, const double* dt, const double* dQ_each, const double* t2,
                         const double* lam_l, const double* lam_r, double* dlam_l, double* dlam_r, double* dQ,
                         cudaStream_t st);

}  // namespace vcsmc
