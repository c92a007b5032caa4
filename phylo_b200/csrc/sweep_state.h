// Host-side state of a sweep (shared by sweep.cu: eager forward + reverse sweep, and lazy.cu: lazy / particle-sharded
// forward).  Internal to libvcsmc_b200.
#pragma once
#include <vector>

#include "launch.h"
#include "smc_device.cuh"

struct vcsmc_sweep;

namespace vcsmc {

inline int64_t align_up(int64_t x, int64_t a = 256) { return (x + a - 1) / a * a; }

struct Layout {
  int64_t off = 0;
  template <typename T>
  int64_t take(int64_t count) {
    const int64_t o = off;
    off = align_up(off + count * (int64_t)sizeof(T));
    return o;
  }
};

struct WeightArgs {
  int r, n, N, tiles;
  int64_t K;
  const double* ell_part;
  const int32_t* ids_new;
  const int32_t* cnt_new;
  const double* ldf;
  const double* lam_l;
  const double* lam_r;
  const double* b_l;
  const double* b_r;
  const double* cum_l_prev;
  const double* cum_r_prev;
  double* cum_l;
  double* cum_r;
  const double* ll_tilde;
  double* ell_node;
  double* lw;
  double* LL;
  int32_t* vminus;
  double q;
  int64_t e_off;       // ell_node index of (local) particle 0's new node: N + r*K (+ k0 under particle sharding)
  const double* qlog;  // VNCSMC: per-particle log-probability of the chosen option (vncsmc.py:315-316), else null
};

int group_particles(vcsmc_sweep* h, const int32_t* lsrc, const int32_t* rsrc, const int32_t* active, int64_t K,
                    int32_t* order_out, int32_t* count_out, cudaStream_t st, int skip_leaf_pairs = 0);
int launch_leaf_ell(const uint8_t* codes, int64_t stride, int N, int S, const double* pi, double* ell_node, cudaStream_t st);
int launch_step_weights(const WeightArgs& w, cudaStream_t st);
int launch_finalize(int N, int64_t K, const double* stats, const double* LL_last, const double* b_l, const double* b_r,
                    const double* lam_l, const double* lam_r, double ldf_root, double* llR, double* elbo, double* logz,
                    double* ess, cudaStream_t st);
double log_double_factorial_host(int m);
int sweep_forward_lazy(vcsmc_sweep* h, const uint8_t* codes, const double* lam_l, const double* lam_r, const double* Q,
                       const double* pi, cudaStream_t st);
bool use_sorted_order(int64_t K, int n_sites);

}  // namespace vcsmc

struct vcsmc_sweep {
  int N, S, jc, keep;
  int M = 0;  // VNCSMC sub-samples (0 = VCSMC)
  int64_t K;
  char* ws;
  int64_t ws_bytes;
  // modes
  bool fwd_gc;        // forward on the garbage-collected slot pool
  bool retain;        // backward reuses the forward's nodes (no recompute)
  int64_t pool_slots; // GC mode capacity
  int chunk_sites;    // backward site-chunk size (== S when retain)
  int tiles_max;
  // offsets into ws
  int64_t o_anc, o_lref, o_rref, o_nleaf, o_rempos, o_b_l, o_b_r, o_t2, o_cum_l, o_cum_r, o_lw, o_LL, o_lltilde, o_llR,
      o_vminus, o_ell_node, o_stats, o_logz, o_ess, o_elbo, o_status, o_P, o_ids[2], o_cnt[2], o_slot[2], o_cdf,
      o_u_pair, o_u_bl, o_u_br, o_u_res, o_ell_part, o_ell_new, o_lsrc, o_rsrc, o_dst, o_ldf, o_flags, o_childsum[2],
      o_Dacc[2], o_cnew, o_consumed, o_bsrc_l, o_bsrc_r, o_bsrc_g, o_bdst, o_dP, o_dpi_each, o_dQ_acc, o_dQ_each, o_dt,
      o_suf_l, o_suf_r, o_cleaf, o_pool, o_keys_in, o_keys_out, o_vals_in, o_order, o_count,
      o_sort_temp, o_order_bwd, o_count_bwd, o_order_rec, o_count_rec, o_act_bwd, o_act_rec, o_act_all, o_dirty, o_cslot,
      o_inh_ids, o_inh_cnt, o_inh_slot, o_pot, o_choice, o_qlog, o_u_cat, o_rows_all, o_nact, o_nbase, o_v_lsrc, o_v_rsrc,
      o_v_coef, o_v_t2, o_v_P, o_v_dP, o_v_dt, o_v_dQ, o_v_dpi, o_v_order, o_v_keys_in, o_v_keys_out, o_v_vals, o_v_count, o_v_temp, o_v_keep, o_v_index, o_v_scan;
  size_t v_scan = 0;
  std::vector<int64_t> pot_off;  // per rank event, offset (doubles) into the potentials
  int64_t v_batch = 0;           // virtual events per batch in the nested reverse sweep
  size_t v_temp = 0;
  const double* x_look_bl = nullptr;
  const double* x_look_br = nullptr;
  const double* x_cat = nullptr;
  size_t sort_temp = 0;
  int64_t pool_bytes;
  std::vector<int64_t> rem_off;  // per step offset (bytes) into rempos
  // uniform source
  const float* x_pair = nullptr;
  const double* x_bl = nullptr;
  const double* x_br = nullptr;
  const double* x_res = nullptr;
  uint64_t seed = 0;
  bool use_seed = true;
  // lazy forward (score every particle, materialise only the survivors) and particle sharding (lazy.cu)
  int lazy = 1;
  bool force_gc = false;
  int rank = 0, world = 1;
  int64_t Kl = 0, k0 = 0;              // particles owned by this rank: logical k0 .. k0 + Kl - 1
  char* peer_ws[vcsmc::kMaxPeers] = {nullptr};  // base of every rank's workspace mapped into this process (identical layouts)
  vcsmc_comm_fn comm = nullptr;
  void* comm_user = nullptr;
  int site_begin = 0, site_end = -1;   // site slice of the reverse sweep (particle-sharded runs shard the backward by site)
  int64_t o_loc = 0, o_slot_id = 0, o_pend = 0, o_surv = 0, o_mat_list = 0, o_fetch_e = 0, o_fetch_src = 0, o_counts = 0,
          o_pF = 0, o_pT = 0, o_pV = 0, o_pLLt = 0, o_mat_ls = 0, o_mat_rs = 0, o_pEll = 0, o_pDirect = 0, o_lsrc2 = 0, o_rsrc2 = 0, o_F0 = 0, o_live = 0, o_haskid = 0, o_gocc = 0, o_leaf_perm = 0, o_leaf_tstate = 0,
          o_lz_ids = 0, o_lz_cnt = 0, o_u_res_all = 0, o_rec = 0, o_cdf_scratch = 0, o_gtab = 0, o_gcnt = 0, o_goff = 0, o_gslot = 0, o_grank = 0, o_leaf_hist = 0, o_F[2] = {0, 0}, o_topo[2] = {0, 0}, o_vm[2] = {0, 0};
  int event_timing = 0;                // option "event_timing": CTA 0 of the event kernel stamps %globaltimer at every phase boundary
  int64_t o_ev_timing = 0;
  int sparse_bwd = 1;                  // reverse sweep: one site-parallel launch when few particles carry an adjoint (option "sparse_bwd")
  int score_streams = 1;               // the two scoring kernels of a rank event on two streams (option "score_streams")
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int force_sorted = 0;                // testing aid: grouped visiting order (and the rows kernel) even for small K
  int leaf_rows = 1;                   // score leaf + internal merges with the rows kernel on state-sorted sites (option "leaf_rows")
  int leaf_patterns = 1;               // score leaf-leaf merges from the site-pattern histogram (option "leaf_patterns")
  int64_t rec_stride = 0;              // bytes of one rank's chunk of the per-event record
  int64_t fetch_cap = 0;
  int64_t o_sig = 0;                   // int32[kMaxPeers]: epochs the peers have signalled (flag barrier over peer memory)
  int peer_sync = 1;                   // 1: barriers and the record exchange run over peer memory; 0: through the collective hook
  int64_t o_epoch_dev = 0, o_seed_dev = 0, o_model = 0;   // device-resident barrier epoch base, seed, copy of (lam_l, lam_r, Q, pi)
  bool ldf_ready = false;
  // the lazy forward as a CUDA graph: captured on the second forward of a given (codes, uniform source), replayed after
  int use_graph = 1;
  int64_t forwards = 0;
  cudaGraphExec_t fwd_graph = nullptr;
  cudaStream_t cap_stream = nullptr;
  const void* graph_key[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t graph_launches = 0;         // kernels + memsets one replay stands for (bench's launch count)
  // hook
  vcsmc_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  double scalar_share = 1.0;
  int skip_zero = 1;
  double skip_below = 5.421010862427522e-20;  // 2^-64: adjoint coefficients this small (relative to dELBO = 1) are treated as zero
  int max_chunk_sites = 0;  // testing aid: cap the backward site chunk (0 = as large as memory allows)
  // model pointers of the last forward (caller keeps them alive until backward)
  const uint8_t* codes = nullptr;
  const double* lam_l = nullptr;
  const double* lam_r = nullptr;
  const double* Q = nullptr;
  const double* pi = nullptr;
  bool forward_done = false;
  // optional per-kernel timing of the merge launches (bench.py roofline): kind 0 = forward merge,
  // 1 = recompute merge of the chunked backward, 2 = backward merge
  bool profile = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_kind;
  size_t ev_used = 0;
  int prof_begin(int kind, cudaStream_t st) {
    if (!profile) return 0;
    if (ev_used + 2 > ev.size()) {
      for (int i = 0; i < 256; ++i) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return -1;
        ev.push_back(e);
      }
    }
    ev_kind.push_back(kind);
    return cudaEventRecord(ev[ev_used++], st) == cudaSuccess ? 0 : -1;
  }
  void prof_end(cudaStream_t st) {
    if (profile) cudaEventRecord(ev[ev_used++], st);
  }
  ~vcsmc_sweep() {
    for (auto e : ev) cudaEventDestroy(e);
    if (fwd_graph) cudaGraphExecDestroy(fwd_graph);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (side_stream) cudaStreamDestroy(side_stream);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
  }

  template <typename T>
  T* p(int64_t off) const { return reinterpret_cast<T*>(ws + off); }
};

