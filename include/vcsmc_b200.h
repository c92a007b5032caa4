/*
 * vcsmc_b200.h -- C ABI of the B200-native VCSMC hot path (libvcsmc_b200.so).
 *
 * The reference (amoretti86/phylo) has NO FFI / plugin interface: the path sits behind the Python
 * class VCSMC (vcsmc.py:103-131) and its TensorFlow graph.  The entry points below are therefore
 * cut at the reference's own function boundaries; each one cites the reference code it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless its name ends in _host;
 *     the library never allocates or frees caller-visible memory (workspaces are sized by a
 *     query and passed in);
 *   - every launch is asynchronous and ordered on the cudaStream_t passed as `void* stream`;
 *   - return value 0 = OK, negative = error (vcsmc_last_error() gives a thread-local message);
 *   - no C++ exception crosses this boundary; there is no CPU fallback: without a CUDA device
 *     every compute entry returns VCSMC_ERR_CUDA;
 *   - all arithmetic is IEEE float64, matching the reference (tf.float64 everywhere except the
 *     float32 pair-proposal uniforms, vcsmc.py:301-303); alphabet size A = 4;
 *   - matrices are row-major P[i*4+j]; partial-likelihood vectors are [site][4] (the reference's
 *     [S,A] layout), 32-byte aligned.
 *   - node reference ("ref", int32): 0..N-1 = leaf i; N + r*K + k = the node created by particle
 *     slot k at rank event r.
 */
#ifndef VCSMC_B200_H
#define VCSMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCSMC_ABI_VERSION 4

#define VCSMC_OK 0
#define VCSMC_ERR_ARG (-1)      /* bad argument */
#define VCSMC_ERR_CUDA (-2)     /* CUDA runtime error (or no device) */
#define VCSMC_ERR_POOL (-3)     /* node pool exhausted (give the sweep a larger workspace) */
#define VCSMC_ERR_DATA (-4)     /* alignment entry is not a 0/1 state mask */
#define VCSMC_ERR_STATE (-5)    /* call order violated (e.g. backward before forward) */

int vcsmc_abi_version(void);
const char* vcsmc_last_error(void);
/* Number of kernels launched by this library in the calling process since load (bench.py's gpu_launches). */
uint64_t vcsmc_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * (a) alignment loader.  Replaces the dense one-hot [N,S,4] f64 genome of runner.py:107-115 and its
 *     K-fold host replication np.array([genome]*K) (vcsmc.py:479): states are packed ONCE into one
 *     byte per (taxon, site) holding a 4-bit mask (bit a set <=> genome[n,s,a] == 1; '-'/'?' = 0xF).
 *     status[0] is set to VCSMC_ERR_DATA if any entry is not exactly 0.0 or 1.0 or a site is all-zero.
 * --------------------------------------------------------------------------------------------- */
int vcsmc_pack_alignment(const double* genome, int n_taxa, int n_sites, uint8_t* codes, int* status, void* stream);
/* np.take(data, slice, axis=2) of vcsmc.py:533 on the packed codes: out[n, j] = codes[n, site_idx[j]]. */
int vcsmc_gather_sites(const uint8_t* codes, int n_taxa, int n_sites, const int32_t* site_idx, int n_sel,
                       uint8_t* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (b) transition matrices.  Replaces tf.linalg.expm(Q * b) on [K,4,4] (vcsmc.py:181-184).
 *     jc != 0: closed form for the reference's JC Q (vcsmc.py:126-129): P_ii = 1/4 + 3/4 e^-t,
 *     P_ij = 1/4 - 1/4 e^-t; Q is ignored.  jc == 0: scaling-and-squaring Taylor (degree 14) on t*Q per matrix.
 *     bwd: given dP (adjoint of P) returns dt[i] = <dP_i, Q P_i> and dQ_each[i] = t_i * L(t_i Q^T, dP_i)
 *     (Frechet adjoint); dQ_each may be NULL (JC).  This is the piece of TF autodiff
 *     (vcsmc.py:488-491) that differentiates expm.
 * --------------------------------------------------------------------------------------------- */
int vcsmc_transition_fwd(const double* Q, const double* t, int64_t n, int jc, double* P, void* stream);
/* The general-Q arithmetic of transition_fwd on the HOST (host pointers, no GPU): Q is the same for every matrix of a
 * call, so Q^k / k! (k <= 14) is tabulated once and P_i is a Horner scheme in the scalar t_i / 2^s on that table, then s
 * squarings -- what the kernels execute, operation for operation.  For tests of that scheme without a device. */
int vcsmc_transition_host(const double* Q, const double* t, int64_t n, double* P);
int vcsmc_transition_bwd(const double* Q, const double* t, const double* dP, int64_t n, int jc, double* dt,
                         double* dQ_each, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (c) merge.  Replaces broadcast_conditional_likelihood_K (vcsmc.py:180-188) fused with the new
 *     node's share of compute_forest_posterior (vcsmc.py:238-242):
 *        new[k,s,:] = (L_l[k,s,:] . P_l[k]) * (L_r[k,s,:] . P_r[k])         (row-vector convention)
 *        ell[k]     = sum_s log( sum_a pi[a] new[k,s,a] )
 *     Children are given by reference: lsrc/rsrc[k] < 0 means leaf -(src+1) (read from `codes`,
 *     row stride `codes_stride`), otherwise a slot of `pool` (slot stride `slot_sites` sites,
 *     each site 4 doubles).  dst[k] is the output slot (dst < 0: compute ell only, store nothing).
 *     P holds [K][32] doubles (P_l row-major then P_r).  ell_part must hold K * vcsmc_merge_tiles(n_sites)
 *     doubles; ell[k] is their fixed-order sum (deterministic).  Sites processed: [0, n_sites).
 *     bwd (reverse pruning, the per-site part of TF autodiff through the while-loop):
 *        given coef[k] = dELBO/d ell[k] and (optionally) the adjoint G_new of the new node,
 *        accumulates the children's adjoints into gpool (atomic, same slot numbering as pool),
 *        dP[k][32] += per-particle 4x4 adjoints, dpi[4] += adjoint of pi (summed over all particles).
 *     jc != 0 selects the JC-specialised kernels: P must be the JC closed form (only P[0], P[1] of each
 *     matrix are read) and dP is COMPRESSED: dP[k][0] += sum_i dP_l[i][i], dP[k][1] += sum_{i!=j} dP_l[i][j],
 *     dP[k][16], dP[k][17] likewise for the right child -- the layout vcsmc_transition_bwd(jc=1) consumes.
 * --------------------------------------------------------------------------------------------- */
int vcsmc_merge_tiles(int n_sites);
int vcsmc_merge_fwd(const uint8_t* codes, int64_t codes_stride, double* pool, int64_t slot_sites,
                    const int32_t* lsrc, const int32_t* rsrc, const int32_t* dst, const double* P,
                    const double* pi, int64_t K, int n_sites, int jc, double* ell_part, double* ell, void* stream);
int vcsmc_merge_bwd(const uint8_t* codes, int64_t codes_stride, const double* pool, double* gpool,
                    int64_t slot_sites, const int32_t* lsrc, const int32_t* rsrc, const int32_t* gsrc,
                    const double* P, const double* pi, const double* coef, int64_t K, int n_sites, int jc,
                    double* dP, double* dpi, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (d) proposal + resampling.
 *     propose_pairs replaces extend_partial_state (vcsmc.py:298-305): Gumbel top-2 on float32
 *     uniforms u[K,n]; because z = -log(-log u) is increasing in u the kernel ranks u itself
 *     (ties -> lower index, like tf.nn.top_k).  coal[K,2] = two largest (largest first),
 *     rem[K,n-2] = the others in ascending order.
 *     resample replaces resample (vcsmc.py:284-285): idx[j] = first i with cdf[i] > u[j]*total over
 *     the fp64 running sum of exp(logit - max); also returns logsumexp(lw) and the ESS.
 *     work must hold vcsmc_resample_work_doubles(K) doubles (CDF, statistics, per-tile partials of the multi-CTA scan).
 * --------------------------------------------------------------------------------------------- */
int vcsmc_propose_pairs(const float* u, int64_t K, int n, int32_t* coal, int32_t* rem, void* stream);
int64_t vcsmc_resample_work_doubles(int64_t K);
int vcsmc_resample(const double* lw, const double* u, int64_t K, int32_t* idx, double* lse, double* ess,
                   double* work, void* stream);

/* Counter-based uniforms (Philox4x32-10) keyed by (seed, rank event r, LOGICAL particle k, lane), so
 * 1/2/4/8-GPU runs consume identical numbers.  Fills what one rank event consumes:
 * u_pair[K,n] float32 in [0,1), u_bl/u_br[K] float64 in [tiny,1), u_res[K] float64 in [0,1). */
int vcsmc_philox_step_uniforms(uint64_t seed, int r, int64_t k0, int64_t K, int n, float* u_pair, double* u_bl,
                               double* u_br, double* u_res, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The sweep: sample_phylogenies + body_rank_update (vcsmc.py:332-451) forward, and the reverse
 * sweep that TF autodiff of `cost = -elbo` performs (vcsmc.py:488-491) backward.
 * --------------------------------------------------------------------------------------------- */
typedef struct vcsmc_sweep vcsmc_sweep_t;

typedef struct {
  int32_t n_taxa;        /* N */
  int32_t n_sites;       /* S (sites held by THIS rank) */
  int64_t n_particles;   /* K */
  int32_t jc;            /* 1 = JC closed form, 0 = general Q */
  int32_t keep_for_backward; /* 0 = forward only (nodes freed when dead) */
  int64_t workspace_bytes;   /* size of the workspace the caller will pass (0 = ask for the minimum) */
  int32_t n_sub;             /* 0 = VCSMC (vcsmc.py); M > 0 = VNCSMC nested look-ahead with M sub-samples (vncsmc.py, --M) */
  int32_t reserved;
} vcsmc_sweep_config;

typedef struct {
  int64_t min_bytes;         /* smallest workspace that can run this config */
  int64_t retain_bytes;      /* workspace at which every node is retained (no recompute in backward) */
} vcsmc_sweep_sizes;

/* Optional cross-rank hook (site sharding): sum `count` doubles in place across ranks, stream-ordered. */
typedef int (*vcsmc_allreduce_fn)(void* user, double* buf, int64_t count, void* stream);

/* Cross-rank hook of PARTICLE sharding (north_star: each GPU holds K/G particles).  The library asks the caller
 * (torch.distributed / NCCL) for the three collectives the protocol needs, all stream-ordered on `stream`:
 *   VCSMC_COMM_ALLGATHER  buf holds world chunks of `bytes` bytes, chunk g belongs to rank g; in place
 *   VCSMC_COMM_BARRIER    every rank's earlier work on `stream` is complete before any rank's later work starts
 *   VCSMC_COMM_ALLREDUCE  sum bytes/8 doubles in place
 * Everything else -- reading a remote ancestor's forest row, copying the nodes a rank lacks -- is done by the library's
 * own kernels through peer pointers (vcsmc_sweep_set_comm's peer_ws). */
#define VCSMC_COMM_ALLGATHER 1
#define VCSMC_COMM_BARRIER 2
#define VCSMC_COMM_ALLREDUCE 3
typedef int (*vcsmc_comm_fn)(void* user, int op, void* buf, int64_t bytes, void* stream);

/* Peer mapping of a device allocation (CUDA IPC), so that every rank can address every rank's workspace:
 * export: handle_host[64] and the byte offset of dev_ptr inside its allocation; open: maps a peer's allocation into
 * this process and returns its base (add the exporter's offset); close: unmaps. */
int vcsmc_ipc_export(void* dev_ptr, void* handle_host, int64_t* offset_host);
int vcsmc_ipc_open(const void* handle_host, void** base_out);
int vcsmc_ipc_close(void* base);

int vcsmc_sweep_query(const vcsmc_sweep_config* cfg, vcsmc_sweep_sizes* out);
int vcsmc_sweep_create(const vcsmc_sweep_config* cfg, void* workspace, vcsmc_sweep_t** out);
void vcsmc_sweep_destroy(vcsmc_sweep_t* h);
int vcsmc_sweep_set_allreduce(vcsmc_sweep_t* h, vcsmc_allreduce_fn fn, void* user);
/* Particle sharding over `world` <= 8 ranks of one NVLink domain: this rank owns logical particles
 * [rank K/world, (rank+1) K/world) (K must be divisible); peer_ws_host[g] is the address, in THIS process, of rank g's
 * workspace (its own for g == rank); all ranks create the sweep with the same config and workspace size.  Forward:
 * every rank scores its particles on all sites; per rank event one all-gather of the step record (weights, branch
 * lengths, child references), identical ancestors on every rank, owners materialise the survivors, ranks that drew a
 * remote ancestor pull the nodes they lack over NVLink.  Backward: sharded by SITE on the gathered tables
 * (options "site_begin"/"site_end"), gradients are summed by the caller.  VCSMC proposal only (n_sub == 0).
 * `fn` may be NULL: by default (option "peer_sync") the synchronisations and the record exchange of the forward run
 * over peer memory and nothing goes through the hook.  The caller must synchronise the ranks (any barrier) between
 * this call and the first forward. */
int vcsmc_sweep_set_comm(vcsmc_sweep_t* h, int rank, int world, vcsmc_comm_fn fn, void* user, void* const* peer_ws_host);
/* Options: "scalar_share" (default 1): fraction of the site-independent gradient terms this rank contributes
 * (site sharding: 1 on rank 0, 0 elsewhere, then sum the gradients across ranks);
 * "skip_zero" (default 1): backward skips rank events whose adjoint is zero (W underflowed to 0 and no descendant
 * uses the node; see skip_below for what counts as zero), set 0 to force the dense reverse sweep;
 * "skip_below" (default 2^-64): with skip_zero, an adjoint coefficient dELBO/d ell of magnitude <= skip_below * |grad_elbo|
 * counts as zero (softmax weights of 1e-200 are representable in fp64 but cannot change any digit of the gradient:
 * the event's coefficients sum to O(1)); 0 restores "exactly zero only".  This is an APPROXIMATION of the gradient:
 * the dropped terms are bounded by K * skip_below * max |d ell / d theta| per rank event; the two settings agree to
 * 1e-11 relative on primate.p (tested);
 * "max_chunk_sites" (default 0 = unlimited): cap on the site chunk of the recompute backward (testing aid);
 * "lazy" (default 1; VCSMC proposal only): the forward scores every particle without storing its node and materialises
 * only the particles that the next resampling draws as an ancestor; 0 = eager (every node stored as it is computed) --
 * results are identical;
 * "leaf_patterns" (default 1; lazy forward): merges of two leaves are scored from the site-pattern counts of the leaf
 * pair (tabulated once per sweep) instead of site by site -- same sum, different summation order;
 * "sparse_bwd" (default 1): when at most 1024 particle-events carry an adjoint (the rule with ESS ~ 1), the per-site part
 * of the reverse sweep is ONE site-parallel launch per chunk (a thread owns its site through all rank events) instead of
 * three launches per rank event; 0 keeps the per-event kernels -- same sums up to the order of the dP reduction;
 * "leaf_rows" (default 1; lazy forward, grouped order): merges of a leaf and an internal node are scored by the rows
 * kernel on the leaf's state-sorted sites (one row of the bilinear form per 128-position unit) -- same sum, different
 * summation order; 0 leaves them to the generic scoring kernel;
 * "force_sorted" (default 0): grouped visiting order even when K is too small for it to pay (testing aid);
 * "event_timing" (default 0; lazy forward): CTA 0 of the event kernel stamps %globaltimer at every phase boundary
 * (output "event_timing", uint64 [N][16]);
 * "peer_sync" (default 1; particle sharding): the two synchronisations of a rank event and the exchange of the step
 * record run over peer memory (flags polled inside the event kernel, peer loads) with no host involvement.  The
 * collective-hook route (0) of the first round is gone: the forward is one cooperative kernel per rank event and
 * cannot call back into the host between its phases;
 * "graph" (default 1; lazy forward): from the second forward on, the launch sequence of the forward sweep (no host
 * synchronisation, seed and model read from the workspace) is captured once into a CUDA graph and replayed;
 * "force_gc" (default 0): use the garbage-collected pool and the recompute backward even when every node fits (testing aid);
 * "site_begin", "site_end" (default 0, n_sites): the site slice this rank's reverse sweep covers;
 * "profile" (default 0): record CUDA events around every merge launch, read with vcsmc_sweep_profile. */
int vcsmc_sweep_set_option(vcsmc_sweep_t* h, const char* name, double value);

/* Uniform source: explicit arrays (u_pair is the ragged concatenation over r of [K, N-r] float32;
 * u_bl/u_br/u_res are [N-1,K] float64) or Philox from a seed. */
int vcsmc_sweep_set_uniforms(vcsmc_sweep_t* h, const float* u_pair, const double* u_bl, const double* u_br,
                             const double* u_res);
int vcsmc_sweep_set_seed(vcsmc_sweep_t* h, uint64_t seed);
/* VNCSMC explicit uniforms: look_bl / look_br are the ragged concatenation over r of [C(N-r,2)][M*K] float64 in
 * [tiny,1) (pair t in r1-major order, column m*K + k, as vncsmc.py:346-353 lays them out); cat and res are [N-1,K]
 * float64 in [0,1): the categorical draw over the C*M options (vncsmc.py:298) and the resampling draw. */
int vcsmc_sweep_set_uniforms_nested(vcsmc_sweep_t* h, const double* look_bl, const double* look_br, const double* cat,
                                    const double* res);

/* codes [N,S] (row stride = S); lam_l, lam_r [N-1] rates; Q [16]; pi [4]. */
int vcsmc_sweep_forward(vcsmc_sweep_t* h, const uint8_t* codes, const double* lam_l, const double* lam_r,
                        const double* Q, const double* pi, void* stream);
/* grads of grad_elbo * ELBO: dlam_l, dlam_r [N-1], dQ [16], dpi [4] (written, not accumulated). */
int vcsmc_sweep_backward(vcsmc_sweep_t* h, double grad_elbo, double* dlam_l, double* dlam_r, double* dQ,
                         double* dpi, void* stream);

/* Per-kernel device time of the merge launches since the option "profile" was set to 1 (CUDA events on the
 * launching stream): out_host[8] = {ms, launches} for the forward merge (eager) or scoring kernel (lazy), the recompute
 * merge of the chunked backward, the backward merge, and the lazy forward's cooperative event kernel (one launch per
 * rank event: weights, CDF, ancestors, rows, survivors, proposal) or, with n_sub > 0, the look-ahead kernel.
 * Synchronises on the recorded events and resets the counters. */
int vcsmc_sweep_profile(vcsmc_sweep_t* h, double* out_host);

/* Device pointers into the workspace, valid after forward:
 *   "elbo"[1] "log_weights"[N-1,K] "log_likelihood"[N-1,K] "log_likelihood_tilde"[K]
 *   "log_likelihood_R"[K] "left_branches"[N-1,K] "right_branches"[N-1,K] (float64)
 *   "v_minus"[K] "ancestors"[N-1,K] "left_ref"[N-1,K] "right_ref"[N-1,K] "leaf_counts"[N-1,K] (int32)
 *   "log_z"[N-1] "ess"[N-1] (float64)   "status"[8] (int32: error, peak pool slots, backward chunks, ...)
 *   "choice"[N-1,K] (int32, VNCSMC only: the chosen option t*M+m of vncsmc.py:298)
 *   "rem_positions" (uint8, ragged: rank event r holds [K, N-r-2] at byte offset sum_{r'<r} align16(K (N-r'-2)):
 *   the positions, in the ancestor's forest, of the subtrees a particle keeps, in the reference's order.  The lazy
 *   forward fills the rows of particles whose normalised weight is not zero in double precision -- the only ones that can
 *   be resampled or carry a gradient -- and leaves the others at position 0)
 *   "event_timing" (uint64 [N][16], lazy forward with option "event_timing")
 * Returns NULL for an unknown name. */
void* vcsmc_sweep_output(vcsmc_sweep_t* h, const char* name);

#ifdef __cplusplus
}
#endif
#endif /* VCSMC_B200_H */
