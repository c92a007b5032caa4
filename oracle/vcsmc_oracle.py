"""CPU oracle for the VCSMC hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  Nothing under ``phylo_b200/`` imports it, and the product path
fails loudly when the CUDA library is missing instead of falling back to this.

What it is: a float64 restatement, in torch-CPU dense ops, of the reference's
per-rank-event arithmetic (``/root/reference/vcsmc.py``), keeping the reference's dense
structure on purpose (K-replicated ``[K,n,S,4]`` core, three gathers + concat per step,
full-forest posterior every step, a batched matrix exponential) so that it can also be
timed as the "restated reference" CPU baseline.  All randomness is INJECTED as explicit
uniform arrays, because the reference seeds nothing.

Parity status: **parity unpinned** against TensorFlow.  TF 1.15 / TFP 0.7 cannot be
installed in this image, and the reference ships no tests or golden vectors.  What pins
this oracle instead (tests/test_oracle.py, tests/golden/):
  * the merge formula against the reference's own ``csmc.py:300-309`` executed live
    (tests/golden/make_golden.py -> tests/golden/csmc_merge.npz);
  * a full Felsenstein log-likelihood on fixed trees against ``csmc.py:318-326``;
  * analytic identities (JC closed form, rows of P sum to 1, K=1 sweep equals a plain
    pruning likelihood, ELBO row 0 contributes 0).
Third-party arithmetic restated from its published algorithm (tensorflow==1.15.0,
tensorflow_probability==0.7.0, requirements.txt:2-3):
  * ``tf.random.categorical``  (vcsmc.py:285): sequential fp64 running sum of
    exp(logit - max), one uniform per draw, ``upper_bound(u * total)``;
  * ``tf.nn.top_k`` (vcsmc.py:304-305): descending, ties -> lower index first;
  * ``tfp.distributions.Exponential.sample`` (vcsmc.py:353-356): ``-log(U)/rate``;
  * ``tf.linalg.expm`` (vcsmc.py:183-184): Taylor + scaling-and-squaring, see ``expm`` below.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

F64 = torch.float64

# ----------------------------------------------------------------------------------------
# loader  (runner.py:83-115)
# ----------------------------------------------------------------------------------------
ALPHABET_DIR_BLANK = {  # runner.py:91-96
    "A": [1, 0, 0, 0], "C": [0, 1, 0, 0], "G": [0, 0, 1, 0], "T": [0, 0, 0, 1],
    "-": [1, 1, 1, 1], "?": [1, 1, 1, 1],
}
ALPHABET_DIR = {k: ALPHABET_DIR_BLANK[k] for k in "ACGT"}  # runner.py:83-86


def form_dataset_from_strings(genome_strings, alphabet_dir=None, alphabet_num=4):
    """runner.py:107-115: strings -> {'taxa': ['S0',..], 'genome': [N,S,4] f64}."""
    alphabet_dir = ALPHABET_DIR_BLANK if alphabet_dir is None else alphabet_dir
    g = np.zeros([len(genome_strings), len(genome_strings[0]), alphabet_num])
    for i in range(g.shape[0]):
        for j in range(g.shape[1]):
            g[i, j] = alphabet_dir[genome_strings[i][j]]
    return {"taxa": ["S" + str(i) for i in range(g.shape[0])], "genome": g}


# ----------------------------------------------------------------------------------------
# small helpers  (vcsmc.py:23-57, :133-148)
# ----------------------------------------------------------------------------------------
def ncr(n: int, r: int) -> float:
    """vcsmc.py:23-27 (int products, true division -> float)."""
    return math.prod(range(n - r + 1, n + 1)) / math.prod(range(1, r + 1))


def log_double_factorial(n: torch.Tensor) -> torch.Tensor:
    """vcsmc.py:30-57: sum_{j = n, n-2, ... >= 2} log j, elementwise, float64."""
    n = n.to(F64).clone()
    result = torch.zeros_like(n)
    while bool((n >= 2).any()):
        result = torch.where(n >= 2, result + torch.log(torch.clamp(n, min=1.0)), result)
        n = n - 2
    return result


def get_Q(y_q: torch.Tensor) -> torch.Tensor:
    """vcsmc.py:122,138-148: off-diagonal row softmax of the logits, diagonal = -rowsum."""
    A = y_q.shape[0]
    off = 1.0 - torch.eye(A, dtype=F64)
    e = torch.exp(y_q * off) * off          # set_diag(y,0) -> exp -> set_diag(.,0)
    q_entry = e / e.sum(dim=1, keepdim=True)
    return q_entry - torch.diag(q_entry.sum(dim=1))


def jc_Q(A: int = 4) -> torch.Tensor:
    """vcsmc.py:126-129: off-diagonal 1/A, diagonal -(A-1)/A."""
    return torch.full((A, A), 1.0 / A, dtype=F64) - torch.eye(A, dtype=F64)


def get_stationary_probs(y_station: torch.Tensor) -> torch.Tensor:
    """vcsmc.py:133-136: softmax, shape [1,A]."""
    e = torch.exp(y_station)
    return (e / e.sum()).unsqueeze(0)


# ----------------------------------------------------------------------------------------
# hot-path pieces
# ----------------------------------------------------------------------------------------
def expm(A: torch.Tensor) -> torch.Tensor:
    """Batched matrix exponential accurate to ~1e-16, differentiable by autograd.

    Stands in for ``tf.linalg.expm`` (vcsmc.py:183-184; Pade + scaling-and-squaring inside TF).
    ``torch.linalg.matrix_exp`` is NOT used: on these 4x4 rate matrices it is only good to
    ~2e-11 (measured against a long-double Taylor series and against scipy.linalg.expm, which
    agree to 4e-17), too loose for a 1e-9 parity bar.  Here: scale each matrix by 2^-s so
    that ||A||_1 <= 1/2, degree-18 Taylor by Horner (remainder < 1e-21), then s squarings.
    """
    shape = A.shape
    A = A.reshape(-1, shape[-2], shape[-1])
    n = A.shape[-1]
    norm = A.detach().abs().sum(dim=1).max(dim=1).values
    s = torch.clamp(torch.ceil(torch.log2(torch.clamp(norm, min=1e-300) / 0.5)), min=0).to(torch.int64)
    B = A * torch.pow(torch.tensor(0.5, dtype=A.dtype), s.to(A.dtype)).reshape(-1, 1, 1)
    eye = torch.eye(n, dtype=A.dtype).expand_as(B)
    X = eye.clone()
    for k in range(18, 0, -1):
        X = eye + torch.matmul(B, X) / k
    for j in range(1, int(s.max()) + 1 if s.numel() else 0):
        X = torch.where((s >= j).reshape(-1, 1, 1), torch.matmul(X, X), X)
    return X.reshape(shape)


def transition_matrices(Q: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """vcsmc.py:181-184: P[k] = expm(b[k] * Q)."""
    return expm(b.reshape(-1, 1, 1) * Q.unsqueeze(0))


def merge(L_l, L_r, b_l, b_r, Q):
    """vcsmc.py:180-188: new[k,s,:] = (L_l[k,s,:] @ P_l[k]) * (L_r[k,s,:] @ P_r[k])."""
    return torch.matmul(L_l, transition_matrices(Q, b_l)) * torch.matmul(L_r, transition_matrices(Q, b_r))


def propose_pairs(u_pair: np.ndarray):
    """vcsmc.py:298-305: Gumbel top-2 on float32 uniforms.

    z = -log(-log u) in float32; coal = top_k(z, 2); rem = top_k(-z, n-2).
    ``tf.nn.top_k`` orders descending with ties broken towards the LOWER index, which a
    stable argsort of the negated key reproduces.  Returns int32 ``coal [K,2]``, ``rem [K,n-2]``.
    """
    u = np.asarray(u_pair, dtype=np.float32)
    n = u.shape[1]
    with np.errstate(divide="ignore"):
        z = -np.log(-np.log(u))
    z = z.astype(np.float32)
    coal = np.argsort(-z, axis=1, kind="stable")[:, :2].astype(np.int32)
    rem = np.argsort(z, axis=1, kind="stable")[:, : n - 2].astype(np.int32)  # top_k(-z)
    return coal, rem


def resample_indices(log_weights: np.ndarray, u_res: np.ndarray) -> np.ndarray:
    """vcsmc.py:284-285 + tf.random.categorical's published CPU algorithm.

    logits = lw - logsumexp(lw); running fp64 sum of exp(logits - max) in index order;
    draw j picks the first i with cdf[i] > u[j] * total (upper_bound), clamped to K-1.
    """
    lw = np.asarray(log_weights, dtype=np.float64)
    m0 = lw.max()
    lse = m0 + np.log(np.exp(lw - m0).sum())
    logits = lw - lse
    w = np.exp(logits - logits.max())
    cdf = np.cumsum(w)                                # sequential double running sum
    t = np.asarray(u_res, dtype=np.float64) * cdf[-1]
    idx = np.searchsorted(cdf, t, side="right")       # upper_bound
    return np.minimum(idx, lw.shape[0] - 1).astype(np.int64)


def gather_across(a: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """vcsmc.py:60-97: out[k,m,...] = a[k, idx[k,m], ...]."""
    K = a.shape[0]
    return a[torch.arange(K).unsqueeze(1), idx]


class _AllReduceSum(torch.autograd.Function):
    """Sum across ranks; the adjoint of a sum of per-rank terms w.r.t. each term is the identity."""

    @staticmethod
    def forward(ctx, x, fn):
        y = x.detach().clone()
        fn(y)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def compute_forest_posterior(core, leafnode_num_record, pi, allreduce=None):
    """vcsmc.py:231-245: sum_x sum_s log(pi . core[k,x,s,:]) - sum_x log (2 max(n_x,2) - 3)!!.

    ``allreduce`` (site sharding): the site sum is completed across ranks before the weights.
    """
    forest_lik = torch.matmul(core, pi.reshape(-1, 1)).squeeze(-1)      # [K,X,S]
    forest_loglik = torch.log(forest_lik).sum(dim=(1, 2))
    if allreduce is not None:
        forest_loglik = _AllReduceSum.apply(forest_loglik, allreduce)
    forest_logprior = (-log_double_factorial(2 * torch.clamp(leafnode_num_record, min=2) - 3)).sum(dim=1)
    return forest_loglik + forest_logprior


def overcounting_correct(leafnode_num_record):
    """vcsmc.py:247-252."""
    return (leafnode_num_record - (leafnode_num_record == 1).to(leafnode_num_record.dtype)).sum(dim=1)


def compute_log_ZSMC(log_weights, K):
    """vcsmc.py:270-277 on the full [N,K] array (row 0 = zeros contributes exactly 0)."""
    return torch.logsumexp(log_weights - math.log(K), dim=1).sum()


# ----------------------------------------------------------------------------------------
# uniforms
# ----------------------------------------------------------------------------------------
@dataclass
class Uniforms:
    """All randomness of one sweep, injected.

    pair[r]  float32 [K, N-r]   pair proposal at rank event r           (vcsmc.py:303)
    bl, br   float64 [N-1, K]   in [tiny,1): branch lengths b=-log(U)/rate   (vcsmc.py:353-356)
    res      float64 [N-1, K]   in [0,1): row r drives the resampling at rank event r (row 0 unused)
    """
    pair: List[np.ndarray]
    bl: np.ndarray
    br: np.ndarray
    res: np.ndarray

    @staticmethod
    def draw(N: int, K: int, seed: int = 0) -> "Uniforms":
        rng = np.random.Generator(np.random.PCG64(seed))
        tiny = np.finfo(np.float64).tiny
        pair = [rng.random((K, N - r), dtype=np.float32) for r in range(N - 1)]
        bl = np.maximum(rng.random((N - 1, K)), tiny)
        br = np.maximum(rng.random((N - 1, K)), tiny)
        res = rng.random((N - 1, K))
        return Uniforms(pair, bl, br, res)


@dataclass
class SweepResult:
    elbo: torch.Tensor
    log_weights: torch.Tensor        # [N-1,K]  (rows 1.. of the reference's array)
    log_likelihood: torch.Tensor     # [N-1,K]
    log_likelihood_tilde: torch.Tensor
    log_likelihood_R: torch.Tensor   # [K]
    left_branches: torch.Tensor      # [N-1,K]
    right_branches: torch.Tensor
    v_minus: torch.Tensor
    ancestors: np.ndarray            # int64 [N-1,K], row 0 = identity
    coal: List[np.ndarray]           # per step int32 [K,2]  (positions in the pre-merge forest)
    rem: List[np.ndarray]            # per step int32 [K,n-2]
    new_nodes: Optional[List[torch.Tensor]] = None   # per step [K,S,4] when keep_nodes
    forests: List[np.ndarray] = field(default_factory=list)  # per step int64 [K,n-1] node ids
    leaf_counts: Optional[np.ndarray] = None


def sweep(genome: np.ndarray, K: int, lam_l: torch.Tensor, lam_r: torch.Tensor, Q: torch.Tensor,
          pi: torch.Tensor, U: Uniforms, keep_nodes: bool = False,
          site_idx: Optional[np.ndarray] = None, allreduce=None, scalar_share: float = 1.0) -> SweepResult:
    """One forward SMC sweep: vcsmc.py:406-451 driving body_rank_update vcsmc.py:332-400.

    ``lam_l``/``lam_r`` are the rates exp(variable) [N-1]; ``Q`` [4,4]; ``pi`` [1,4] or [4].
    ``site_idx`` selects a site minibatch exactly like np.take(data, slice, axis=2) (vcsmc.py:533).
    Node ids in ``forests``: leaf i -> i; the node created at event (r,k) -> N + r*K + k.
    ``allreduce`` / ``scalar_share`` restate the site-sharding protocol of the product:
    each rank holds a slice of the sites, the forest log-likelihood is summed across ranks, and the gradient of
    the site-INDEPENDENT terms (branch priors, proposal density) is scaled by ``scalar_share`` so that the sum of
    the per-rank gradients is the full gradient (share 1 on one rank, 0 on the others).
    """
    def shared(t):
        return t if scalar_share == 1.0 else scalar_share * t + (1.0 - scalar_share) * t.detach()

    g = np.asarray(genome, dtype=np.float64)
    if site_idx is not None:
        g = np.take(g, site_idx, axis=1)
    N, S, A = g.shape
    pi = pi.reshape(-1)
    core = torch.from_numpy(np.array([g] * K))                        # vcsmc.py:479  [K,N,S,A]
    record = torch.ones((K, N), dtype=torch.int64)                    # :415
    left_branches = torch.zeros((1, K), dtype=F64)                    # :417-418
    right_branches = torch.zeros((1, K), dtype=F64)
    log_weights = torch.zeros((1, K), dtype=F64)                      # :420-421
    log_likelihood = torch.zeros((1, K), dtype=F64)
    ll_tilde = torch.full((K,), math.log(1.0 / K), dtype=F64)         # :422
    ids = np.tile(np.arange(N, dtype=np.int64), (K, 1))
    ancestors = np.tile(np.arange(K, dtype=np.int64), (N - 1, 1))
    coal_hist, rem_hist, new_nodes, forests = [], [], [], []
    v_minus = torch.ones((K,), dtype=torch.int64)

    for r in range(N - 1):
        n = N - r
        # -- resample (vcsmc.py:340-344, :279-289, :318-325); unconditional for r > 0
        if r > 0:
            idx_np = resample_indices(log_weights[r].detach().numpy(), U.res[r])
            idx = torch.from_numpy(idx_np)
            core = core[idx]
            record = record[idx]
            ids = ids[idx_np]
            ll_tilde = log_likelihood[r][idx]
            ancestors[r] = idx_np
        # -- pair proposal (vcsmc.py:291-316)
        q = 1.0 / ncr(n, 2)
        coal_np, rem_np = propose_pairs(U.pair[r])
        coal, rem = torch.from_numpy(coal_np.astype(np.int64)), torch.from_numpy(rem_np.astype(np.int64))
        coal_hist.append(coal_np); rem_hist.append(rem_np)
        # -- branch lengths (vcsmc.py:351-358)
        b_l = -torch.log(torch.from_numpy(U.bl[r])) / lam_l[r]
        b_r = -torch.log(torch.from_numpy(U.br[r])) / lam_r[r]
        left_branches = torch.cat([left_branches, b_l.unsqueeze(0)], dim=0)
        right_branches = torch.cat([right_branches, b_r.unsqueeze(0)], dim=0)
        # -- merge + state update (vcsmc.py:361-373)
        remaining_core = gather_across(core, rem)
        L_l = gather_across(core, coal[:, 0:1]).squeeze(1)
        L_r = gather_across(core, coal[:, 1:2]).squeeze(1)
        new = merge(L_l, L_r, b_l, b_r, Q)
        core = torch.cat([remaining_core, new.unsqueeze(1)], dim=1)
        rem_rec = gather_across(record, rem)
        new_rec = gather_across(record, coal).sum(dim=1, keepdim=True)
        record = torch.cat([rem_rec, new_rec], dim=1)
        new_id = (N + r * K + np.arange(K, dtype=np.int64))[:, None]
        ids = np.concatenate([np.take_along_axis(ids, rem_np.astype(np.int64), axis=1), new_id], axis=1)
        forests.append(ids.copy())
        if keep_nodes:
            new_nodes.append(new.detach().clone())
        # -- weights (vcsmc.py:376-395)
        ll_r = compute_forest_posterior(core, record, pi, allreduce)
        lsel = left_branches[1:r + 2]
        rsel = right_branches[1:r + 2]
        ll_r = ll_r + shared((-lam_l[r] * lsel + torch.log(lam_l[r])).sum(dim=0)
                             + (-lam_r[r] * rsel + torch.log(lam_r[r])).sum(dim=0))
        v_minus = overcounting_correct(record)
        lw_r = ll_r - ll_tilde \
            - shared(torch.log(lam_l[r]) - lam_l[r] * b_l + torch.log(lam_r[r]) - lam_r[r] * b_r) \
            + torch.log(v_minus.to(F64)) - q
        log_weights = torch.cat([log_weights, lw_r.unsqueeze(0)], dim=0)
        log_likelihood = torch.cat([log_likelihood, ll_r.unsqueeze(0)], dim=0)

    elbo = compute_log_ZSMC(log_weights, K)                            # :445
    lb, rb = left_branches[1:], right_branches[1:]                     # :443-444
    # get_log_likelihood, vcsmc.py:254-268 (quirk Q4: the right multiplier uses log(left param))
    l_prior = (torch.log(lam_l).unsqueeze(0) - lb.t() * lam_l.unsqueeze(0)).sum(dim=1)
    r_prior = (torch.log(lam_l).unsqueeze(0) - rb.t() * lam_r.unsqueeze(0)).sum(dim=1)
    ll_R = log_likelihood[N - 1] + log_double_factorial(torch.tensor(2.0 * N - 3)) - l_prior - r_prior
    return SweepResult(elbo=elbo, log_weights=log_weights[1:], log_likelihood=log_likelihood[1:],
                       log_likelihood_tilde=ll_tilde, log_likelihood_R=ll_R, left_branches=lb,
                       right_branches=rb, v_minus=v_minus, ancestors=ancestors, coal=coal_hist,
                       rem=rem_hist, new_nodes=new_nodes if keep_nodes else None, forests=forests,
                       leaf_counts=record.numpy())


# ----------------------------------------------------------------------------------------
# parameters + gradients (vcsmc.py:119-131 variables; vcsmc.py:488-491 autodiff)
# ----------------------------------------------------------------------------------------
@dataclass
class Params:
    """The reference's four TF variables (vcsmc.py:119-124)."""
    left_branches_param: torch.Tensor    # [N-1] log-rates
    right_branches_param: torch.Tensor   # [N-1]
    y_q: Optional[torch.Tensor]          # [4,4] logits (None in JC mode)
    y_station: Optional[torch.Tensor]    # [4]   logits (None in JC mode)

    @staticmethod
    def init(N: int, jcmodel: bool, branch_prior: float = math.log(10.0), A: int = 4) -> "Params":
        lb = torch.full((N - 1,), branch_prior, dtype=F64)
        rb = torch.full((N - 1,), branch_prior, dtype=F64)
        if jcmodel:
            return Params(lb, rb, None, None)
        return Params(lb, rb, torch.full((A, A), 1.0 / A, dtype=F64), torch.full((A,), 1.0 / A, dtype=F64))

    def tensors(self):
        return [t for t in (self.left_branches_param, self.right_branches_param, self.y_q, self.y_station)
                if t is not None]


def model_from_params(p: Params, A: int = 4):
    """vcsmc.py:119-131: (lam_l, lam_r, Q, pi[1,A])."""
    lam_l, lam_r = torch.exp(p.left_branches_param), torch.exp(p.right_branches_param)
    if p.y_q is None:
        return lam_l, lam_r, jc_Q(A), torch.full((1, A), 1.0 / A, dtype=F64)
    return lam_l, lam_r, get_Q(p.y_q), get_stationary_probs(p.y_station)


def elbo_and_grads(genome, K, p: Params, U: Uniforms, site_idx=None):
    """ELBO and d(ELBO)/d(variables) by torch autograd on the restatement.

    Mirrors what ``optimizer.minimize(self.cost)`` differentiates (vcsmc.py:488-491, cost = -ELBO):
    resampling indices and pair choices are integer constants, branch lengths are
    reparameterised through the rates.
    """
    leaves = [t.detach().clone().requires_grad_(True) for t in p.tensors()]
    if p.y_q is None:
        q = Params(leaves[0], leaves[1], None, None)
    else:
        q = Params(*leaves)
    lam_l, lam_r, Q, pi = model_from_params(q)
    res = sweep(genome, K, lam_l, lam_r, Q, pi, U, site_idx=site_idx)
    grads = torch.autograd.grad(res.elbo, leaves)
    return res, grads


# ----------------------------------------------------------------------------------------
# plain Felsenstein pruning on a fixed tree (for the K=1 identity and csmc.py cross-check)
# ----------------------------------------------------------------------------------------
def pruning_loglik(genome: np.ndarray, merges, Q: torch.Tensor, pi: torch.Tensor) -> float:
    """csmc.py:292-326: post-order message passing, then sum_s log(pi . root[s]).

    ``merges`` is a list of (left_id, right_id, b_l, b_r); node ids < N are leaves, the j-th
    merge creates node N + j.
    """
    g = np.asarray(genome, dtype=np.float64)
    nodes: Dict[int, torch.Tensor] = {i: torch.from_numpy(g[i]) for i in range(g.shape[0])}
    nid = g.shape[0]
    for (l, r, bl, br) in merges:
        P_l = expm(Q * bl)
        P_r = expm(Q * br)
        nodes[nid] = (nodes[l] @ P_l) * (nodes[r] @ P_r)
        nid += 1
    return float(torch.log(nodes[nid - 1] @ pi.reshape(-1)).sum())


# ----------------------------------------------------------------------------------------
# VNCSMC: nested look-ahead proposal (vncsmc.py:295-499); everything else is shared with VCSMC
# ----------------------------------------------------------------------------------------
@dataclass
class UniformsNested:
    """Randomness of one VNCSMC sweep.

    look_bl[r], look_br[r]  float64 [C(N-r,2), M*K] in [tiny,1): branch samples of pair t (r1-major enumeration,
                            vncsmc.py:324-377), sub-particle m, particle k at column m*K + k  (vncsmc.py:350-353)
    cat   float64 [N-1, K]  in [0,1): the categorical draw over the C*M options of each particle (vncsmc.py:298)
    res   float64 [N-1, K]  in [0,1): resampling (row 0 unused)
    """
    look_bl: List[np.ndarray]
    look_br: List[np.ndarray]
    cat: np.ndarray
    res: np.ndarray

    @staticmethod
    def draw(N: int, K: int, M: int, seed: int = 0) -> "UniformsNested":
        rng = np.random.Generator(np.random.PCG64(seed))
        tiny = np.finfo(np.float64).tiny
        bl = [np.maximum(rng.random((int(ncr(N - r, 2)), M * K)), tiny) for r in range(N - 1)]
        br = [np.maximum(rng.random((int(ncr(N - r, 2)), M * K)), tiny) for r in range(N - 1)]
        return UniformsNested(bl, br, rng.random((N - 1, K)), rng.random((N - 1, K)))


def categorical_rows(logits: np.ndarray, u: np.ndarray) -> np.ndarray:
    """tf.random.categorical(logits [K,C], 1) (vncsmc.py:298), one injected uniform per row: per row the running
    fp64 sum of exp(logit - rowmax) in column order, then upper_bound(u * total), clamped."""
    lg = np.asarray(logits, dtype=np.float64)
    w = np.exp(lg - lg.max(axis=1, keepdims=True))
    cdf = np.cumsum(w, axis=1)
    t = np.asarray(u, dtype=np.float64) * cdf[:, -1]
    idx = (cdf <= t[:, None]).sum(axis=1)
    return np.minimum(idx, lg.shape[1] - 1).astype(np.int64)


def tree_posterior_K(data, leaf_counts, pi):
    """vncsmc.py:217-233 (MK variant included: one leaf-count column): sum_s log(pi . data[k,s,:]) - log (2 max(c,2)-3)!!."""
    lik = torch.matmul(data, pi.reshape(-1, 1)).squeeze(-1)
    return torch.log(lik).sum(dim=1) - log_double_factorial(2 * torch.clamp(leaf_counts, min=2) - 3)


def compute_potentials(core, record, lam_l_r, lam_r_r, Q, pi, M, u_bl, u_br):
    """vncsmc.py:324-416: log-softmaxed look-ahead potentials [K, C*M] (column t*M+m) and the branch samples."""
    K, n = core.shape[0], core.shape[1]
    pots, lbs, rbs, pairs = [], [], [], []
    t = 0
    for r1 in range(n - 1):
        for r2 in range(r1 + 1, n):
            L_l, L_r = core[:, r1], core[:, r2]
            L_l_MK, L_r_MK = L_l.repeat(M, 1, 1), L_r.repeat(M, 1, 1)              # index m*K + k  (:346-347)
            b_l = -torch.log(torch.from_numpy(u_bl[t])) / lam_l_r                    # (:350-353)
            b_r = -torch.log(torch.from_numpy(u_br[t])) / lam_r_r
            merged = merge(L_l_MK, L_r_MK, b_l, b_r, Q)
            c_l, c_r = record[:, r1].repeat(M), record[:, r2].repeat(M)
            joint = tree_posterior_K(merged, c_l + c_r, pi) - tree_posterior_K(L_l_MK, c_l, pi) \
                - tree_posterior_K(L_r_MK, c_r, pi)                                  # (:363-365)
            pots.append(joint); lbs.append(b_l); rbs.append(b_r); pairs.append((r1, r2))
            t += 1
    C = len(pairs)
    pot = torch.stack(pots).reshape(C * M, K).t()                                    # (:404-406)
    pot = pot - torch.logsumexp(pot, dim=1, keepdim=True)                            # (:407)
    l_br = torch.stack(lbs).reshape(C * M, K).t()
    r_br = torch.stack(rbs).reshape(C * M, K).t()
    return pot, np.array(pairs, dtype=np.int64), l_br, r_br


def sweep_nested(genome: np.ndarray, K: int, M: int, lam_l, lam_r, Q, pi, U: UniformsNested,
                 site_idx: Optional[np.ndarray] = None) -> SweepResult:
    """One forward VNCSMC sweep: vncsmc.py:511-555 driving body_rank_update vncsmc.py:432-499."""
    g = np.asarray(genome, dtype=np.float64)
    if site_idx is not None:
        g = np.take(g, site_idx, axis=1)
    N, S, A = g.shape
    pi = pi.reshape(-1)
    core = torch.from_numpy(np.array([g] * K))
    record = torch.ones((K, N), dtype=torch.int64)
    left_branches = torch.zeros((1, K), dtype=F64)
    right_branches = torch.zeros((1, K), dtype=F64)
    log_weights = torch.zeros((1, K), dtype=F64)
    log_likelihood = torch.zeros((1, K), dtype=F64)
    ll_tilde = torch.full((K,), math.log(1.0 / K), dtype=F64)
    ids = np.tile(np.arange(N, dtype=np.int64), (K, 1))
    ancestors = np.tile(np.arange(K, dtype=np.int64), (N - 1, 1))
    coal_hist, rem_hist, forests, choices = [], [], [], []
    v_minus = torch.ones((K,), dtype=torch.int64)
    ar = torch.arange(K)

    for r in range(N - 1):
        n = N - r
        if r > 0:                                                            # vncsmc.py:440-446
            idx_np = resample_indices(log_weights[r].detach().numpy(), U.res[r])
            idx = torch.from_numpy(idx_np)
            core, record, ids = core[idx], record[idx], ids[idx_np]
            ll_tilde = log_likelihood[r][idx]
            ancestors[r] = idx_np
        # twist the proposal (vncsmc.py:449) and extend (vncsmc.py:295-322)
        pot, pairs, l_all, r_all = compute_potentials(core, record, lam_l[r], lam_r[r], Q, pi, M, U.look_bl[r], U.look_br[r])
        choice = categorical_rows(pot.detach().numpy(), U.cat[r])            # :298
        choices.append(choice)
        pair_idx = choice // M                                               # :299
        coal_np = pairs[pair_idx].astype(np.int32)                           # :301  (r1 < r2)
        rem_np = np.stack([np.array([i for i in range(n - 1, -1, -1) if i not in (c[0], c[1])], dtype=np.int32)
                           for c in coal_np]).reshape(K, n - 2)              # :302-305 descending index order
        ch = torch.from_numpy(choice)
        q_log = pot[ar, ch]                                                  # :315-316 (a true log here)
        b_l, b_r = l_all[ar, ch], r_all[ar, ch]                              # :317-320
        coal, rem = torch.from_numpy(coal_np.astype(np.int64)), torch.from_numpy(rem_np.astype(np.int64))
        coal_hist.append(coal_np); rem_hist.append(rem_np)
        left_branches = torch.cat([left_branches, b_l.unsqueeze(0)], dim=0)
        right_branches = torch.cat([right_branches, b_r.unsqueeze(0)], dim=0)
        # the chosen merge is recomputed (vncsmc.py:458-465)
        remaining_core = gather_across(core, rem)
        L_l = gather_across(core, coal[:, 0:1]).squeeze(1)
        L_r = gather_across(core, coal[:, 1:2]).squeeze(1)
        new = merge(L_l, L_r, b_l, b_r, Q)
        core = torch.cat([remaining_core, new.unsqueeze(1)], dim=1)
        record = torch.cat([gather_across(record, rem), gather_across(record, coal).sum(dim=1, keepdim=True)], dim=1)
        new_id = (N + r * K + np.arange(K, dtype=np.int64))[:, None]
        ids = np.concatenate([np.take_along_axis(ids, rem_np.astype(np.int64), axis=1), new_id], axis=1)
        forests.append(ids.copy())
        # weights (vncsmc.py:472-491): identical to VCSMC except that q_log_proposal is a log-probability
        ll_r = compute_forest_posterior(core, record, pi)
        ll_r = ll_r + (-lam_l[r] * left_branches[1:r + 2] + torch.log(lam_l[r])).sum(dim=0) \
                    + (-lam_r[r] * right_branches[1:r + 2] + torch.log(lam_r[r])).sum(dim=0)
        v_minus = overcounting_correct(record)
        lw_r = ll_r - ll_tilde - (torch.log(lam_l[r]) - lam_l[r] * b_l + torch.log(lam_r[r]) - lam_r[r] * b_r) \
            + torch.log(v_minus.to(F64)) - q_log
        log_weights = torch.cat([log_weights, lw_r.unsqueeze(0)], dim=0)
        log_likelihood = torch.cat([log_likelihood, ll_r.unsqueeze(0)], dim=0)

    elbo = compute_log_ZSMC(log_weights, K)
    lb, rb = left_branches[1:], right_branches[1:]
    l_prior = (torch.log(lam_l).unsqueeze(0) - lb.t() * lam_l.unsqueeze(0)).sum(dim=1)
    r_prior = (torch.log(lam_l).unsqueeze(0) - rb.t() * lam_r.unsqueeze(0)).sum(dim=1)      # quirk Q4 as in vcsmc.py
    ll_R = log_likelihood[N - 1] + log_double_factorial(torch.tensor(2.0 * N - 3)) - l_prior - r_prior
    res = SweepResult(elbo=elbo, log_weights=log_weights[1:], log_likelihood=log_likelihood[1:],
                      log_likelihood_tilde=ll_tilde, log_likelihood_R=ll_R, left_branches=lb, right_branches=rb,
                      v_minus=v_minus, ancestors=ancestors, coal=coal_hist, rem=rem_hist, new_nodes=None,
                      forests=forests, leaf_counts=record.numpy())
    res.choices = choices
    return res


def elbo_and_grads_nested(genome, K, M, p: Params, U: UniformsNested, site_idx=None):
    """VNCSMC ELBO and d(ELBO)/d(variables): autograd flows through the chosen potential AND the log-softmax
    normaliser, i.e. through every look-ahead merge."""
    leaves = [t.detach().clone().requires_grad_(True) for t in p.tensors()]
    q = Params(leaves[0], leaves[1], None, None) if p.y_q is None else Params(*leaves)
    lam_l, lam_r, Q, pi = model_from_params(q)
    res = sweep_nested(genome, K, M, lam_l, lam_r, Q, pi, U, site_idx=site_idx)
    return res, torch.autograd.grad(res.elbo, leaves)
