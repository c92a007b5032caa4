"""N>1 path on CPU: the PARTICLE-sharding protocol with world_size 2 over gloo.

The product has no CPU path, so the protocol is exercised with the oracle's primitives standing in for the kernels:
every rank owns K/G particles on all sites, the log-weights are all-gathered (phylo_b200.comm.Comm, the same object
the GPU path hands to vcsmc_sweep_set_comm), every rank derives the ancestors of ALL particles from the same uniforms,
and a rank that drew a remote ancestor takes that ancestor's forest from its owner.  ELBO, weights and ancestors must
equal the single-process sweep.  The reverse sweep of a particle-sharded run is sharded by site on the gathered tables:
that decomposition is what tests/test_sharding_gloo.py checks.
"""
import math
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vcsmc_oracle as O
from phylo_b200.sharding import choose_sharding, owner_of, particle_range, site_slice
from vcsmc_test_helpers import random_params, synthetic_genome

F64 = torch.float64


def test_partition_helpers():
    assert particle_range(64, 3, 4) == (48, 64)
    np.testing.assert_array_equal(owner_of(np.array([0, 15, 16, 63]), 64, 4), [0, 0, 1, 3])
    for S in (1, 7, 10000):
        for world in (1, 2, 8):
            cuts = [site_slice(S, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == S
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
    assert choose_sharding(None, 64, 1, False) == "none"
    assert choose_sharding(None, 64, 8, False) == "particles"
    assert choose_sharding(None, 60, 8, False) == "sites"          # K not divisible
    assert choose_sharding(None, 64, 8, True) == "sites"           # nested proposal
    assert choose_sharding("sites", 64, 8, False) == "sites"
    try:
        choose_sharding("particles", 60, 8, False)
        assert False
    except ValueError:
        pass


def _gather_rows(comm, local: torch.Tensor, K: int) -> torch.Tensor:
    """All ranks' rows of a per-particle tensor, through Comm.all_gather_inplace (the per-event record collective)."""
    kl = local.shape[0]
    flat = local.contiguous().view(torch.uint8).reshape(-1)
    buf = torch.empty(flat.numel() * comm.world, dtype=torch.uint8)
    buf[comm.rank * flat.numel():(comm.rank + 1) * flat.numel()] = flat
    comm.all_gather_inplace(buf, flat.numel())
    return buf.view(local.dtype).reshape((K,) + tuple(local.shape[1:]))


def _sharded_forward(g, K, lam_l, lam_r, Q, pi, U, comm):
    """One rank's share of the forward sweep (vcsmc.py:332-451) under particle sharding."""
    N, S, _ = g.shape
    k0, k1 = particle_range(K, comm.rank, comm.world)
    kl = k1 - k0
    pi = pi.reshape(-1)
    core = torch.from_numpy(np.array([g] * kl))
    record = torch.ones((kl, N), dtype=torch.int64)
    cum_l = torch.zeros(kl, dtype=F64)
    cum_r = torch.zeros(kl, dtype=F64)
    ll_tilde = torch.full((kl,), math.log(1.0 / K), dtype=F64)
    lw_all, ll_all, anc_all = [], [], []
    for r in range(N - 1):
        n = N - r
        if r > 0:
            idx = O.resample_indices(lw_all[-1].numpy(), U.res[r])          # ALL K ancestors, identical on every rank
            anc_all.append(idx)
            mine = torch.from_numpy(idx[k0:k1])
            core = _gather_rows(comm, core, K)[mine]                        # "migration": take the ancestors' forests
            record = _gather_rows(comm, record, K)[mine]
            ll_tilde = ll_all[-1][mine]
        coal_np, rem_np = O.propose_pairs(U.pair[r][k0:k1])
        coal, rem = torch.from_numpy(coal_np.astype(np.int64)), torch.from_numpy(rem_np.astype(np.int64))
        b_l = -torch.log(torch.from_numpy(U.bl[r][k0:k1])) / lam_l[r]
        b_r = -torch.log(torch.from_numpy(U.br[r][k0:k1])) / lam_r[r]
        cum_l, cum_r = cum_l + b_l, cum_r + b_r                              # slot-wise histories (quirk Q1)
        new = O.merge(O.gather_across(core, coal[:, 0:1]).squeeze(1), O.gather_across(core, coal[:, 1:2]).squeeze(1), b_l, b_r, Q)
        core = torch.cat([O.gather_across(core, rem), new.unsqueeze(1)], dim=1)
        record = torch.cat([O.gather_across(record, rem), O.gather_across(record, coal).sum(dim=1, keepdim=True)], dim=1)
        ll_r = O.compute_forest_posterior(core, record, pi, None)
        ll_r = ll_r + (-lam_l[r] * cum_l + (r + 1) * torch.log(lam_l[r])) + (-lam_r[r] * cum_r + (r + 1) * torch.log(lam_r[r]))
        v_minus = O.overcounting_correct(record)
        lw_r = ll_r - ll_tilde - (torch.log(lam_l[r]) - lam_l[r] * b_l + torch.log(lam_r[r]) - lam_r[r] * b_r) \
            + torch.log(v_minus.to(F64)) - 1.0 / O.ncr(n, 2)
        rec = _gather_rows(comm, torch.stack([lw_r, ll_r], dim=1), K)        # the step record: one all-gather per event
        lw_all.append(rec[:, 0].clone())
        ll_all.append(rec[:, 1].clone())
    log_weights = torch.stack(lw_all)
    elbo = O.compute_log_ZSMC(torch.cat([torch.zeros((1, K), dtype=F64), log_weights]), K)
    return float(elbo), log_weights.numpy(), np.stack(anc_all) if anc_all else np.zeros((0, K), dtype=np.int64)


def _worker(rank, world, port, jc, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from phylo_b200.comm import Comm
    comm = Comm()
    assert comm.min_int(10 + rank) == 10
    t = torch.full((3,), float(rank + 1), dtype=F64)
    comm.all_reduce(t)
    assert float(t[0]) == sum(range(1, world + 1))
    comm.barrier("cpu")
    g = synthetic_genome(7, 23, seed=4, gaps=0.05)
    N, K = 7, 24
    p = random_params(N, jc, seed=3)
    U = O.Uniforms.draw(N, K, seed=8)
    lam_l, lam_r, Q, pi = [x.detach() for x in O.model_from_params(p)]
    out[rank] = _sharded_forward(g, K, lam_l, lam_r, Q, pi, U, comm)
    dist.destroy_process_group()


def _run(jc, port, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, jc, out), nprocs=world, join=True)
    g = synthetic_genome(7, 23, seed=4, gaps=0.05)
    p = random_params(7, jc, seed=3)
    lam_l, lam_r, Q, pi = [x.detach() for x in O.model_from_params(p)]
    res = O.sweep(g, 24, lam_l, lam_r, Q, pi, O.Uniforms.draw(7, 24, seed=8))
    for rank in range(world):
        elbo, lw, anc = out[rank]
        assert abs(elbo - float(res.elbo)) <= 1e-11 * abs(float(res.elbo))
        np.testing.assert_allclose(lw, res.log_weights.numpy(), rtol=1e-10)   # summation order of the branch priors differs
        np.testing.assert_array_equal(anc, res.ancestors[1:])
    assert out[0][0] == out[1][0]


def test_particle_sharding_world2_jc():
    _run(True, 29621)


def test_particle_sharding_world3_gtr():
    _run(False, 29622, world=3)
