"""N>1 path on CPU: the site-sharding protocol (phylo_b200/sharding.py + DESIGN.md section 6) with world_size 2 over gloo.

The product has no CPU path, so the protocol is exercised with the oracle standing in for the kernels: each rank
sweeps its slice of the sites, all-reduces the forest log-likelihood sums, and the rank-sum of the gradients (site-
independent terms owned by rank 0) must equal the single-process gradient; weights and ancestors must be identical
on every rank.  The same protocol runs on GPUs through Sweep.set_allreduce / scalar_share.
"""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vcsmc_oracle as O
from phylo_b200.sharding import local_sites, scalar_share
from vcsmc_test_helpers import random_params, synthetic_genome


def test_local_sites_partition():
    idx = np.arange(103, dtype=np.int32)[::-1].copy()
    for world in (1, 2, 3, 8):
        parts = [local_sites(idx, r, world) for r in range(world)]
        np.testing.assert_array_equal(np.concatenate(parts), idx)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert scalar_share(0, 4) == 1.0 and scalar_share(3, 4) == 0.0 and scalar_share(0, 1) == 1.0


def _worker(rank, world, port, jc, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g = synthetic_genome(7, 61, seed=4, gaps=0.05)
    N, K = 7, 24
    p = random_params(N, jc, seed=3)
    U = O.Uniforms.draw(N, K, seed=8)
    leaves = [t.detach().clone().requires_grad_(True) for t in p.tensors()]
    q = O.Params(leaves[0], leaves[1], None, None) if jc else O.Params(*leaves)
    lam_l, lam_r, Q, pi = O.model_from_params(q)
    sites = local_sites(np.arange(g.shape[1], dtype=np.int32), rank, world)
    res = O.sweep(g, K, lam_l, lam_r, Q, pi, U, site_idx=sites, allreduce=lambda t: dist.all_reduce(t),
                  scalar_share=scalar_share(rank, world))
    grads = list(torch.autograd.grad(res.elbo, leaves))
    for gr in grads:
        dist.all_reduce(gr)                       # the per-step gradient all-reduce of VCSMC._allreduce_grads
    out[rank] = (float(res.elbo), res.ancestors.copy(), res.log_weights.detach().numpy().copy(), [x.numpy().copy() for x in grads])
    dist.destroy_process_group()


def _run(jc, port):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, jc, out), nprocs=2, join=True)
    g = synthetic_genome(7, 61, seed=4, gaps=0.05)
    res, grads = O.elbo_and_grads(g, 24, random_params(7, jc, seed=3), O.Uniforms.draw(7, 24, seed=8))
    for rank in (0, 1):
        elbo, anc, lw, gr = out[rank]
        assert abs(elbo - float(res.elbo)) <= 1e-12 * abs(float(res.elbo))
        np.testing.assert_array_equal(anc, res.ancestors)
        np.testing.assert_allclose(lw, res.log_weights.detach().numpy(), rtol=1e-12)
        for a, b in zip(gr, grads):
            np.testing.assert_allclose(a, b.numpy(), rtol=1e-9, atol=1e-10 * float(b.abs().max()))
    assert out[0][0] == out[1][0]                 # bit-identical ELBO on both ranks


def test_site_sharding_world2_jc():
    _run(True, 29611)


def test_site_sharding_world2_gtr():
    _run(False, 29612)


def _seed_worker(rank, world, port, explicit, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    from phylo_b200.sharding import shared_seed
    seed = shared_seed((100 + rank) if explicit else None, dist)      # ranks start from different seeds (or none at all)
    rng = random.Random(seed)
    sites = list(range(50))
    out[rank] = (seed, [rng.sample(sites, 8) for _ in range(3)])      # what VCSMC.batch_slices draws from
    dist.destroy_process_group()


def test_seed_and_minibatches_are_shared_across_ranks():
    """runner.py's default is seed=None: without a shared seed every rank would draw its own ancestors and its own site
    minibatches.  Rank 0's seed, explicit or drawn, is the run's."""
    for explicit, port in ((False, 29641), (True, 29642)):
        mgr = mp.Manager()
        out = mgr.dict()
        mp.spawn(_seed_worker, args=(2, port, explicit, out), nprocs=2, join=True)
        assert out[0] == out[1]
        if explicit:
            assert out[0][0] == 100
