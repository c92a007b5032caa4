"""Particle sharding on GPUs through the C ABI (vcsmc_sweep_set_comm): every rank owns K/G particles, the step record is
all-gathered per rank event, nodes of remote ancestors are pulled through peer pointers, the reverse sweep is sharded
by site.  Results must equal the single-process oracle: ancestors / child references bit-exact, weights and ELBO to
1e-9, summed gradients to 1e-7.

With fewer GPUs than ranks the ranks SHARE a device: the peer mapping (CUDA IPC) and every kernel of the protocol run
exactly as on separate GPUs, only the two collectives go through gloo instead of NCCL (NCCL refuses two ranks per
device).  On a multi-GPU box the same test runs over NCCL / NVLink.
"""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import vcsmc_oracle as O
from vcsmc_test_helpers import gpu_uniforms, random_params, refs_from_oracle, synthetic_genome

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _case(name):
    if name == "peaked":      # many sites: ESS ~ 1, one or two survivors per event, almost every ancestor is remote
        g = synthetic_genome(8, 300, seed=21, gaps=0.02)
        return g, 64
    if name == "flat":        # few sites: flat weights, hundreds of survivors, many nodes pulled per event
        g = synthetic_genome(9, 5, seed=3, gaps=0.1)
        return g, 128
    if name == "flat_large":  # thousands of survivors per event: long fetch lists, every rank pulls from every rank
        g = synthetic_genome(6, 7, seed=12, gaps=0.1)
        return g, 4096
    if name == "flat_grouped":  # K/G large enough for the grouped visiting order of the scoring kernel (regression)
        g = synthetic_genome(6, 7, seed=12, gaps=0.1)
        return g, 32768
    raise KeyError(name)


def _worker(rank, world, port, jc, name, seeded, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(rank % ndev)
    if ndev >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank % ndev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from phylo_b200 import ops
    from phylo_b200.comm import Comm
    from phylo_b200.sharding import scalar_share, site_slice

    g, K = _case(name)
    N, S = g.shape[0], g.shape[1]
    p = random_params(N, jc, seed=5)
    U = O.Uniforms.draw(N, K, seed=11)
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    dev = lambda x: torch.as_tensor(x).cuda().contiguous()
    comm = Comm()
    codes = ops.pack_alignment(dev(g))
    sw = ops.Sweep(N, S, K, jc, comm=comm)
    if seeded:
        sw.set_seed(1234)
    else:
        sw.set_uniforms(*gpu_uniforms(U))
    s0, s1 = site_slice(S, rank, world)
    sw.set_option("site_begin", float(s0))
    sw.set_option("site_end", float(s1))
    sw.set_option("scalar_share", scalar_share(rank, world))
    res = {}
    for it in range(2):   # twice: the second sweep reuses pool, tables and peer mappings
        elbo = sw.forward(codes, dev(lam_l), dev(lam_r), None if jc else dev(Q), dev(pi.reshape(-1)))
        grads = sw.backward(1.0)
        flat = torch.cat([t.reshape(-1) for t in grads if t is not None])
        comm.all_reduce(flat)
        res = {k: sw.output(k).cpu().numpy().copy() for k in
               ("log_weights", "log_likelihood", "log_likelihood_tilde", "log_likelihood_R", "left_branches",
                "right_branches", "v_minus", "ancestors", "left_ref", "right_ref", "leaf_counts")}
        res["elbo"] = float(elbo.item())
        res["grads"] = flat.cpu().numpy()
        res["info"] = sw.check_status()
    out[rank] = res
    del sw
    dist.destroy_process_group()


def _run(world, jc, name, port, seeded=False):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, jc, name, seeded, out), nprocs=world, join=True)
    return [out[r] for r in range(world)]


def _oracle(jc, name):
    g, K = _case(name)
    N = g.shape[0]
    p = random_params(N, jc, seed=5)
    U = O.Uniforms.draw(N, K, seed=11)
    lam_l, lam_r, Q, pi = [t.detach().clone().requires_grad_(True) for t in O.model_from_params(p)]
    res = O.sweep(g, K, lam_l, lam_r, Q, pi, U)
    gs = torch.autograd.grad(res.elbo, [lam_l, lam_r, Q, pi], allow_unused=True)
    names = ["dlam_l", "dlam_r", "dQ", "dpi"]
    flat = np.concatenate([np.asarray(x.numpy()).reshape(-1) for nme, x in zip(names, gs) if not (jc and nme == "dQ")])
    return res, flat, N, K


@pytest.mark.parametrize("name", ["peaked", "flat"])
@pytest.mark.parametrize("jc", [True, False])
@pytest.mark.parametrize("world", [2, 4])
def test_particle_sharding_matches_oracle(world, jc, name):
    res, g_ref, N, K = _oracle(jc, name)
    lref, rref = refs_from_oracle(res, N, K)
    outs = _run(world, jc, name, 29700 + world * 10 + int(jc) * 2 + (name == "flat"))
    for o in outs:
        np.testing.assert_array_equal(o["ancestors"][1:], res.ancestors[1:])
        np.testing.assert_array_equal(o["left_ref"], lref)
        np.testing.assert_array_equal(o["right_ref"], rref)
        np.testing.assert_allclose(o["left_branches"], res.left_branches.detach().numpy(), rtol=1e-14)
        np.testing.assert_allclose(o["log_weights"], res.log_weights.detach().numpy(), rtol=RTOL)
        np.testing.assert_allclose(o["log_likelihood"], res.log_likelihood.detach().numpy(), rtol=RTOL)
        np.testing.assert_allclose(o["log_likelihood_tilde"], res.log_likelihood_tilde.detach().numpy(), rtol=RTOL)
        np.testing.assert_allclose(o["log_likelihood_R"], res.log_likelihood_R.detach().numpy(), rtol=RTOL)
        np.testing.assert_array_equal(o["v_minus"], res.v_minus.numpy())
        assert o["elbo"] == pytest.approx(float(res.elbo), rel=RTOL)
        scale = np.abs(g_ref).max()
        np.testing.assert_allclose(o["grads"], g_ref, rtol=1e-7, atol=1e-9 * scale)
    # every rank holds the same gathered tables
    for o in outs[1:]:
        assert o["elbo"] == outs[0]["elbo"]
        np.testing.assert_array_equal(o["log_weights"], outs[0]["log_weights"])


def test_particle_sharding_many_pulls():
    res, g_ref, N, K = _oracle(False, "flat_large")
    assert max(len(np.unique(res.ancestors[r])) for r in range(1, N - 1)) > 50   # many distinct lineages alive
    outs = _run(4, False, "flat_large", 29760)
    for o in outs:
        np.testing.assert_array_equal(o["ancestors"][1:], res.ancestors[1:])
        np.testing.assert_allclose(o["log_weights"], res.log_weights.detach().numpy(), rtol=RTOL)
        assert o["elbo"] == pytest.approx(float(res.elbo), rel=RTOL)
        np.testing.assert_allclose(o["grads"], g_ref, rtol=1e-7, atol=1e-9 * np.abs(g_ref).max())
        assert o["info"]["peak_pool_slots"] > 20


def test_particle_sharding_seeded_equals_single_gpu():
    """Philox uniforms are keyed by the LOGICAL particle index: 1 and 2 ranks draw the same numbers and agree."""
    from phylo_b200 import ops
    g, K = _case("peaked")
    N, S = g.shape[0], g.shape[1]
    p = random_params(N, False, seed=5)
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    dev = lambda x: torch.as_tensor(x).cuda().contiguous()
    sw = ops.Sweep(N, S, K, False)
    sw.set_seed(1234)
    elbo = float(sw.forward(ops.pack_alignment(dev(g)), dev(lam_l), dev(lam_r), dev(Q), dev(pi.reshape(-1))).item())
    anc = sw.output("ancestors").cpu().numpy().copy()
    grads = torch.cat([t.reshape(-1) for t in sw.backward(1.0)]).cpu().numpy()
    outs = _run(2, False, "peaked", 29790, seeded=True)
    for o in outs:
        assert o["elbo"] == pytest.approx(elbo, rel=1e-12)
        np.testing.assert_array_equal(o["ancestors"][1:], anc[1:])
        np.testing.assert_allclose(o["grads"], grads, rtol=1e-7, atol=1e-9 * np.abs(grads).max())


def _train_worker(rank, world, port, sharding, out, nested=False):
    """Two optimiser steps through the drop-in class under torch.distributed (the path runner.py takes under torchrun)."""
    import argparse
    import math
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(rank % ndev)
    if world > 1:
        if ndev >= world:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank % ndev))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
    from phylo_b200.vcsmc import VCSMC
    g = synthetic_genome(8, 300, seed=21, gaps=0.02)
    args = argparse.Namespace(dataset="synthetic", n_particles=64, batch_size=128, learning_rate=0.01, num_epoch=1,
                              optimizer="GradientDescentOptimizer", branch_prior=math.log(10.0), M=3, nested=nested,
                              jcmodel=False, memory_optimization="on")
    m = VCSMC({"taxa": ["t%d" % i for i in range(8)], "genome": g}, 64, args, seed=3, sharding=sharding)
    assert m.sharding == (sharding if world > 1 else "none")
    opt = torch.optim.SGD(m.trainable_variables(), lr=0.01)
    sites = np.random.default_rng(1).permutation(300)[:128].astype(np.int32)
    elbos = []
    for it in range(2):
        opt.zero_grad(set_to_none=True)
        cost = -m.sample_phylogenies(sites, need_grad=True, seed=500 + it)
        cost.backward()
        m._allreduce_grads()
        opt.step()
        elbos.append(float(-cost))
    full = float(m.sample_phylogenies(need_grad=False, seed=900))
    out[rank] = (elbos, full, [v.detach().cpu().numpy().copy() for v in m.trainable_variables()])
    m.release()
    if world > 1:
        dist.destroy_process_group()


def test_nested_proposal_site_sharded_agrees_with_one_rank():
    """VNCSMC (look-ahead proposal, vncsmc.py:295-499) across GPUs runs site-sharded: two ranks against one rank --
    same ELBOs, same updated variables (the single-rank path is checked against the oracle and the reference goldens)."""
    mgr = mp.Manager()
    ref, two = mgr.dict(), mgr.dict()
    mp.spawn(_train_worker, args=(1, 29830, "sites", ref, True), nprocs=1, join=True)
    mp.spawn(_train_worker, args=(2, 29831, "sites", two, True), nprocs=2, join=True)
    e1, f1, v1 = ref[0]
    for rank in (0, 1):
        e2, f2, v2 = two[rank]
        np.testing.assert_allclose(e2, e1, rtol=1e-10)
        assert f2 == pytest.approx(f1, rel=1e-10)
        for a, b in zip(v2, v1):
            np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("sharding", ["particles", "sites"])
def test_training_steps_agree_across_world_sizes(sharding):
    """VCSMC(...).sample_phylogenies / backward / SGD under 1 and 2 ranks: same ELBOs, same updated variables."""
    mgr = mp.Manager()
    ref, two = mgr.dict(), mgr.dict()
    mp.spawn(_train_worker, args=(1, 29800, sharding, ref), nprocs=1, join=True)
    mp.spawn(_train_worker, args=(2, 29801 + (sharding == "sites"), sharding, two), nprocs=2, join=True)
    e1, f1, v1 = ref[0]
    for rank in (0, 1):
        e2, f2, v2 = two[rank]
        np.testing.assert_allclose(e2, e1, rtol=1e-10)
        assert f2 == pytest.approx(f1, rel=1e-10)
        for a, b in zip(v2, v1):
            np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10)


def _unseeded_train_worker(rank, world, port, sharding, out):
    """runner.py's default under torchrun: seed=None.  Every rank must end up with rank 0's seed and slices."""
    import argparse
    import math
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(rank % ndev)
    if ndev >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank % ndev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from phylo_b200.vcsmc import VCSMC
    g = synthetic_genome(8, 300, seed=21, gaps=0.02)
    args = argparse.Namespace(dataset="synthetic", n_particles=64, batch_size=64, learning_rate=0.01, num_epoch=1,
                              optimizer="GradientDescentOptimizer", branch_prior=math.log(10.0), M=10, nested=False,
                              jcmodel=False, memory_optimization="on")
    m = VCSMC({"taxa": ["t%d" % i for i in range(8)], "genome": g}, 64, args, seed=None, sharding=sharding)
    slices = m.batch_slices(300, 64)
    res = m.train(epochs=2, batch_size=64, learning_rate=0.01, save=False, verbose=False)
    out[rank] = (m.seed, slices, res["cost"], [v.detach().cpu().numpy().copy() for v in m.trainable_variables()])
    m._sweeps.clear()
    m._last = None
    dist.destroy_process_group()


@pytest.mark.parametrize("sharding", ["particles", "sites"])
def test_unseeded_training_is_consistent_across_ranks(sharding):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_unseeded_train_worker, args=(2, 29820 + (sharding == "sites"), sharding, out), nprocs=2, join=True)
    s0, sl0, c0, v0 = out[0]
    s1, sl1, c1, v1 = out[1]
    assert s0 == s1 and sl0 == sl1
    assert np.isfinite(c0).all() and -4000 < c0[0] < -1000
    np.testing.assert_allclose(c1, c0, rtol=1e-10)
    for a, b in zip(v1, v0):
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10)
