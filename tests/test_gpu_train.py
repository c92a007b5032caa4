"""The drop-in class on the GPU: VCSMC(datadict, K, args).train(...) protocol, outputs and one optimiser step vs the oracle."""
import argparse
import math
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import vcsmc_oracle as O

pytestmark = pytest.mark.gpu


def make_args(**kw):
    d = dict(dataset="primate_data", n_particles=32, batch_size=256, learning_rate=0.001, num_epoch=2,
             optimizer="GradientDescentOptimizer", branch_prior=math.log(10.0), M=10, nested=False, jcmodel=False,
             memory_optimization="on")
    d.update(kw)
    return argparse.Namespace(**d)


def oracle_uniforms(ops, seed, N, K):
    pair, bl, br, rs = [], [], [], []
    for r in range(N - 1):
        a, b, c, d = ops.philox_step_uniforms(seed, r, 0, K, N - r)
        pair.append(a.cpu().numpy()); bl.append(b.cpu().numpy()); br.append(c.cpu().numpy()); rs.append(d.cpu().numpy())
    return O.Uniforms(pair, np.stack(bl), np.stack(br), np.stack(rs))


@pytest.mark.parametrize("jc", [True, False])
def test_one_sgd_step_matches_oracle(primate_genome, jc):
    """cost = -ELBO on a site minibatch, one GradientDescent step (vcsmc.py:488-491,:533-534): variables after the
    step equal variable - lr * oracle gradient, for the reference's initial parameters."""
    from phylo_b200 import ops
    from phylo_b200.vcsmc import VCSMC
    g = primate_genome[:9]
    args = make_args(jcmodel=jc, n_particles=48)
    m = VCSMC({"taxa": ["t%d" % i for i in range(9)], "genome": g}, 48, args, seed=11)
    sites = np.random.default_rng(0).permutation(g.shape[1])[:256].astype(np.int32)
    opt = torch.optim.SGD(m.trainable_variables(), lr=0.01)
    cost = -m.sample_phylogenies(sites, need_grad=True, seed=777)
    cost.backward()
    opt.step()
    U = oracle_uniforms(ops, 777, 9, 48)
    p = O.Params.init(9, jc)
    res, grads = O.elbo_and_grads(g, 48, p, U, site_idx=sites)
    assert float(-cost) == pytest.approx(float(res.elbo), rel=1e-9)
    for v, p0, gr in zip(m.trainable_variables(), p.tensors(), grads):
        expect = p0 + 0.01 * gr                      # minimising -ELBO
        np.testing.assert_allclose(v.detach().cpu().numpy(), expect.numpy(), rtol=1e-9, atol=1e-9)


def test_train_protocol_and_results_file(primate_genome, tmp_path, monkeypatch):
    from phylo_b200.vcsmc import VCSMC
    monkeypatch.chdir(tmp_path)
    args = make_args(jcmodel=True, n_particles=32, optimizer="Adam")
    taxa = ["S%d" % i for i in range(12)]
    m = VCSMC({"taxa": taxa, "genome": primate_genome}, 32, args, seed=5)
    res = m.train(epochs=2, batch_size=256, learning_rate=0.01, verbose=False)
    keys = {"cost", "nParticles", "nTaxa", "lr", "log_weights", "Qmatrices", "left_branches", "right_branches", "log_lik",
            "ll_tilde", "log_lik_R", "jump_chain_evolution", "best_epoch", "best_log_lik", "best_jump_chain"}
    assert set(res.keys()) == keys                   # vcsmc.py:622-636
    assert res["cost"].shape == (2,) and np.isfinite(res["cost"]).all() and -8000 < res["cost"][0] < -6000
    assert res["log_weights"].shape == (2, 11, 32) and res["log_lik_R"].shape == (2, 32)
    assert res["Qmatrices"].shape == (2, 4, 4) and res["nTaxa"] == 12
    # 898 sites / 256 -> 3 full slices + 1 remainder; the LAST slice is never trained on (quirk Q8): 3 steps/epoch
    jc = res["best_jump_chain"]
    assert jc.shape == (32, 1 + sum(range(2, 13)))
    assert jc[0, 0] == "" and sorted(jc[0, 1:13]) == sorted(taxa)
    assert all(sorted(t.split("+")) == sorted(taxa) for t in m.final_trees)   # every particle's final tree has all taxa
    with open(os.path.join(m.save_dir, "results.p"), "rb") as f:
        saved = pickle.load(f)
    assert set(saved.keys()) == keys
    assert "Initial evaluation of ELBO" in open(os.path.join(m.save_dir, "run_parameters.txt")).read()
    # variables moved (3 Adam steps per epoch)
    assert not torch.allclose(m.left_branches_var, torch.full_like(m.left_branches_var, math.log(10.0)))


def test_runner_cli_flags():
    from phylo_b200.runner import parse_args
    a = parse_args(["--dataset=primate_data", "--n_particles=16", "--batch_size=1", "--jcmodel=true", "--twisting=true"])
    assert a.n_particles == 16 and a.batch_size == 1 and a.jcmodel is True and a.nested is True
    d = parse_args([])                                # the reference's defaults, runner.py:15-54
    assert (d.dataset, d.n_particles, d.batch_size, d.learning_rate, d.num_epoch, d.optimizer, d.M, d.nested, d.jcmodel) == \
        ("primate_data", 10, 256, 0.001, 100, "GradientDescentOptimizer", 10, False, False)
    assert d.branch_prior == pytest.approx(math.log(10))
