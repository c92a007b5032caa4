"""Regression: particle sharding with enough particles per rank that the scoring kernel runs on the GROUPED order
(K/G * tiles >= 9472).  The grouping tables are carved for K particles but used for K/G: a rank must clear the table it
actually uses (a stale count table corrupted the visiting order on multi-GPU runs with large K)."""
import numpy as np
import pytest

from test_gpu_particle_sharding import RTOL, _run

pytestmark = pytest.mark.gpu


def test_particle_sharding_grouped_order():
    """2 ranks x 16,384 particles against ONE rank x 32,768 on the same GPU code (seeded): same ELBO, same ancestors, same
    gradients.  (Against the single-GPU run rather than the oracle: with 32,768 flat weights a last-bit difference between
    numpy's and the kernels' CDF could legitimately move one draw across a boundary.)"""
    import torch
    from oracle import vcsmc_oracle as O
    from phylo_b200 import ops
    from vcsmc_test_helpers import random_params
    import test_gpu_particle_sharding as T
    g, K = T._case("flat_grouped")
    N, S = g.shape[0], g.shape[1]
    p = random_params(N, False, seed=5)
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    dev = lambda x: torch.as_tensor(x).cuda().contiguous()
    sw = ops.Sweep(N, S, K, False)
    sw.set_seed(1234)
    elbo = float(sw.forward(ops.pack_alignment(dev(g)), dev(lam_l), dev(lam_r), dev(Q), dev(pi.reshape(-1))).item())
    anc = sw.output("ancestors").cpu().numpy().copy()
    lw = sw.output("log_weights").cpu().numpy().copy()
    grads = torch.cat([t.reshape(-1) for t in sw.backward(1.0)]).cpu().numpy()
    assert max(len(np.unique(anc[r])) for r in range(1, N - 1)) > 100       # flat weights: many lineages alive
    del sw
    outs = _run(2, False, "flat_grouped", 29770, seeded=True)
    for o in outs:
        assert o["elbo"] == pytest.approx(elbo, rel=1e-11)
        np.testing.assert_array_equal(o["ancestors"][1:], anc[1:])
        np.testing.assert_allclose(o["log_weights"], lw, rtol=1e-10)
        np.testing.assert_allclose(o["grads"], grads, rtol=1e-6, atol=1e-9 * np.abs(grads).max())


def test_seeded_sweep_with_more_than_64_taxa():
    """N = 70: forest rows no longer fit two entries per lane (the proposal kernel's 8-entries-per-lane instantiation,
    two Philox blocks per lane, a 256-key bitonic sort); seeded mode against the oracle fed the same Philox uniforms."""
    import torch
    from oracle import vcsmc_oracle as O
    from phylo_b200 import ops
    from vcsmc_test_helpers import synthetic_genome
    N, S, K, seed = 70, 24, 12, 424242
    g = synthetic_genome(N, S, seed=9, gaps=0.05)
    p = O.Params.init(N, False)
    pair, bl, br, rs = [], [], [], []
    for r in range(N - 1):
        a, b, c, d = ops.philox_step_uniforms(seed, r, 0, K, N - r)
        pair.append(a.cpu().numpy()); bl.append(b.cpu().numpy()); br.append(c.cpu().numpy()); rs.append(d.cpu().numpy())
    U = O.Uniforms(pair, np.stack(bl), np.stack(br), np.stack(rs))
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    res = O.sweep(g, K, lam_l, lam_r, Q, pi, U)
    dev = lambda x: torch.as_tensor(x).cuda().contiguous()
    sw = ops.Sweep(N, S, K, False, keep_for_backward=False)
    sw.set_seed(seed)
    for _ in range(2):   # the second forward replays the captured graph
        elbo = sw.forward(ops.pack_alignment(dev(g)), dev(lam_l), dev(lam_r), dev(Q), dev(pi.reshape(-1)))
        assert float(elbo) == pytest.approx(float(res.elbo), rel=RTOL)
        np.testing.assert_array_equal(sw.output("ancestors").cpu().numpy()[1:], res.ancestors[1:])
        np.testing.assert_allclose(sw.output("log_weights").cpu().numpy(), res.log_weights.detach().numpy(), rtol=RTOL)
