"""Pins the CPU oracle (oracle/vcsmc_oracle.py) before anything is compared against it.

Golden vectors come from the reference's own csmc.py executed live in the build container
(tests/golden/make_golden.py); the rest are analytic identities listed in SURVEY.md section 8c.
"""
import math
import os

import numpy as np
import pytest
import scipy.linalg as spl
import torch

from oracle import vcsmc_oracle as O


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "csmc_merge.npz"))


def test_merge_matches_csmc_live(gold):
    """oracle.merge == reference csmc.py:300-309 on dense, one-hot and gap children."""
    Q = torch.from_numpy(gold["Q"])
    out = O.merge(torch.from_numpy(gold["L_l"]), torch.from_numpy(gold["L_r"]),
                  torch.from_numpy(gold["b_l"]), torch.from_numpy(gold["b_r"]), Q).numpy()
    np.testing.assert_allclose(out, gold["merged"], rtol=1e-12, atol=0)


def test_pruning_matches_csmc_live(gold):
    """oracle.pruning_loglik == reference csmc.py:318-326 on a fixed 6-taxon tree."""
    merges = [(int(l), int(r), float(bl), float(br)) for l, r, bl, br in gold["merges"]]
    ll = O.pruning_loglik(gold["genome"], merges, torch.from_numpy(gold["Q"]), torch.from_numpy(gold["prior"]))
    assert ll == pytest.approx(float(gold["tree_loglik"]), rel=1e-12)


def test_ncr_matches_csmc_live(gold):
    np.testing.assert_array_equal(np.array([O.ncr(n, 2) for n in range(2, 13)]), gold["ncr_vals"])


def test_loader_matches_runner_lines(golden_dir):
    z = np.load(os.path.join(golden_dir, "loader.npz"))
    dd = O.form_dataset_from_strings([str(s) for s in z["toy_strings"]])
    np.testing.assert_array_equal(dd["genome"], z["toy_genome"])
    assert dd["taxa"] == [str(t) for t in z["toy_taxa"]]
    dd = O.form_dataset_from_strings([str(s) for s in z["primate_strings"]])
    np.testing.assert_array_equal(dd["genome"], z["primate_genome"].astype(np.float64))
    assert dd["genome"].shape == (12, 898, 4)


def test_jc_closed_form_vs_expm():
    """JC Q of vcsmc.py:126-129: P(t) = 1/4 + 3/4 e^-t on the diagonal, 1/4 - 1/4 e^-t off it."""
    Q = O.jc_Q()
    for t in (1e-6, 0.013, 0.1, 1.0, 7.3):
        P = spl.expm(Q.numpy() * t)
        d, o = 0.25 + 0.75 * math.exp(-t), 0.25 - 0.25 * math.exp(-t)
        ref = np.full((4, 4), o) + np.eye(4) * (d - o)
        np.testing.assert_allclose(P, ref, rtol=0, atol=5e-16)
        np.testing.assert_allclose(O.transition_matrices(Q, torch.tensor([t]))[0].numpy(), P, atol=1e-15)


def test_gtr_Q_rows_and_P_rows():
    y = torch.from_numpy(np.random.default_rng(0).normal(size=(4, 4)))
    Q = O.get_Q(y)
    np.testing.assert_allclose(Q.sum(dim=1).numpy(), 0, atol=1e-15)
    np.testing.assert_allclose(torch.diag(Q).numpy(), -1, atol=1e-15)
    P = O.transition_matrices(Q, torch.tensor([0.01, 0.3, 4.0]))
    np.testing.assert_allclose(P.sum(dim=2).numpy(), 1.0, atol=1e-14)
    np.testing.assert_allclose(P[1].numpy(), spl.expm(Q.numpy() * 0.3), atol=1e-15)


def test_log_double_factorial():
    n = torch.tensor([1, 2, 3, 5, 7, 8, 125])
    ref = [sum(math.log(j) for j in range(int(m), 1, -2)) for m in n]
    np.testing.assert_allclose(O.log_double_factorial(n).numpy(), ref, rtol=1e-14)


def test_propose_pairs_order_and_ties():
    u = np.array([[0.2, 0.9, 0.5, 0.7, 0.1]], dtype=np.float32)
    coal, rem = O.propose_pairs(u)
    assert coal.tolist() == [[1, 3]]
    assert rem.tolist() == [[4, 0, 2]]
    # exact ties: tf.nn.top_k puts the lower index first, in both calls (reference tie quirk)
    u = np.array([[0.5, 0.3, 0.3]], dtype=np.float32)
    coal, rem = O.propose_pairs(u)
    assert coal.tolist() == [[0, 1]] and rem.tolist() == [[1]]
    # ranking by u itself reproduces the float32 Gumbel ranking (what the CUDA kernel does)
    rng = np.random.default_rng(5)
    u = rng.random((512, 17), dtype=np.float32)
    coal, rem = O.propose_pairs(u)
    np.testing.assert_array_equal(coal, np.argsort(-u, axis=1, kind="stable")[:, :2])
    np.testing.assert_array_equal(rem, np.argsort(u, axis=1, kind="stable")[:, :15])


def test_resample_indices_basic():
    lw = np.log(np.array([0.1, 0.2, 0.3, 0.4]))
    idx = O.resample_indices(lw, np.array([0.0, 0.0999, 0.1001, 0.3001, 0.6001, 0.999999]))
    assert idx.tolist() == [0, 0, 1, 2, 3, 3]
    # -inf weight never chosen, shift invariance
    lw2 = np.array([-np.inf, 5.0, 5.0])
    assert O.resample_indices(lw2, np.array([0.0, 0.49, 0.51])).tolist() == [1, 1, 2]


def test_k1_sweep_equals_pruning(primate_genome):
    """K=1: the final log_likelihood_R is a plain Felsenstein likelihood of the sampled tree."""
    g = primate_genome[:7, :200]
    N = g.shape[0]
    p = O.Params.init(N, jcmodel=False)
    p.y_q = torch.from_numpy(np.random.default_rng(1).normal(size=(4, 4)) * 0.3)
    p.y_station = torch.from_numpy(np.random.default_rng(2).normal(size=4) * 0.3)
    # lam_l == lam_r so that quirk Q4 (right prior uses log(left param), vcsmc.py:262) is invisible
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    U = O.Uniforms.draw(N, 1, seed=3)
    res = O.sweep(g, 1, lam_l, lam_r, Q, pi, U)
    merges = []
    forest = list(range(N))
    for r in range(N - 1):
        c, rem = res.coal[r][0], res.rem[r][0]
        merges.append((forest[c[0]], forest[c[1]], float(res.left_branches[r, 0]), float(res.right_branches[r, 0])))
        forest = [forest[i] for i in rem] + [N + r]
    ll = O.pruning_loglik(g, merges, Q, pi)
    assert float(res.log_likelihood_R[0]) == pytest.approx(ll, rel=1e-12)


def test_elbo_row0_and_magnitude(primate_genome):
    """ELBO on primate.p, JC, untrained, K=16 lands where the README figure's curves start."""
    N = primate_genome.shape[0]
    p = O.Params.init(N, jcmodel=True)
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    res = O.sweep(primate_genome, 16, lam_l, lam_r, Q, pi, O.Uniforms.draw(N, 16, seed=1))
    assert -7600 < float(res.elbo) < -6600
    manual = sum(float(torch.logsumexp(res.log_weights[r] - math.log(16), 0)) for r in range(N - 1))
    assert float(res.elbo) == pytest.approx(manual, rel=1e-14)
    assert res.ancestors.shape == (N - 1, 16)


def test_oracle_grads_finite_difference(primate_genome):
    """autograd through the restatement agrees with central differences (ints held fixed by the same uniforms)."""
    g = primate_genome[:5, :60]
    N, K = 5, 8
    p = O.Params.init(N, jcmodel=False)
    rng = np.random.default_rng(7)
    p.y_q = torch.from_numpy(rng.normal(size=(4, 4)) * 0.2)
    p.y_station = torch.from_numpy(rng.normal(size=4) * 0.2)
    p.left_branches_param = torch.from_numpy(math.log(10) + rng.normal(size=N - 1) * 0.1)
    U = O.Uniforms.draw(N, K, seed=11)
    res, grads = O.elbo_and_grads(g, K, p, U)

    def f(q):
        return float(O.sweep(g, K, *O.model_from_params(q), U).elbo)

    eps = 1e-6
    for name, gi, sel in (("left_branches_param", 0, (1,)), ("y_q", 2, (1, 2)), ("y_station", 3, (2,))):
        base = getattr(p, name)
        plus, minus = base.clone(), base.clone()
        plus[sel] += eps
        minus[sel] -= eps
        qp = O.Params(**{**p.__dict__, name: plus})
        qm = O.Params(**{**p.__dict__, name: minus})
        fd = (f(qp) - f(qm)) / (2 * eps)
        assert float(grads[gi][sel]) == pytest.approx(fd, rel=2e-5, abs=1e-6)


# ------------------------------------------------------------------------------------------------------------
# The oracle against the reference's OWN source: /root/reference/vcsmc.py and vncsmc.py were imported unmodified
# and executed under an eager TensorFlow stand-in (tests/golden/tf_shim.py) by tests/golden/make_golden.py; the
# outputs are tests/golden/ref_sweeps.npz.  These tests pin rows a1, a3-a13 of SURVEY section 8 (quirks Q1-Q5 are
# thereby confirmed by the reference's code and not by reading it).
# ------------------------------------------------------------------------------------------------------------
from vcsmc_test_helpers import RefCase, ref_case_names  # noqa: E402

RTOL = 1e-12


def _check_common(c, res, grads):
    assert float(res.elbo) == pytest.approx(float(c["elbo"]), rel=RTOL)
    np.testing.assert_array_equal(res.ancestors, c["ancestors"])
    np.testing.assert_array_equal(res.v_minus.numpy(), c["v_minus"])
    for name, got in (("log_weights", res.log_weights), ("log_likelihood", res.log_likelihood),
                      ("log_likelihood_tilde", res.log_likelihood_tilde), ("log_likelihood_R", res.log_likelihood_R),
                      ("left_branches", res.left_branches), ("right_branches", res.right_branches)):
        np.testing.assert_allclose(got.detach().numpy(), c[name], rtol=RTOL, atol=1e-12, err_msg=name)
    for g, ref in zip(grads, c.grads_elbo()):
        np.testing.assert_allclose(g.numpy(), ref, rtol=1e-9, atol=1e-10 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize("name", ref_case_names())
def test_oracle_sweep_matches_reference_source(name):
    """sample_phylogenies + body_rank_update (vcsmc.py:332-451) and the autodiff of cost (vcsmc.py:488-491)."""
    c = RefCase(name)
    p = c.params()
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    np.testing.assert_allclose(Q.numpy(), c["Qmatrix"], rtol=1e-14, atol=1e-16)          # vcsmc.py:138-148 / :126-129
    np.testing.assert_allclose(pi.numpy(), c["stationary_probs"], rtol=1e-14)            # vcsmc.py:133-136
    res, grads = O.elbo_and_grads(c.genome, c.K, p, c.uniforms())
    _check_common(c, res, grads)
    for r in range(c.N - 1):                                                             # vcsmc.py:304-305
        np.testing.assert_array_equal(res.coal[r], c.coal(r))
        np.testing.assert_array_equal(res.rem[r], c.rem(r))


@pytest.mark.parametrize("name", ref_case_names(nested=True))
def test_oracle_nested_sweep_matches_reference_source(name):
    """compute_potentials / nested extend_partial_state / body_rank_update (vncsmc.py:295-499)."""
    c = RefCase(name)
    res, grads = O.elbo_and_grads_nested(c.genome, c.K, c.M, c.params(), c.uniforms())
    _check_common(c, res, grads)
    np.testing.assert_array_equal(np.stack(res.choices), c["choices"])                   # vncsmc.py:298


def test_reference_jump_chains_are_particle0_labels():
    """Quirk Q6 in the reference's own output: vcsmc.py:306-307 index the flattened label tensor without the
    k*(N-r) offset, so every particle's kept and coalesced labels are read from particle 0's row (after
    resampling).  The stored strings are therefore not the particles' trees; the product rebuilds trees from the
    integer tables instead (phylo_b200/trees.py)."""
    for name in ("gtr_n4_k2", "gtr_n8_k64_pert"):
        c = RefCase(name)
        N, K = c.N, c.K
        jck = np.array([["S%d" % i for i in range(N)]] * K, dtype=object)
        cols = [np.full((K, 1), "", dtype=object)]
        for r in range(N - 1):
            if r > 0:
                jck = jck[c["ancestors"][r]]                      # vcsmc.py:288
            cols.append(jck.copy())                               # vcsmc.py:324 / :329
            row0 = jck[0]
            keep = row0[c.rem(r)]
            new = row0[c.coal(r)[:, 0]] + "+" + row0[c.coal(r)[:, 1]]
            jck = np.concatenate([keep, new[:, None]], axis=1)    # vcsmc.py:313
        np.testing.assert_array_equal(np.concatenate(cols, axis=1).astype(str), c["jump_chains"])
