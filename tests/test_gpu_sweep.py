"""GPU parity of the whole sweep (forward + reverse sweep) against the CPU oracle, through the C ABI."""
import math

import numpy as np
import pytest
import torch

from oracle import vcsmc_oracle as O
from vcsmc_test_helpers import gpu_uniforms, random_params, refs_from_oracle, synthetic_genome

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from phylo_b200 import ops as _ops
    return _ops


def dev(x):
    return torch.as_tensor(x).cuda().contiguous()


def run_gpu(ops, genome, K, p, U, jc, workspace_bytes=None, grads=True, skip_zero=True, max_chunk_sites=0, lazy=True,
            force_gc=False, skip_below=None):
    N, S = genome.shape[0], genome.shape[1]
    codes = ops.pack_alignment(dev(genome))
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    sw = ops.Sweep(N, S, K, jc, keep_for_backward=grads, workspace_bytes=workspace_bytes)
    sw.set_uniforms(*gpu_uniforms(U))
    sw.set_option("skip_zero", 1.0 if skip_zero else 0.0)
    sw.set_option("max_chunk_sites", float(max_chunk_sites))
    sw.set_option("lazy", 1.0 if lazy else 0.0)
    if force_gc:
        sw.set_option("force_gc", 1.0)
    if skip_below is not None:
        sw.set_option("skip_below", float(skip_below))
    elbo = sw.forward(codes, dev(lam_l), dev(lam_r), None if jc else dev(Q), dev(pi.reshape(-1)))
    out = {k: sw.output(k).cpu().numpy().copy() for k in
           ("log_weights", "log_likelihood", "log_likelihood_tilde", "log_likelihood_R", "left_branches",
            "right_branches", "v_minus", "ancestors", "left_ref", "right_ref", "leaf_counts")}
    out["elbo"] = float(elbo.item())
    g = None
    if grads:
        g = [None if t is None else t.cpu().numpy() for t in sw.backward(1.0)]
    out["info"] = sw.check_status()
    return out, g, sw


def oracle_param_grads(genome, K, p, U):
    """Oracle gradients w.r.t. (lam_l, lam_r, Q, pi) -- the quantities the C ABI differentiates."""
    lam_l, lam_r, Q, pi = [t.detach().clone().requires_grad_(True) for t in O.model_from_params(p)]
    res = O.sweep(genome, K, lam_l, lam_r, Q, pi, U)
    gs = torch.autograd.grad(res.elbo, [lam_l, lam_r, Q, pi], allow_unused=True)
    return res, [None if g is None else g.numpy() for g in gs]


def compare_forward(out, res, N, K):
    anc = res.ancestors.copy()
    np.testing.assert_array_equal(out["ancestors"][1:], anc[1:])                       # bit-exact ancestors
    lref, rref = refs_from_oracle(res, N, K)
    np.testing.assert_array_equal(out["left_ref"], lref)                                # bit-exact pair choices
    np.testing.assert_array_equal(out["right_ref"], rref)
    np.testing.assert_allclose(out["left_branches"], res.left_branches.detach().numpy(), rtol=1e-14)
    np.testing.assert_allclose(out["log_weights"], res.log_weights.detach().numpy(), rtol=RTOL)
    np.testing.assert_allclose(out["log_likelihood"], res.log_likelihood.detach().numpy(), rtol=RTOL)
    np.testing.assert_allclose(out["log_likelihood_tilde"], res.log_likelihood_tilde.detach().numpy(), rtol=RTOL)
    np.testing.assert_allclose(out["log_likelihood_R"], res.log_likelihood_R.detach().numpy(), rtol=RTOL)
    np.testing.assert_array_equal(out["v_minus"], res.v_minus.numpy())
    assert out["elbo"] == pytest.approx(float(res.elbo), rel=RTOL)


def compare_grads(g, g_ref, jc):
    names = ["dlam_l", "dlam_r", "dQ", "dpi"]
    for name, a, b in zip(names, g, g_ref):
        if jc and name in ("dQ",):
            continue
        b = np.asarray(b).reshape(np.asarray(a).shape)
        scale = np.abs(b).max() + 1e-300
        np.testing.assert_allclose(a, b, rtol=1e-7, atol=1e-9 * scale, err_msg=name)


@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("jc", [True, False])
@pytest.mark.parametrize("K", [1, 16, 64])
def test_sweep_primate_subset(ops, primate_genome, jc, K, lazy):
    g = primate_genome[:8, :300]
    N = g.shape[0]
    p = random_params(N, jc, seed=K)
    U = O.Uniforms.draw(N, K, seed=100 + K)
    res, g_ref = oracle_param_grads(g, K, p, U)
    out, grads, sw = run_gpu(ops, g, K, p, U, jc, lazy=lazy)
    assert sw.retained
    compare_forward(out, res, N, K)
    compare_grads(grads, g_ref, jc)


@pytest.mark.parametrize("jc", [True, False])
def test_sweep_primate_full(ops, primate_genome, jc):
    """primate.p, all 12 taxa x 898 sites (BASELINE config 1/2 shape at an oracle-sized K)."""
    g = primate_genome
    N, K = g.shape[0], 32
    p = O.Params.init(N, jc)           # the reference's initial parameters (vcsmc.py:119-131)
    U = O.Uniforms.draw(N, K, seed=7)
    res, g_ref = oracle_param_grads(g, K, p, U)
    out, grads, _ = run_gpu(ops, g, K, p, U, jc)
    compare_forward(out, res, N, K)
    compare_grads(grads, g_ref, jc)
    assert -7600 < out["elbo"] < -6500   # where the README figure's curves start
    # adjoint coefficients below 2^-64 are skipped by default; "exact zeros only" visits more events, same gradients
    out0, grads0, _ = run_gpu(ops, g, K, p, U, jc, skip_below=0.0)
    assert out0["info"]["backward_events_visited"] >= out["info"]["backward_events_visited"]
    for a, b in zip(grads, grads0):
        if a is not None:
            np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-13 * np.abs(b).max())
    compare_grads(grads0, g_ref, jc)


@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("jc", [True, False])
def test_sweep_flat_weights_many_lineages(ops, jc, lazy):
    """Short alignment => flat weights => many distinct ancestors and shared nodes with several consumers."""
    g = synthetic_genome(9, 5, seed=3, gaps=0.1)
    N, K = 9, 128
    p = random_params(N, jc, seed=5)
    U = O.Uniforms.draw(N, K, seed=11)
    res, g_ref = oracle_param_grads(g, K, p, U)
    assert len(np.unique(res.ancestors[4])) > 10
    out, grads, _ = run_gpu(ops, g, K, p, U, jc, lazy=lazy)
    compare_forward(out, res, N, K)
    compare_grads(grads, g_ref, jc)
    # dense reverse sweep (no zero-adjoint skipping) gives the same gradients
    _, grads_dense, _ = run_gpu(ops, g, K, p, U, jc, skip_zero=False, lazy=lazy)
    compare_grads(grads_dense, g_ref, jc)


@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("jc", [True, False])
def test_sweep_gc_pool_and_chunked_backward(ops, primate_genome, jc, lazy):
    """Smallest workspace: garbage-collected forward pool + site-chunked recompute backward == retained mode."""
    g = primate_genome[:10]
    N, K = g.shape[0], 48
    p = random_params(N, jc, seed=2)
    U = O.Uniforms.draw(N, K, seed=3)
    res, g_ref = oracle_param_grads(g, K, p, U)
    probe = ops.Sweep(N, g.shape[1], K, jc, workspace_bytes=None)
    small = probe.min_bytes
    assert small < probe.retain_bytes
    del probe
    out, grads, sw = run_gpu(ops, g, K, p, U, jc, workspace_bytes=small, lazy=lazy)
    # eager: every event allocates K slots; lazy: only survivors are materialised
    assert not sw.retained and out["info"]["backward_chunks"] >= 1 and out["info"]["peak_pool_slots"] >= (1 if lazy else K)
    if lazy:
        assert out["info"]["peak_pool_slots"] < K
    compare_forward(out, res, N, K)
    compare_grads(grads, g_ref, jc)
    # several site chunks (ragged last chunk: 898 = 3 x 256 + 130), dense reverse sweep
    out, grads, sw = run_gpu(ops, g, K, p, U, jc, workspace_bytes=small, max_chunk_sites=256, skip_zero=False, lazy=lazy)
    assert out["info"]["backward_chunks"] == 4
    compare_grads(grads, g_ref, jc)


@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("jc", [True, False])
def test_sweep_flat_weights_gc_pool_sorted_order(ops, jc, lazy):
    """Many particles on few sites: the visiting order is sorted by child pair, thousands of distinct survivors go
    through the slot allocator of the garbage-collected pool, and the backward recomputes by chunk."""
    g = synthetic_genome(6, 7, seed=12, gaps=0.1)
    N, K = 6, 12288
    p = random_params(N, jc, seed=6)
    U = O.Uniforms.draw(N, K, seed=13)
    res, g_ref = oracle_param_grads(g, K, p, U)
    assert len(np.unique(res.ancestors[3])) > 1000
    probe = ops.Sweep(N, g.shape[1], K, jc)
    roomy = probe.retain_bytes + (32 << 20)   # the slot tables of the garbage-collected pool need a little extra room
    del probe
    for force_gc in (False, True):
        out, grads, sw = run_gpu(ops, g, K, p, U, jc, lazy=lazy, force_gc=force_gc, workspace_bytes=roomy)
        assert (out["info"]["peak_pool_slots"] > 1000) == force_gc
        compare_forward(out, res, N, K)
        compare_grads(grads, g_ref, jc)


@pytest.mark.parametrize("jc", [True, False])
@pytest.mark.parametrize("force_gc", [False, True])
def test_repeated_forward_replays_a_cuda_graph(ops, primate_genome, jc, force_gc):
    """From the second forward on the launch sequence is a captured CUDA graph: new seeds / parameters must take effect
    (they live in the workspace), and results must equal those of a fresh sweep object."""
    g = primate_genome[:9, :500]
    N, S, K = 9, 500, 96
    codes = ops.pack_alignment(dev(g))
    probe = ops.Sweep(N, S, K, jc)
    roomy = probe.retain_bytes + (32 << 20)
    del probe

    def fresh(seed, p):
        lam_l, lam_r, Q, pi = O.model_from_params(p)
        sw = ops.Sweep(N, S, K, jc, workspace_bytes=roomy)
        if force_gc:
            sw.set_option("force_gc", 1.0)
        sw.set_option("graph", 0.0)
        sw.set_seed(seed)
        e = float(sw.forward(codes, dev(lam_l), dev(lam_r), None if jc else dev(Q), dev(pi.reshape(-1))).item())
        return e, sw.output("ancestors").cpu().numpy().copy(), [None if t is None else t.cpu().numpy() for t in sw.backward(1.0)]

    sw = ops.Sweep(N, S, K, jc, workspace_bytes=roomy)
    if force_gc:
        sw.set_option("force_gc", 1.0)
    for it, (seed, pseed) in enumerate([(11, 1), (12, 1), (13, 2), (11, 1)]):
        p = random_params(N, jc, seed=pseed)
        lam_l, lam_r, Q, pi = O.model_from_params(p)
        sw.set_seed(seed)
        e = float(sw.forward(codes, dev(lam_l), dev(lam_r), None if jc else dev(Q), dev(pi.reshape(-1))).item())
        anc = sw.output("ancestors").cpu().numpy().copy()
        grads = [None if t is None else t.cpu().numpy() for t in sw.backward(1.0)]
        e0, anc0, grads0 = fresh(seed, p)
        assert e == pytest.approx(e0, rel=1e-13), it
        np.testing.assert_array_equal(anc, anc0)
        for a, b in zip(grads, grads0):
            if a is not None:
                np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12 * np.abs(b).max())


@pytest.mark.parametrize("jc", [True, False])
def test_lazy_equals_eager_at_scale(ops, primate_genome, jc):
    """K = 8192 on primate.p (too large for the CPU oracle): the lazy schedule -- grouped scoring with 32-particle groups
    and several site chunks, leaf pairs from site patterns, leaf rows, survivors only, graph replay, thresholded reverse
    sweep -- against the eager schedule with the dense reverse sweep, same seed."""
    g = primate_genome
    N, S, K = g.shape[0], g.shape[1], 8192
    p = random_params(N, jc, seed=4)
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    codes = ops.pack_alignment(dev(g))
    args = (codes, dev(lam_l), dev(lam_r), None if jc else dev(Q), dev(pi.reshape(-1)))

    def run(lazy, reps):
        sw = ops.Sweep(N, S, K, jc)
        sw.set_option("lazy", 1.0 if lazy else 0.0)
        sw.set_option("skip_zero", 1.0 if lazy else 0.0)
        sw.set_seed(77)
        for _ in range(reps):
            elbo = float(sw.forward(*args).item())
            grads = [None if t is None else t.cpu().numpy() for t in sw.backward(1.0)]
        sw.check_status()
        return elbo, sw.output("ancestors").cpu().numpy().copy(), sw.output("log_weights").cpu().numpy().copy(), grads

    e0, a0, w0, g0 = run(False, 1)
    e1, a1, w1, g1 = run(True, 3)      # the third forward replays the captured graph
    assert e1 == pytest.approx(e0, rel=1e-12)
    np.testing.assert_array_equal(a1[1:], a0[1:])
    np.testing.assert_allclose(w1, w0, rtol=1e-10)
    for a, b in zip(g1, g0):
        if a is not None:
            np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10 * np.abs(b).max())


def test_sweep_pool_exhaustion_is_reported(ops):
    """Flat weights keep many nodes alive; a 2K-slot pool must fail loudly, not silently corrupt."""
    from phylo_b200._lib import VcsmcError
    g = synthetic_genome(16, 3, seed=1)
    N, K = 16, 256
    p = O.Params.init(N, True)
    U = O.Uniforms.draw(N, K, seed=1)
    probe = ops.Sweep(N, 3, K, True, keep_for_backward=False)
    small = probe.min_bytes
    del probe
    codes = ops.pack_alignment(dev(g))
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    sw = ops.Sweep(N, 3, K, True, keep_for_backward=False, workspace_bytes=small)
    sw.set_uniforms(*gpu_uniforms(U))
    sw.forward(codes, dev(lam_l), dev(lam_r), None, dev(pi.reshape(-1)))
    try:
        info = sw.check_status()
        assert info["peak_pool_slots"] <= 2 * K + 64   # it fitted: fine, and says how much it used
    except VcsmcError as e:
        assert e.code == -3


def test_single_site_batch(ops, primate_genome):
    """batch_size=1 (BASELINE config 1): the reference mis-shapes here (tf.squeeze, quirk Q9); intended semantics."""
    g = primate_genome[:6, 17:18]
    N, K = 6, 16
    p = O.Params.init(N, True)
    U = O.Uniforms.draw(N, K, seed=4)
    res, g_ref = oracle_param_grads(g, K, p, U)
    out, grads, _ = run_gpu(ops, g, K, p, U, True)
    compare_forward(out, res, N, K)
    compare_grads(grads, g_ref, True)


@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("N,S,K", [(2, 40, 7), (3, 1, 5), (5, 513, 33), (4, 1025, 1)])
def test_sweep_odd_shapes(ops, N, S, K, lazy):
    """Smallest trees (one rank event), particle counts that are not a multiple of a warp, site counts around tile edges."""
    g = synthetic_genome(N, S, seed=N + S, gaps=0.1)
    for jc in (True, False):
        p = random_params(N, jc, seed=K)
        U = O.Uniforms.draw(N, K, seed=S)
        res, g_ref = oracle_param_grads(g, K, p, U)
        out, grads, _ = run_gpu(ops, g, K, p, U, jc, lazy=lazy)
        compare_forward(out, res, N, K)
        compare_grads(grads, g_ref, jc)
        if lazy:   # and again on the replayed graph
            out, grads, _ = run_gpu(ops, g, K, p, U, jc, lazy=lazy, force_gc=True, workspace_bytes=64 << 20)
            compare_forward(out, res, N, K)
            compare_grads(grads, g_ref, jc)


def test_autograd_through_parameterisation(ops, primate_genome):
    """sweep_elbo under torch autograd reproduces the oracle's gradients w.r.t. the reference's four variables."""
    g = primate_genome[:7, :200]
    N, K = 7, 32
    p = random_params(N, False, seed=9)
    U = O.Uniforms.draw(N, K, seed=10)
    res, g_ref = O.elbo_and_grads(g, K, p, U)
    codes = ops.pack_alignment(dev(g))
    sw = ops.Sweep(N, g.shape[1], K, False)
    sw.set_uniforms(*gpu_uniforms(U))
    leaves = [dev(t).requires_grad_(True) for t in p.tensors()]
    lam_l, lam_r = torch.exp(leaves[0]), torch.exp(leaves[1])
    off = 1.0 - torch.eye(4, dtype=torch.float64, device="cuda")
    e = torch.exp(leaves[2] * off) * off
    qe = e / e.sum(dim=1, keepdim=True)
    Q = qe - torch.diag(qe.sum(dim=1))
    pi = torch.softmax(leaves[3], dim=0)
    elbo = ops.sweep_elbo(sw, codes, lam_l, lam_r, Q, pi)
    assert float(elbo) == pytest.approx(float(res.elbo), rel=RTOL)
    (-elbo).backward()
    for a, b in zip(leaves, g_ref):
        scale = float(b.abs().max())
        np.testing.assert_allclose(-a.grad.cpu().numpy(), b.numpy(), rtol=1e-7, atol=1e-9 * scale)


def test_seeded_sweep_matches_oracle_fed_the_same_philox_uniforms(ops, primate_genome):
    """Philox mode: dump the uniforms the sweep consumes, feed them to the oracle, compare."""
    g = primate_genome[:9, :400]
    N, K, seed = 9, 64, 20261018
    p = O.Params.init(N, False)
    pair, bl, br, rs = [], [], [], []
    for r in range(N - 1):
        a, b, c, d = ops.philox_step_uniforms(seed, r, 0, K, N - r)
        pair.append(a.cpu().numpy()); bl.append(b.cpu().numpy()); br.append(c.cpu().numpy()); rs.append(d.cpu().numpy())
    U = O.Uniforms(pair, np.stack(bl), np.stack(br), np.stack(rs))
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    res = O.sweep(g, K, lam_l, lam_r, Q, pi, U)
    codes = ops.pack_alignment(dev(g))
    sw = ops.Sweep(N, g.shape[1], K, False, keep_for_backward=False)
    sw.set_seed(seed)
    elbo = sw.forward(codes, dev(lam_l), dev(lam_r), dev(Q), dev(pi.reshape(-1)))
    assert float(elbo) == pytest.approx(float(res.elbo), rel=RTOL)
    np.testing.assert_array_equal(sw.output("ancestors").cpu().numpy()[1:], res.ancestors[1:])


@pytest.mark.parametrize("leaf_rows,patterns", [(True, True), (False, True), (True, False)])
@pytest.mark.parametrize("jc", [True, False])
def test_grouped_scoring_kernels_against_oracle(ops, primate_genome, jc, leaf_rows, patterns):
    """The grouped visiting order and its three scoring kernels (site patterns for two leaves, state-sorted rows for a
    leaf + an internal node, the bilinear form for two internal nodes) are what large runs use; `force_sorted` puts a
    small run on that path so that it can be checked against the oracle: gaps, ambiguity codes other than gaps (class
    "other" of the leaf sort), more sites than one 1024-site tile, and a site count that is not a multiple of anything."""
    g = np.concatenate([primate_genome[:9], primate_genome[:9, :403]], axis=1)   # 1301 sites
    rng = np.random.default_rng(4)
    amb = rng.random(g.shape[:2]) < 0.03
    g[amb] = np.array([1.0, 0.0, 1.0, 0.0])          # R = A|G: neither one-hot nor a gap
    g[rng.random(g.shape[:2]) < 0.03] = 1.0           # gaps
    N, K = g.shape[0], 96
    p = random_params(N, jc, seed=21)
    U = O.Uniforms.draw(N, K, seed=31)
    res, g_ref = oracle_param_grads(g, K, p, U)
    codes = ops.pack_alignment(dev(g))
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    sw = ops.Sweep(N, g.shape[1], K, jc)
    sw.set_uniforms(*gpu_uniforms(U))
    sw.set_option("force_sorted", 1.0)
    sw.set_option("leaf_rows", 1.0 if leaf_rows else 0.0)
    sw.set_option("leaf_patterns", 1.0 if patterns else 0.0)   # 0: cherries go through the generic kernel, site by site
    for it in range(2):   # the second sweep replays the captured graph
        elbo = sw.forward(codes, dev(lam_l), dev(lam_r), None if jc else dev(Q), dev(pi.reshape(-1)))
        out = {k: sw.output(k).cpu().numpy().copy() for k in
               ("log_weights", "log_likelihood", "log_likelihood_tilde", "log_likelihood_R", "left_branches",
                "right_branches", "v_minus", "ancestors", "left_ref", "right_ref", "leaf_counts")}
        out["elbo"] = float(elbo.item())
        grads = [None if t is None else t.cpu().numpy() for t in sw.backward(1.0)]
        sw.check_status()
        compare_forward(out, res, N, K)
        compare_grads(grads, g_ref, jc)


@pytest.mark.parametrize("leaf_rows", [True, False])
@pytest.mark.parametrize("jc", [True, False])
def test_scoring_kernels_fall_back_to_one_log_per_site(ops, jc, leaf_rows):
    """Branch rates of 1e80 give transition matrices I + O(1e-80): a site where the children disagree has a likelihood of
    1e-80 .. 1e-240 (a normal double, its log is finite), but the product of a thread's four such sites -- or of a
    run's tiles -- leaves the double range.  The scoring kernels then find the running product outside
    [2^-959, 2^1024), poison the particle and redo it with one log per site, like the reference (score.cu:
    rows_slow / score_slow): the results must still match the oracle.  i.i.d. nucleotides (three sites out of four
    disagree), gaps, four taxa (a deeper tree would underflow per SITE, in the reference too)."""
    g = synthetic_genome(4, 1301, seed=5, gaps=0.02)
    N, K = 4, 48
    p = O.Params.init(N, jc)
    p.left_branches_param = torch.full((N - 1,), float(np.log(1e80)), dtype=torch.float64)
    p.right_branches_param = torch.full((N - 1,), float(np.log(1e80)), dtype=torch.float64)
    U = O.Uniforms.draw(N, K, seed=7)
    lam_l, lam_r, Q, pi = O.model_from_params(p)
    res = O.sweep(g, K, lam_l, lam_r, Q, pi, U)
    assert np.isfinite(np.asarray(res.log_weights)).all() and float(np.asarray(res.log_likelihood).min()) < -4e5
    codes = ops.pack_alignment(dev(g))
    sw = ops.Sweep(N, g.shape[1], K, jc, keep_for_backward=False)
    sw.set_uniforms(*gpu_uniforms(U))
    sw.set_option("force_sorted", 1.0)
    sw.set_option("leaf_rows", 1.0 if leaf_rows else 0.0)
    elbo = sw.forward(codes, dev(lam_l.detach()), dev(lam_r.detach()), None if jc else dev(Q.detach()), dev(pi.detach().reshape(-1)))
    sw.check_status()
    out = {k: sw.output(k).cpu().numpy().copy() for k in
           ("log_weights", "log_likelihood", "log_likelihood_tilde", "log_likelihood_R", "left_branches",
            "right_branches", "v_minus", "ancestors", "left_ref", "right_ref", "leaf_counts")}
    out["elbo"] = float(elbo.item())
    compare_forward(out, res, N, K)
