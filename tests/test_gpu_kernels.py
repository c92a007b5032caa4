"""GPU parity of each kernel behind the C ABI against the CPU oracle (and the csmc.py golden vectors)."""
import os

import numpy as np
import pytest
import torch

from oracle import vcsmc_oracle as O
from vcsmc_test_helpers import synthetic_genome

pytestmark = pytest.mark.gpu
RTOL = 1e-9  # north_star: log-likelihoods within 1e-9 relative in fp64 mode


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from phylo_b200 import ops as _ops
    return _ops


def dev(x):
    return torch.as_tensor(x).cuda().contiguous()


# ---------------------------------------------------------------- (a) loader
def test_pack_alignment_and_gather(ops, primate_genome):
    codes = ops.pack_alignment(dev(primate_genome))
    g = primate_genome.astype(np.int64)
    ref = (g[..., 0] | (g[..., 1] << 1) | (g[..., 2] << 2) | (g[..., 3] << 3)).astype(np.uint8)
    np.testing.assert_array_equal(codes.cpu().numpy(), ref)
    assert (ref == 15).sum() == 30  # primate.p has 30 gap characters
    idx = np.random.default_rng(0).permutation(898)[:100].astype(np.int32)
    sub = ops.gather_sites(codes, dev(idx))
    np.testing.assert_array_equal(sub.cpu().numpy(), ref[:, idx])


def test_pack_alignment_rejects_non_masks(ops):
    from phylo_b200._lib import VcsmcError
    g = synthetic_genome(3, 20)
    g[1, 4, 2] = 0.5
    with pytest.raises(VcsmcError):
        ops.pack_alignment(dev(g))
    g = synthetic_genome(3, 20)
    g[2, 7] = 0.0  # all-zero site: log(0) in the reference
    with pytest.raises(VcsmcError):
        ops.pack_alignment(dev(g))


# ---------------------------------------------------------------- (b) transition matrices
@pytest.mark.parametrize("jc", [True, False])
def test_transition_fwd(ops, jc):
    rng = np.random.default_rng(1)
    t = np.concatenate([[1e-300, 1e-12, 1e-6, 0.013, 0.1, 0.5, 1.0, 7.3, 40.0, 700.0], rng.exponential(0.1, 500)])
    Q = O.jc_Q() if jc else O.get_Q(torch.from_numpy(rng.normal(size=(4, 4))))
    P = ops.transition_fwd(None if jc else dev(Q), dev(t), jc).cpu()
    ref = O.transition_matrices(Q, torch.from_numpy(t))
    np.testing.assert_allclose(P.numpy(), ref.numpy(), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(P.sum(dim=2).numpy(), 1.0, atol=1e-13)


def test_transition_fwd_initial_gtr_Q(ops):
    """The reference's initial 'GTR' Q (off-diagonal 1/3, diagonal -1) has a triple eigenvalue."""
    Q = O.get_Q(torch.full((4, 4), 0.25, dtype=torch.float64))
    t = np.array([0.05, 0.2, 3.0])
    P = ops.transition_fwd(dev(Q), dev(t), False).cpu()
    np.testing.assert_allclose(P.numpy(), O.transition_matrices(Q, torch.from_numpy(t)).numpy(), rtol=1e-13, atol=1e-16)


def test_transition_bwd_general(ops):
    rng = np.random.default_rng(2)
    n = 300
    Q = O.get_Q(torch.from_numpy(rng.normal(size=(4, 4)))).requires_grad_(True)
    t = torch.from_numpy(np.concatenate([[1e-8, 3.0, 25.0], rng.exponential(0.1, n - 3)])).requires_grad_(True)
    G = torch.from_numpy(rng.normal(size=(n, 4, 4)))
    (O.transition_matrices(Q, t) * G).sum().backward()
    dt, dQ = ops.transition_bwd(dev(Q.detach()), dev(t.detach()), dev(G), False)
    np.testing.assert_allclose(dt.cpu().numpy(), t.grad.numpy(), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(dQ.sum(dim=0).cpu().numpy(), Q.grad.numpy(), rtol=1e-10, atol=1e-12)


def test_transition_custom_op_autograd(ops):
    rng = np.random.default_rng(3)
    Qc = O.get_Q(torch.from_numpy(rng.normal(size=(4, 4))))
    tc = torch.from_numpy(rng.exponential(0.2, 50))
    G = torch.from_numpy(rng.normal(size=(50, 4, 4)))
    for jc in (False, True):
        Qo = (O.jc_Q() if jc else Qc).clone().requires_grad_(True)
        to = tc.clone().requires_grad_(True)
        (O.transition_matrices(Qo, to) * G).sum().backward()
        Qg = dev(Qo.detach()).requires_grad_(True)
        tg = dev(tc).requires_grad_(True)
        (torch.ops.vcsmc.transition(Qg, tg, jc) * dev(G)).sum().backward()
        np.testing.assert_allclose(tg.grad.cpu().numpy(), to.grad.numpy(), rtol=1e-10, atol=1e-13)
        if not jc:
            np.testing.assert_allclose(Qg.grad.cpu().numpy(), Qo.grad.numpy(), rtol=1e-10, atol=1e-12)


# ---------------------------------------------------------------- (c) merge
def test_merge_matches_csmc_golden(ops, golden_dir):
    """The merge kernel reproduces the reference's own csmc.py:300-309 outputs (tests/golden/csmc_merge.npz)."""
    z = np.load(os.path.join(golden_dir, "csmc_merge.npz"))
    K, S = z["L_l"].shape[0], z["L_l"].shape[1]
    Q = dev(z["Q"])
    P_l = ops.transition_fwd(Q, dev(z["b_l"]), False)
    P_r = ops.transition_fwd(Q, dev(z["b_r"]), False)
    new, ell = torch.ops.vcsmc.merge(dev(z["L_l"]), dev(z["L_r"]), P_l, P_r, dev(z["prior"]))
    np.testing.assert_allclose(new.cpu().numpy(), z["merged"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(ell.cpu().numpy(), np.log(z["merged"] @ z["prior"]).sum(axis=1), rtol=1e-12)


@pytest.mark.parametrize("jc", [True, False])
@pytest.mark.parametrize("S", [1, 37, 1024, 2500])
def test_merge_fwd_leaf_and_pool_children(ops, jc, S):
    """Leaf x leaf, leaf x node and node x node children, ragged site counts (partial tiles), gaps."""
    rng = np.random.default_rng(10 + S)
    N, K = 5, 7
    g = synthetic_genome(N, S, seed=S, gaps=0.05)
    codes = ops.pack_alignment(dev(g))
    Q = O.jc_Q() if jc else O.get_Q(torch.from_numpy(rng.normal(size=(4, 4)) * 0.5))
    pi = torch.full((4,), 0.25, dtype=torch.float64) if jc else O.get_stationary_probs(torch.from_numpy(rng.normal(size=4))).reshape(-1)
    dense = rng.random((K, S, 4))
    pool = torch.zeros((2 * K, S, 4), dtype=torch.float64, device="cuda")
    pool[:K] = dev(dense)
    lsrc = np.array([-1, -2, -3, 0, 1, 2, 3], dtype=np.int32)      # leaves 0,1,2 then pool slots
    rsrc = np.array([-2, 3, -5, -1, 4, 5, 6], dtype=np.int32)
    dst = (K + np.arange(K)).astype(np.int32)
    b = rng.exponential(0.1, (2, K))
    P = torch.cat([ops.transition_fwd(None if jc else dev(Q), dev(b[0]), jc).reshape(K, 16),
                   ops.transition_fwd(None if jc else dev(Q), dev(b[1]), jc).reshape(K, 16)], dim=1).contiguous()
    ell = ops.merge_fwd(codes, pool, dev(lsrc), dev(rsrc), dev(dst), P, dev(pi), S, jc)

    def child(src, k):
        return torch.from_numpy(g[-src - 1] if src < 0 else dense[src])

    L_l = torch.stack([child(lsrc[k], k) for k in range(K)])
    L_r = torch.stack([child(rsrc[k], k) for k in range(K)])
    ref = O.merge(L_l, L_r, torch.from_numpy(b[0]), torch.from_numpy(b[1]), Q)
    np.testing.assert_allclose(pool[K:].cpu().numpy(), ref.numpy(), rtol=1e-12, atol=0)
    np.testing.assert_allclose(ell.cpu().numpy(), torch.log(ref @ pi).sum(dim=1).numpy(), rtol=RTOL)


def test_merge_custom_op_autograd(ops):
    """Reverse pruning of one merge == torch autograd of the oracle formula."""
    rng = np.random.default_rng(5)
    K, S = 6, 1500
    Q = O.get_Q(torch.from_numpy(rng.normal(size=(4, 4)) * 0.5))
    b = torch.from_numpy(rng.exponential(0.1, (2, K)))
    P_l, P_r = O.transition_matrices(Q, b[0]), O.transition_matrices(Q, b[1])
    pi = O.get_stationary_probs(torch.from_numpy(rng.normal(size=4))).reshape(-1)
    L_l, L_r = torch.from_numpy(rng.random((K, S, 4))), torch.from_numpy(rng.random((K, S, 4)))
    g_new, g_ell = torch.from_numpy(rng.normal(size=(K, S, 4))), torch.from_numpy(rng.normal(size=K))
    cpu_in = [x.clone().requires_grad_(True) for x in (L_l, L_r, P_l, P_r, pi)]
    new = torch.matmul(cpu_in[0], cpu_in[2]) * torch.matmul(cpu_in[1], cpu_in[3])
    ((new * g_new).sum() + (torch.log(new @ cpu_in[4]).sum(dim=1) * g_ell).sum()).backward()
    gpu_in = [dev(x).requires_grad_(True) for x in (L_l, L_r, P_l, P_r, pi)]
    new_g, ell_g = torch.ops.vcsmc.merge(*gpu_in)
    ((new_g * dev(g_new)).sum() + (ell_g * dev(g_ell)).sum()).backward()
    for a, b_ in zip(gpu_in, cpu_in):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b_.grad.numpy(), rtol=1e-9, atol=1e-9)


# ---------------------------------------------------------------- (d) proposal / resampling
@pytest.mark.parametrize("n", [2, 3, 12, 33, 64, 200])
def test_propose_pairs_bit_exact(ops, n):
    rng = np.random.default_rng(n)
    K = 1000
    u = rng.random((K, n), dtype=np.float32)
    u[:50, : min(n, 4)] = np.float32(0.5)   # exact ties (tf.nn.top_k tie rule incl. the duplicate quirk)
    u[50:60, 0] = 0.0                       # u = 0 -> z = -inf
    coal, rem = ops.propose_pairs(dev(u))
    c_ref, r_ref = O.propose_pairs(u)
    np.testing.assert_array_equal(coal.cpu().numpy(), c_ref)
    np.testing.assert_array_equal(rem.cpu().numpy(), r_ref)


@pytest.mark.parametrize("K", [1, 16, 1000, 4096, 65536])
def test_resample_bit_exact(ops, K):
    rng = np.random.default_rng(K)
    for spread in (0.5, 30.0, 3000.0):     # near-uniform, skewed, degenerate (ESS ~ 1)
        lw = rng.normal(size=K) * spread - 7000.0
        u = rng.random(K)
        u[0] = 0.0
        idx, lse, ess = ops.resample(dev(lw), dev(u))
        np.testing.assert_array_equal(idx.cpu().numpy().astype(np.int64), O.resample_indices(lw, u))
        m = lw.max()
        assert float(lse) == pytest.approx(m + np.log(np.exp(lw - m).sum()), rel=1e-13)
        w = np.exp(lw - m)
        assert float(ess) == pytest.approx(w.sum() ** 2 / (w * w).sum(), rel=1e-10)


def test_resample_custom_op_and_minus_inf(ops):
    lw = np.array([-np.inf, 5.0, 5.0, -np.inf])
    idx, lse, ess = torch.ops.vcsmc.resample(dev(lw), dev(np.array([0.0, 0.49, 0.51, 0.9999])))
    assert idx.cpu().tolist() == [1, 1, 2, 2]
    assert float(ess) == pytest.approx(2.0)


def test_philox_uniform_ranges_and_shard_invariance(ops):
    """Uniforms are keyed by LOGICAL particle index: a shard sees exactly the slice of the full stream."""
    K, n, r = 4096, 13, 5
    full = ops.philox_step_uniforms(1234, r, 0, K, n)
    part = ops.philox_step_uniforms(1234, r, 1024, 512, n)
    for f, p in zip(full, part):
        assert torch.equal(f[1024:1536], p)
    u_pair, u_bl, u_br, u_res = [x.cpu().numpy() for x in full]
    assert u_pair.dtype == np.float32 and 0.0 <= u_pair.min() and u_pair.max() < 1.0
    assert u_bl.min() > 0.0 and u_bl.max() < 1.0 and 0.0 <= u_res.min() and u_res.max() < 1.0
    assert abs(u_pair.mean() - 0.5) < 0.01 and abs(u_bl.mean() - 0.5) < 0.02
    other = ops.philox_step_uniforms(1234, r + 1, 0, K, n)
    assert not torch.equal(other[1], full[1])
