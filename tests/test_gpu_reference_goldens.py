"""The CUDA path against outputs of the reference's OWN source.

tests/golden/ref_sweeps.npz was produced by importing /root/reference/vcsmc.py and vncsmc.py unmodified under an
eager TensorFlow stand-in (tests/golden/tf_shim.py, driven by tests/golden/make_golden.py) and running
``sample_phylogenies`` (vcsmc.py:406-451 / vncsmc.py:511-555) plus the autodiff of ``cost`` (vcsmc.py:488-491) on
primate.p subsets with injected randomness.  Here the same inputs go through the C ABI (lazy and eager schedules)
and are compared with those files directly -- no oracle in between.  Bars: integer tables bit-exact; log-weights,
likelihoods, ELBO 1e-9 relative; gradients w.r.t. the reference's variables 1e-7 relative.
"""
import numpy as np
import pytest
import torch

from vcsmc_test_helpers import RefCase, gpu_uniforms, ref_case_names

pytestmark = pytest.mark.gpu
RTOL = 1e-9
F64 = torch.float64


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from phylo_b200 import ops as _ops
    return _ops


def dev(x):
    return torch.as_tensor(x).cuda().contiguous()


def model_on_device(c):
    """The reference's parameterisation (vcsmc.py:119-148) of the golden variable values, as torch autograd leaves."""
    from phylo_b200.vcsmc import VCSMC
    import types
    args = types.SimpleNamespace(M=c.M or 1, branch_prior=np.log(10), jcmodel=c.jc, nested=c.M is not None,
                                 optimizer="GradientDescentOptimizer")
    m = VCSMC({"taxa": ["S%d" % i for i in range(c.N)], "genome": c.genome}, c.K, args, seed=1)
    with torch.no_grad():
        m.left_branches_var.copy_(dev(c["var_left_branches_param"]))
        m.right_branches_var.copy_(dev(c["var_right_branches_param"]))
        if not c.jc:
            m.y_q.copy_(dev(c["var_Qmatrix"]))
            m.y_station.copy_(dev(c["var_Stationary_probs"]))
    return m


def check_forward(c, sw, elbo):
    assert float(elbo) == pytest.approx(float(c["elbo"]), rel=RTOL)
    out = lambda k: sw.output(k).cpu().numpy()
    np.testing.assert_array_equal(out("ancestors")[1:], c["ancestors"][1:])
    np.testing.assert_array_equal(out("v_minus"), c["v_minus"])
    for name in ("log_weights", "log_likelihood", "log_likelihood_tilde", "log_likelihood_R"):
        np.testing.assert_allclose(out(name), c[name], rtol=RTOL, atol=1e-9, err_msg=name)
    for name in ("left_branches", "right_branches"):
        np.testing.assert_allclose(out(name), c[name], rtol=1e-13, err_msg=name)


@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("name", ref_case_names())
def test_cuda_sweep_matches_reference_source(ops, name, lazy):
    c = RefCase(name)
    m = model_on_device(c)
    sw = m._sweep_for(c.S, True)
    sw.set_uniforms(*gpu_uniforms(c.uniforms()))
    sw.set_option("lazy", 1.0 if lazy else 0.0)
    lam_l, lam_r, Q, pi = m._model()
    elbo = ops.sweep_elbo(sw, m.codes, lam_l, lam_r, None if c.jc else Q, pi)
    (-elbo).backward()                                        # cost = -ELBO (vcsmc.py:447)
    sw.check_status()
    check_forward(c, sw, elbo.detach())
    # pair choices (vcsmc.py:304-305) and kept order, as positions in the pre-merge forest
    rem = sw.rem_positions()
    lref, rref = sw.output("left_ref").cpu().numpy(), sw.output("right_ref").cpu().numpy()
    forest = np.tile(np.arange(c.N, dtype=np.int64), (c.K, 1))
    ar = np.arange(c.K)
    for r in range(c.N - 1):
        if r > 0:
            forest = forest[c["ancestors"][r]]
        np.testing.assert_array_equal(lref[r], forest[ar, c.coal(r)[:, 0]])
        np.testing.assert_array_equal(rref[r], forest[ar, c.coal(r)[:, 1]])
        # the lazy schedule rebuilds rows (and kept positions) only for particles whose normalised weight is not zero
        # in double precision: nobody else can be resampled or carry a gradient
        alive = np.exp(c["log_weights"][r] - c["log_weights"][r].max()) > 0 if lazy else np.ones(c.K, dtype=bool)
        if r < c.N - 2:
            np.testing.assert_array_equal(rem[r][alive], c.rem(r)[alive])
        new_id = (c.N + r * c.K + ar)[:, None]
        forest = np.concatenate([np.take_along_axis(forest, c.rem(r).astype(np.int64), axis=1), new_id], axis=1)
    for v, g in zip(m.trainable_variables(), c.grads_elbo()):
        np.testing.assert_allclose(-v.grad.cpu().numpy(), g, rtol=1e-7, atol=1e-9 * max(np.abs(g).max(), 1e-300))


@pytest.mark.parametrize("name", ref_case_names(nested=True))
def test_cuda_nested_sweep_matches_reference_source(ops, name):
    c = RefCase(name)
    m = model_on_device(c)
    sw = m._sweep_for(c.S, True)
    U = c.uniforms()
    sw.set_uniforms_nested(dev(np.concatenate([a.reshape(-1) for a in U.look_bl])),
                           dev(np.concatenate([a.reshape(-1) for a in U.look_br])), dev(U.cat), dev(U.res))
    lam_l, lam_r, Q, pi = m._model()
    elbo = ops.sweep_elbo(sw, m.codes, lam_l, lam_r, None if c.jc else Q, pi)
    (-elbo).backward()
    sw.check_status()
    check_forward(c, sw, elbo.detach())
    np.testing.assert_array_equal(sw.output("choice").cpu().numpy(), c["choices"])       # vncsmc.py:298
    for v, g in zip(m.trainable_variables(), c.grads_elbo()):
        np.testing.assert_allclose(-v.grad.cpu().numpy(), g, rtol=1e-7, atol=1e-9 * max(np.abs(g).max(), 1e-300))
