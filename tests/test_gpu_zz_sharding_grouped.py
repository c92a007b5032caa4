"""Regression: particle sharding with enough particles per rank that the scoring kernel runs on the GROUPED order
(K/G * tiles >= 9472).  The grouping tables are carved for K particles but used for K/G: a rank must clear the table it
actually uses (a stale count table corrupted the visiting order on multi-GPU runs with large K)."""
import numpy as np
import pytest

from test_gpu_particle_sharding import RTOL, _oracle, _run

pytestmark = pytest.mark.gpu


def test_particle_sharding_grouped_order():
    res, g_ref, N, K = _oracle(False, "flat_grouped")
    outs = _run(2, False, "flat_grouped", 29770)
    for o in outs:
        np.testing.assert_array_equal(o["ancestors"][1:], res.ancestors[1:])
        np.testing.assert_allclose(o["log_weights"], res.log_weights.detach().numpy(), rtol=RTOL)
        assert o["elbo"] == pytest.approx(float(res.elbo), rel=RTOL)
        np.testing.assert_allclose(o["grads"], g_ref, rtol=1e-7, atol=1e-9 * np.abs(g_ref).max())
